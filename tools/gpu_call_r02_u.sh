#!/usr/bin/env bash
# round 2, call u: column statistics by the epilogue group instead of Gram MMAs; transition-backward ILP; stem maxpool diet
set -u
out=gpurun_out/r02u
mkdir -p "$out"
timeout 900 python -m pytest tests -x -q -m gpu > "$out/pytest_gpu.log" 2>&1; tail -6 "$out/pytest_gpu.log"
for v in "colstats:" "gram:RXB_DBG_GRAM_STATS=1" ; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 120 python bench.py --quick --steps 10 --warmup 3 > "$out/bench_quick_$name.json" 2> "$out/bench_quick_$name.err"; echo "$name $(cat $out/bench_quick_$name.json | python -c 'import json,sys; d=json.load(sys.stdin); print(d["ms_per_step"], d["value"], d.get("gpu_launches"))') $(tail -1 $out/bench_quick_$name.err | cut -c1-200)"
done
{
python tools/bench_conv.py one fwd 128 128 128 224 256 128 1 1 1
python tools/bench_conv.py one fwd 128 128 128 64 256 128 1 1 1
python tools/bench_conv.py one fwd 128 32 32 992 1024 128 1 1 1
RXB_DBG_GRAM_STATS=1 python tools/bench_conv.py one fwd 128 128 128 224 256 128 1 1 1
RXB_DBG_GRAM_STATS=1 python tools/bench_conv.py one fwd 128 128 128 64 256 128 1 1 1
} > "$out/times.log" 2>&1; cat "$out/times.log"
