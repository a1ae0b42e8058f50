#!/usr/bin/env bash
# round 2, call s: BatchNorm-backward reductions folded into the fused kernels' tails (no bn_bwd_finalize launches)
set -u
out=gpurun_out/r02s
mkdir -p "$out"
timeout 900 python -m pytest tests -x -q -m gpu > "$out/pytest_gpu.log" 2>&1; tail -6 "$out/pytest_gpu.log"
for v in "tail:" "notail:RXB_DBG_NO_BNTAIL=1" ; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 120 python bench.py --quick --steps 10 --warmup 3 > "$out/bench_quick_$name.json" 2> "$out/bench_quick_$name.err"; echo "$name $(cat $out/bench_quick_$name.json | python -c 'import json,sys; d=json.load(sys.stdin); print(d["ms_per_step"], d["value"], d.get("gpu_launches"))') $(tail -1 $out/bench_quick_$name.err | cut -c1-200)"
done
