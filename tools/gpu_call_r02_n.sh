#!/usr/bin/env bash
# round 2, call n: fused 3x3 weight gradient inside the 3x3 data-gradient kernel - parity, then the step
set -u
out=gpurun_out/r02n
mkdir -p "$out"
timeout 600 python -m pytest tests/test_gpu_conv.py -x -q -k "fused_wgrad or dgrad" > "$out/pytest_conv.log" 2>&1; tail -15 "$out/pytest_conv.log"
timeout 600 python -m pytest tests/test_gpu_densenet.py tests/test_gpu_aa_regime.py -x -q > "$out/pytest_net.log" 2>&1; tail -8 "$out/pytest_net.log"
for v in "fused:" "separate3:RXB_DBG_NO_WGFUSE3=1" ; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 120 python bench.py --quick --steps 10 --warmup 3 > "$out/bench_quick_$name.json" 2> "$out/bench_quick_$name.err"; echo "$name $(cat $out/bench_quick_$name.json | python -c 'import json,sys; d=json.load(sys.stdin); print(d["ms_per_step"], d["value"], d.get("loss"))') $(tail -1 $out/bench_quick_$name.err | cut -c1-200)"
done
