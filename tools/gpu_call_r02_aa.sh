#!/usr/bin/env bash
# round 2, call aa: grad_fixup folded into the fused 3x3 kernel's load path - parity, then same-box A/B of the step
set -u
out=gpurun_out/r02aa
mkdir -p "$out"
timeout 600 python -m pytest tests/test_gpu_conv.py -x -q -k "fixup_fold or fused_wgrad" > "$out/pytest_conv.log" 2>&1; tail -5 "$out/pytest_conv.log"
timeout 600 python -m pytest tests/test_gpu_densenet.py tests/test_gpu_aa_regime.py -x -q > "$out/pytest_net.log" 2>&1; tail -4 "$out/pytest_net.log"
for i in 1 2; do
  RXB_DBG_NO_FIXFOLD=1 timeout 120 python bench.py --quick --steps 10 --warmup 3 > "$out/a$i.json" 2>/dev/null; echo "separate fixup $i $(python -c 'import json,sys; d=json.load(open(sys.argv[1])); print(d["ms_per_step"], d["gpu_launches"])' $out/a$i.json)"
  timeout 120 python bench.py --quick --steps 10 --warmup 3 > "$out/b$i.json" 2> "$out/b$i.err"; echo "folded $i $(python -c 'import json,sys; d=json.load(open(sys.argv[1])); print(d["ms_per_step"], d["gpu_launches"])' $out/b$i.json) $(tail -1 $out/b$i.err | cut -c1-150)"
done
