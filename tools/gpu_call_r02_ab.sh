#!/usr/bin/env bash
# round 2, call ab: alternating traversal direction of consecutive kernels (L2 reuse at kernel boundaries) - parity, A/B
set -u
out=gpurun_out/r02ab
mkdir -p "$out"
timeout 600 python -m pytest tests/test_gpu_densenet.py tests/test_gpu_aa_regime.py -x -q > "$out/pytest_net.log" 2>&1; tail -3 "$out/pytest_net.log"
for i in 1 2; do
  RXB_DBG_NO_SNAKE=1 timeout 120 python bench.py --quick --steps 10 --warmup 3 > "$out/a$i.json" 2>/dev/null; echo "same direction $i $(python -c 'import json,sys; d=json.load(open(sys.argv[1])); print(d["ms_per_step"])' $out/a$i.json)"
  timeout 120 python bench.py --quick --steps 10 --warmup 3 > "$out/b$i.json" 2> "$out/b$i.err"; echo "alternating $i $(python -c 'import json,sys; d=json.load(open(sys.argv[1])); print(d["ms_per_step"])' $out/b$i.json) $(tail -1 $out/b$i.err | cut -c1-150)"
done
