#!/usr/bin/env bash
# round 2, call ag: narrow store epilogue - statistics by a pass over the staged tile instead of warp shuffles
set -u
out=gpurun_out/r02ag
mkdir -p "$out"
timeout 600 python -m pytest tests/test_gpu_conv.py -x -q -k "conv_fwd" > "$out/pytest_conv.log" 2>&1; tail -3 "$out/pytest_conv.log"
timeout 600 python -m pytest tests/test_gpu_densenet.py tests/test_gpu_aa_regime.py tests/test_gpu_resnet.py -x -q > "$out/pytest_net.log" 2>&1; tail -3 "$out/pytest_net.log"
for i in 1 2; do
  RXB_DBG_SHFL_STATS=1 timeout 120 python bench.py --quick --steps 10 --warmup 3 > "$out/a$i.json" 2>/dev/null; echo "shuffles $i $(python -c 'import json,sys; d=json.load(open(sys.argv[1])); print(d["ms_per_step"])' $out/a$i.json)"
  timeout 120 python bench.py --quick --steps 10 --warmup 3 > "$out/b$i.json" 2> "$out/b$i.err"; echo "column pass $i $(python -c 'import json,sys; d=json.load(open(sys.argv[1])); print(d["ms_per_step"])' $out/b$i.json) $(tail -1 $out/b$i.err | cut -c1-150)"
done
for args in "fwd 128 128 128 128 128 32 3 1 1" "fwd 128 64 64 128 128 32 3 1 1" "fwd 128 32 32 128 128 32 3 1 1" "fwd 128 256 256 32 32 64 4 0 1"; do
  echo "A: $(RXB_DBG_SHFL_STATS=1 python tools/bench_conv.py one $args)"
  echo "B: $(python tools/bench_conv.py one $args)"
done
