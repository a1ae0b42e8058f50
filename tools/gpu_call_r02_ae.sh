#!/usr/bin/env bash
# round 2, call ae: ONE staging buffer shared by the two epilogue groups of the forward 1x1 (two more operand stages)
set -u
out=gpurun_out/r02ae
mkdir -p "$out"
RXB_DBG_STG1=1 timeout 600 python -m pytest tests/test_gpu_conv.py -x -q -k "conv_fwd or conv_matches or forward" > "$out/pytest_conv.log" 2>&1; tail -3 "$out/pytest_conv.log"
RXB_DBG_STG1=1 timeout 600 python -m pytest tests/test_gpu_densenet.py tests/test_gpu_aa_regime.py -x -q > "$out/pytest_net.log" 2>&1; tail -3 "$out/pytest_net.log"
for i in 1 2; do
  timeout 120 python bench.py --quick --steps 10 --warmup 3 > "$out/a$i.json" 2>/dev/null; echo "two buffers $i $(python -c 'import json,sys; d=json.load(open(sys.argv[1])); print(d["ms_per_step"])' $out/a$i.json)"
  RXB_DBG_STG1=1 timeout 120 python bench.py --quick --steps 10 --warmup 3 > "$out/b$i.json" 2> "$out/b$i.err"; echo "one buffer $i $(python -c 'import json,sys; d=json.load(open(sys.argv[1])); print(d["ms_per_step"])' $out/b$i.json) $(tail -1 $out/b$i.err | cut -c1-150)"
done
for args in "fwd 128 128 128 224 256 128 1 1 1" "fwd 128 128 128 64 256 128 1 1 1" "fwd 128 64 64 256 512 128 1 1 1" "fwd 128 32 32 992 1024 128 1 1 1"; do
  echo "A: $(python tools/bench_conv.py one $args)"
  echo "B: $(RXB_DBG_STG1=1 python tools/bench_conv.py one $args)"
done
