#!/usr/bin/env bash
# same-box A/B of two builds of librxb: tools/exp/ab/librxb_a.so (A) against the in-tree library (B); box-to-box variance
# of the step is ~3 %, so only same-call comparisons decide
set -u
out=gpurun_out/ab_$(date +%H%M%S)
mkdir -p "$out"
A=$PWD/tools/exp/ab/librxb_a.so
for i in 1 2; do
  RXB_LIB=$A timeout 120 python bench.py --quick --steps 10 --warmup 3 > "$out/a$i.json" 2>/dev/null; echo "A$i $(python -c 'import json,sys; d=json.load(open(sys.argv[1])); print(d["ms_per_step"])' $out/a$i.json)"
  timeout 120 python bench.py --quick --steps 10 --warmup 3 > "$out/b$i.json" 2>/dev/null; echo "B$i $(python -c 'import json,sys; d=json.load(open(sys.argv[1])); print(d["ms_per_step"])' $out/b$i.json)"
done
for args in "dgrad 128 128 128 128 224 256 1 2 1" "dgrad 128 32 32 128 992 1024 1 2 1" "dgrad 128 128 128 32 128 128 3 0 1" "fwd 128 128 128 128 128 32 3 1 1" "fwd 128 128 128 224 256 128 1 1 1"; do
  echo "A: $(RXB_LIB=$A python tools/bench_conv.py one $args)"
  echo "B: $(python tools/bench_conv.py one $args)"
done
