#!/usr/bin/env bash
# round 2, call af: shared staging buffer for forward 1x1 with Cin > 128 only - step A/B
set -u
out=gpurun_out/r02af
mkdir -p "$out"
for i in 1 2 3; do
  RXB_DBG_STG1=0 timeout 120 python bench.py --quick --steps 10 --warmup 3 > "$out/a$i.json" 2>/dev/null; echo "two buffers $i $(python -c 'import json,sys; d=json.load(open(sys.argv[1])); print(d["ms_per_step"])' $out/a$i.json)"
  timeout 120 python bench.py --quick --steps 10 --warmup 3 > "$out/b$i.json" 2> "$out/b$i.err"; echo "one buffer (Cin>128) $i $(python -c 'import json,sys; d=json.load(open(sys.argv[1])); print(d["ms_per_step"])' $out/b$i.json) $(tail -1 $out/b$i.err | cut -c1-150)"
done
