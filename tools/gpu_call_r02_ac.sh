#!/usr/bin/env bash
# round 2, call ac: in-place buffer count A/B after the role split (RXB_DBG_NX), same box
set -u
out=gpurun_out/r02ac
mkdir -p "$out"
for v in "nx3:" "nx4:RXB_DBG_NX=4" "nx2:RXB_DBG_NX=2" "nx3b:" ; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 120 python bench.py --quick --steps 10 --warmup 3 > "$out/$name.json" 2> "$out/$name.err"; echo "$name $(python -c 'import json,sys; d=json.load(open(sys.argv[1])); print(d["ms_per_step"])' $out/$name.json) $(tail -1 $out/$name.err | cut -c1-150)"
done
