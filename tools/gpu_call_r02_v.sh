#!/usr/bin/env bash
# round 2, call v: buffer-count A/B for the fused data-gradient kernels; ncu --set full of the conv launches (one each)
set -u
out=gpurun_out/r02v
mkdir -p "$out"
for v in "nx3:" "nx4:RXB_DBG_NX=4" "nx2:RXB_DBG_NX=2" ; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 120 python bench.py --quick --steps 10 --warmup 3 > "$out/bench_quick_$name.json" 2> "$out/bench_quick_$name.err"; echo "$name $(cat $out/bench_quick_$name.json | python -c 'import json,sys; d=json.load(sys.stdin); print(d["ms_per_step"], d["value"], d.get("gpu_launches"))') $(tail -1 $out/bench_quick_$name.err | cut -c1-200)"
done
timeout 300 python tools/bench_conv.py prof 128 > "$out/prof_plain.log" 2>&1 || { tail -5 "$out/prof_plain.log"; exit 1; }
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:conv_ -o "$out/prof_conv_b128" -f python tools/bench_conv.py prof 128 > "$out/ncu_conv_b128.log" 2>&1
tail -2 "$out/ncu_conv_b128.log"; ls -la "$out"
