#!/usr/bin/env bash
set -u
out=gpurun_out/r02k
mkdir -p "$out"
for v in "default:" "pdl_graph:RXB_PDL=1" ; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 120 python bench.py --quick --steps 10 --warmup 3 > "$out/bench_quick_$name.json" 2> "$out/bench_quick_$name.err"; echo "$name $(cat $out/bench_quick_$name.json | python -c 'import json,sys; d=json.load(sys.stdin); print(d["ms_per_step"], d["value"])') $(tail -1 $out/bench_quick_$name.err | cut -c1-200)"
done
RXB_PDL=1 timeout 120 python bench.py --quick --no-graph --steps 10 --warmup 3 > "$out/bench_quick_pdl_nograph.json" 2>/dev/null; echo "pdl_nograph $(cat $out/bench_quick_pdl_nograph.json | python -c 'import json,sys; d=json.load(sys.stdin); print(d["ms_per_step"], d["value"])')"
timeout 120 python bench.py --quick --no-graph --steps 10 --warmup 3 > "$out/bench_quick_nograph.json" 2>/dev/null; echo "nograph $(cat $out/bench_quick_nograph.json | python -c 'import json,sys; d=json.load(sys.stdin); print(d["ms_per_step"], d["value"])')"
