#!/usr/bin/env bash
set -u
out=gpurun_out/r02j
mkdir -p "$out"
( time timeout 600 python -m pytest tests/test_gpu_aa_regime.py tests/test_gpu_densenet.py tests/test_gpu_shims.py -m gpu -q -x ) > "$out/pytest.log" 2>&1; echo "pytest rc=$?"; tail -3 "$out/pytest.log"
for v in "overlap:" "no_overlap:RXB_NO_OVERLAP=1" "overlap_nograph:RXB_NO_GRAPH=1" ; do
  name=${v%%:*}; envs=${v#*:}; extra=""; [ "$name" = "overlap_nograph" ] && extra="--no-graph"
  env $envs timeout 120 python bench.py --quick --steps 10 --warmup 3 $extra > "$out/bench_quick_$name.json" 2>/dev/null; echo "$name $(cat $out/bench_quick_$name.json | python -c 'import json,sys; d=json.load(sys.stdin); print(d["ms_per_step"], d["value"])')"
done
