#!/usr/bin/env bash
# round 2, call r: EPI 4 with the transform one tile ahead: parity, timings, timelines (fused 3x3, forward 3x3)
set -u
out=gpurun_out/r02r
mkdir -p "$out"
timeout 600 python -m pytest tests/test_gpu_conv.py -x -q -k "fused_wgrad" > "$out/pytest_conv.log" 2>&1; tail -3 "$out/pytest_conv.log"
timeout 600 python -m pytest tests/test_gpu_densenet.py tests/test_gpu_aa_regime.py -x -q > "$out/pytest_net.log" 2>&1; tail -4 "$out/pytest_net.log"
timeout 120 python bench.py --quick --steps 10 --warmup 3 > "$out/bench_quick.json" 2> "$out/bench_quick.err"; echo "step $(cat $out/bench_quick.json | python -c 'import json,sys; d=json.load(sys.stdin); print(d["ms_per_step"], d["value"])')"
{
for g in "128 128" "64 64" "32 32" "16 16"; do
  set -- $g
  python tools/bench_conv.py one dgrad 128 $1 $2 32 128 128 3 0 1
done
python tools/bench_conv.py one fwd 128 128 128 128 128 32 3 1 1
python tools/bench_conv.py one fwd 128 64 64 128 128 32 3 1 1
} > "$out/times.log" 2>&1; cat "$out/times.log"
RXB_DBG_TIMELINE=1 python tools/bench_conv.py one dgrad 128 128 128 32 128 128 3 0 1 > "$out/timeline_fused3.log" 2>&1; head -26 "$out/timeline_fused3.log"
RXB_DBG_TIMELINE=1 python tools/bench_conv.py one fwd 128 128 128 128 128 32 3 1 1 > "$out/timeline_fwd3.log" 2>&1; head -38 "$out/timeline_fwd3.log"
