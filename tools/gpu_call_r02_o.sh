#!/usr/bin/env bash
# round 2, call o: where the fused 3x3 kernel spends its time - per-block timings, CTA-0 timeline
set -u
out=gpurun_out/r02o
mkdir -p "$out"
{
for g in "128 128" "64 64" "32 32" "16 16"; do
  set -- $g
  python tools/bench_conv.py one dgrad 128 $1 $2 32 128 128 3 0 0
  python tools/bench_conv.py one dgrad 128 $1 $2 32 128 128 3 0 1
  python tools/bench_conv.py one wgrad 128 $1 $2 128 128 32 3 1
done
python tools/bench_conv.py one dgrad 128 128 128 128 224 256 1 2 0
python tools/bench_conv.py one dgrad 128 128 128 128 224 256 1 2 1
python tools/bench_conv.py one dgrad 128 32 32 128 992 1024 1 2 0
python tools/bench_conv.py one dgrad 128 32 32 128 992 1024 1 2 1
python tools/bench_conv.py one dgrad 128 16 16 128 992 1024 1 2 0
python tools/bench_conv.py one dgrad 128 16 16 128 992 1024 1 2 1
} > "$out/times.log" 2>&1; cat "$out/times.log"
RXB_DBG_TIMELINE=1 python tools/bench_conv.py one dgrad 128 128 128 32 128 128 3 0 1 > "$out/timeline_fused3.log" 2>&1; head -45 "$out/timeline_fused3.log"
RXB_DBG_TIMELINE=1 python tools/bench_conv.py one dgrad 128 128 128 32 128 128 3 0 0 > "$out/timeline_dgrad3.log" 2>&1; head -45 "$out/timeline_dgrad3.log"
