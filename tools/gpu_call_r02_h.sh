#!/usr/bin/env bash
set -u
out=gpurun_out/r02h
mkdir -p "$out"
timeout 300 python tools/bench_resnet.py --batch 32 --steps 10 --library > "$out/bench_resnet.json" 2> "$out/bench_resnet.err"; echo "rc=$?"; cat "$out/bench_resnet.json"; tail -3 "$out/bench_resnet.err"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 900 --csv --log-file "$out/launches_resnet.csv" \
  python tools/bench_resnet.py --batch 32 --steps 1 > "$out/ncu_resnet.log" 2>&1; echo "ncu rc=$?"
