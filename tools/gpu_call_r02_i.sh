#!/usr/bin/env bash
set -u
out=gpurun_out/r02i
mkdir -p "$out"
( time timeout 600 python -m pytest tests/test_gpu_resnet.py -m gpu -q -s -x ) > "$out/pytest_resnet.log" 2>&1; echo "pytest rc=$?"; grep -E "^ResNet-50 TwoSitesNN train step B|passed|failed|Error" "$out/pytest_resnet.log" | cut -c1-400
timeout 300 python tools/bench_resnet.py --batch 32 --steps 10 > "$out/bench_resnet.json" 2> "$out/bench_resnet.err"; echo "rc=$?"; cat "$out/bench_resnet.json"; tail -3 "$out/bench_resnet.err"
