#!/usr/bin/env bash
set -u
out=gpurun_out/r02f
mkdir -p "$out"
( time timeout 600 python -m pytest tests/test_gpu_resnet.py -m gpu -q -s -x ) > "$out/pytest_resnet.log" 2>&1; echo "pytest rc=$?"; tail -40 "$out/pytest_resnet.log" | cut -c1-600
