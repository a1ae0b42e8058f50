#!/usr/bin/env bash
set -u
out=gpurun_out/r02e
mkdir -p "$out"
( time timeout 600 python -m pytest tests/test_gpu_resnet.py tests/test_gpu_conv.py -m gpu -q -s -k "resnet or reference or oracle or shim or fwd" ) > "$out/pytest_resnet.log" 2>&1; echo "pytest rc=$?"; tail -25 "$out/pytest_resnet.log" | cut -c1-300
