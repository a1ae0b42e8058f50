#!/usr/bin/env bash
set -u
out=gpurun_out/r02l
mkdir -p "$out"
timeout 200 python tools/repeat_step_diag.py 16 512 > "$out/repeat_16_512.log" 2>&1; cat "$out/repeat_16_512.log"
