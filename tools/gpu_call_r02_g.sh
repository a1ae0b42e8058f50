#!/usr/bin/env bash
# experiment: weight-gradient time against the number of pixel-split CTAs (each adds its partial dW by L2 reduce-adds)
set -u
out=gpurun_out/r02g
mkdir -p "$out"
for cap in 0 96 64 48 32 24 16; do
  for spec in "128 16 16 128 128 32 3 1" "128 32 32 128 128 32 3 1" "128 64 64 128 128 32 3 1" "128 128 128 128 128 32 3 1" \
              "128 16 16 768 1024 128 1 1" "128 32 32 640 1024 128 1 1" "128 64 64 320 512 128 1 1" "128 128 128 224 256 128 1 1"; do
    echo -n "cap=$cap " >> "$out/wgrad_pix_sweep.log"
    RXB_DBG_WG_PIX=$cap timeout 60 python tools/bench_conv.py one wgrad $spec 2>&1 | tail -1 >> "$out/wgrad_pix_sweep.log"
  done
done
cat "$out/wgrad_pix_sweep.log"
