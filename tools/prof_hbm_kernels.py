#!/usr/bin/env python
"""Profiling driver (ncu) for the HBM-bound kernels timed alone in bench.py: a few launches of the statistics kernel and
of the fused D4 loader on 128 six-channel 512x512 images (201 MB of u8, larger than the 126 MB L2).
    ncu --set full -k regex:stats_planar_kernel|loader_kernel -s 2 -c 2 ... python tools/prof_hbm_kernels.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recursion_cellular_image_classification_b200 import ops  # noqa: E402
from recursion_cellular_image_classification_b200.synth import synth_planes_torch  # noqa: E402

dev = torch.device("cuda:0")
n, n_exp = 128, 4
src = synth_planes_torch(7, n, dev)
exp = (torch.arange(n, device=dev) % n_exp).to(torch.int32)
acc = None
for _ in range(3):
    acc = ops.stats_accumulate(src, exp, n_exp, acc)
mean, std = ops.stats_finalize(ops.stats_accumulate(src, exp, n_exp))
m, d = ops.normalize_constants(mean.cpu().numpy(), std.cpu().numpy())
norm_m, norm_d = torch.from_numpy(m).to(dev), torch.from_numpy(d).to(dev)
idx = torch.arange(n, dtype=torch.int32, device=dev)
crop = torch.zeros(n, 2, dtype=torch.int32, device=dev)
aug = torch.randint(0, 16, (n,), device=dev, dtype=torch.uint8)
out = torch.empty(n, 256, 256, 32, dtype=torch.bfloat16, device=dev)
for _ in range(3):
    ops.load_norm_aug(src, idx, exp, aug, crop, norm_m, norm_d, (512, 512), ops.OUT_BF16_S2D32, out=out)
torch.cuda.synchronize()
print("ok")
