#!/usr/bin/env python
"""Diagnostic: the same DenseNet-121 training step several times from the same state; per parameter tensor the largest
relative run-to-run difference of the gradient, listed along the depth of the network (the backward pass reaches
features.conv0 last).  Shows how far fp32 accumulation-order noise (atomics, L2 reduce-adds) is amplified by the bf16
backward of a randomly initialised 121-layer network."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recursion_cellular_image_classification_b200.cell_classifier.models import DenseNet121  # noqa: E402

B, S = int(sys.argv[1]) if len(sys.argv) > 1 else 16, int(sys.argv[2]) if len(sys.argv) > 2 else 512
dev = torch.device("cuda:0")
net = DenseNet121(nb_classes=1108, device=dev, seed=6)
g = torch.Generator().manual_seed(7)
x = torch.randn(B, 6, S, S, generator=g).to(torch.bfloat16).float().to(dev)
y = torch.randint(0, 1108, (B,), generator=g).to(dev)
net.train()
buf0 = net.bn_buffers.clone()
runs = []
for _ in range(4):
    net.bn_buffers.copy_(buf0)
    loss = net.train_step(x, y).item()
    runs.append((loss, net.flat.grad.clone()))
print("losses", [r[0] for r in runs])
names = list(net._views)
step = max(1, len(names) // 60)
for i, name in enumerate(names):
    off, k, _ = net._views[name]
    ref = runs[0][1][off:off + k]
    d = max(((r[1][off:off + k] - ref).norm() / ref.norm().clamp_min(1e-30)).item() for r in runs[1:])
    if i % step == 0 or i < 6 or i > len(names) - 6:
        print("%4d %-52s |g| %.3e  max rel run-to-run diff %.3e" % (i, name, ref.norm().item(), d))
