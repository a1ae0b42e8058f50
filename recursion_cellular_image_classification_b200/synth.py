"""Seeded synthetic RxRx1-shaped inputs shared by tests, fixtures and bench.py."""
import numpy as np


def synth_planes(seed, n, C=6, H=512, W=512):
    """Fluorescence-like u8 planes: clip(Gamma(2, 8*(1+0.25*ch)) * (1+0.1*(seed%7)), 0, 255), [n,C,H,W]."""
    rng = np.random.default_rng(seed)
    out = np.empty((n, C, H, W), dtype=np.uint8)
    scale = 1.0 + 0.1 * (seed % 7)
    for ch in range(C):
        g = rng.gamma(2.0, 8.0 * (1 + 0.25 * ch), size=(n, H, W)) * scale
        out[:, ch] = np.clip(g, 0, 255).astype(np.uint8)
    return out


def synth_logits(seed, N, C=1108):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((N, C)) * 3.0).astype(np.float32)


def synth_plate_groups(seed, C=1108):
    """Four 4x277 partitions like main.py:157-166 builds: column t assigns every class one plate 1..4."""
    rng = np.random.default_rng(seed)
    pg = np.zeros((C, 4), dtype=np.int64)
    for t in range(4):
        pg[:, t] = rng.permutation(np.repeat(np.arange(1, 5), C // 4)).astype(np.int64)
    return pg


def synth_planes_torch(seed, n, device, C=6, H=512, W=512):
    """Same distribution generated on the device (bench-sized corpora); not bit-identical to synth_planes."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(1234 + seed)
    scale = 1.0 + 0.1 * (seed % 7)
    out = torch.empty(n, C, H, W, dtype=torch.uint8, device=device)
    for ch in range(C):
        theta = 8.0 * (1 + 0.25 * ch) * scale
        # Gamma(2, theta) = -theta * (log U1 + log U2)
        u = torch.rand(2, n, H, W, device=device, generator=g).clamp_min_(1e-12).log_().sum(0).mul_(-theta)
        out[:, ch] = u.clamp_(0, 255).to(torch.uint8)
    return out
