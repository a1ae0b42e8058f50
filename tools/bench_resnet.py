#!/usr/bin/env python
"""The reference's OWN model and training configuration on one B200 (informational, beside bench.py's headline):
TwoSitesNN = ResNet-50 trunk + control thirds + MLP head (reference models.py:7-57), 364x364 crops (dataloader.py:47),
B samples x 3 images per step, native train step + nesterov SGD (csrc/resnet.cu), and evaluation at 512x512 with the
6-image test items (test.py:23-27).  Prints one JSON line; `--library` adds the stock PyTorch path (the same module
structure in torchvision/torch.nn, bf16 autocast, channels_last, cuDNN) on the same box.

    python tools/bench_resnet.py [--batch 32] [--steps 10] [--library]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
FLOP_FWD_364 = 22.93e9          # ResNet-50 trunk, 6x364x364, per image (2*MAC)
FLOP_FWD_512 = 22.93e9 * (512 * 512) / (364 * 364)


def timed(fn, steps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def library_model(nb_classes=1108):
    import torchvision

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.base_nn = torchvision.models.resnet50(weights=None)
            self.base_nn.conv1 = torch.nn.Conv2d(6, 64, kernel_size=7, stride=2, padding=3, bias=False)
            self.base_nn.fc = torch.nn.Identity()
            self.mlp = torch.nn.Sequential(torch.nn.BatchNorm1d(6144), torch.nn.Dropout(0.3), torch.nn.Linear(6144, 1024),
                                           torch.nn.ReLU(), torch.nn.BatchNorm1d(1024), torch.nn.Dropout(0.3),
                                           torch.nn.Linear(1024, nb_classes))

        def forward(self, x):
            bs = x.shape[0]
            f = self.base_nn(x.reshape(-1, *x.shape[2:])).reshape(bs, -1, 2048)
            k = f.shape[1] // 3
            return self.mlp(torch.cat([f[:, :k].mean(1), f[:, k:2 * k].mean(1), f[:, 2 * k:].mean(1)], dim=1))

    return Net()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--library", action="store_true")
    args = ap.parse_args()
    from recursion_cellular_image_classification_b200.cell_classifier.models import TwoSitesResNet50
    dev = torch.device("cuda:0")
    B, G, S = args.batch, 3, 364
    net = TwoSitesResNet50(device=dev, seed=0)
    xs = torch.randn(B * G, S // 2, S // 2, 32, device=dev).to(torch.bfloat16)
    y = torch.randint(0, 1108, (B,), device=dev)
    masks = net.dropout_masks(B)
    net.train()

    def step():
        net.train_step(xs, y, G=G, masks=masks)
        net.sgd_step(B, G, S, S, lr=0.001)

    ms = timed(step, args.steps)
    out = {"model": "TwoSitesNN: ResNet-50 trunk + control thirds + MLP head (reference models.py)",
           "train": {"samples_per_step": B, "images_per_step": B * G, "image_size": S, "ms_per_step": ms,
                     "samples_per_s": B / ms * 1e3, "images_per_s": B * G / ms * 1e3,
                     "trunk_tflops": 3 * FLOP_FWD_364 * B * G / (ms * 1e-3) / 1e12}}
    net.eval()
    Be, Ge, Se = max(args.batch // 4, 4), 6, 512
    xe = torch.randn(Be * Ge, Se // 2, Se // 2, 32, device=dev).to(torch.bfloat16)
    ms_e = timed(lambda: net(xe, G=Ge), args.steps)
    out["eval"] = {"samples_per_call": Be, "images_per_call": Be * Ge, "image_size": Se, "ms_per_call": ms_e,
                   "images_per_s": Be * Ge / ms_e * 1e3, "trunk_tflops": FLOP_FWD_512 * Be * Ge / (ms_e * 1e-3) / 1e12}
    del net
    torch.cuda.empty_cache()
    if args.library:
        torch.backends.cudnn.benchmark = True
        lib = library_model().to(dev).to(memory_format=torch.channels_last).train()
        opt = torch.optim.SGD(lib.parameters(), lr=0.001, momentum=0.9, nesterov=True, weight_decay=3e-5)
        x5 = torch.randn(B, G, 6, S, S, device=dev)
        lossf = torch.nn.CrossEntropyLoss()

        def lib_step():
            opt.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                loss = lossf(lib(x5).float(), y)
            loss.backward()
            opt.step()

        ms_l = timed(lib_step, args.steps)
        out["library_gpu_baseline"] = {"what": "NOT the reference arm: the same module structure in torchvision / torch.nn, "
                                               "bf16 autocast, eager cuDNN/cuBLAS, same box, resident input",
                                       "ms_per_step": ms_l, "images_per_s": B * G / ms_l * 1e3}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
