#!/usr/bin/env python
"""train() itself on N GPUs (round 1 measured the N>1 step only through bench.py's own loop; train()'s multi-rank path
is covered on CPU by gloo tests with a stand-in model).  Launch:

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu/run_train_ranks.py

Builds a small RxRx1-shaped tree (lossless PNG bytes under .jpeg names), trains DenseNet-121 for two epochs at 64x64
with a global batch of 8 and checks, on every rank: the replicas' parameters are bit-identical after training (gradient
all-reduce + replicated SGD), the loss is finite, and rank 0 wrote the checkpoint.  Prints one JSON line from rank 0.

    ... tests/multi_gpu/run_train_ranks.py --size 512 --samples 256 --bs 64 --decode gpu --workers 4
measures train()'s own throughput at the benchmark's image size: real q95 JPEG files through torch's DataLoader
(worker processes), host or device decode, fused loader, native step, phased all-reduce; images/s = all ranks' images
of the second epoch / the slowest rank's wall time for it."""
import json
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _write_jpeg_tree(root, S, n_wells):
    """RxRx1-shaped tree of real q95 JPEG files (what png_to_jpeg.py writes), smooth fluorescence-like content."""
    import cv2
    import pandas as pd
    from recursion_cellular_image_classification_b200.synth import synth_planes
    rows, ctrl = [], []
    exp, plate = "HEPG2-01", 1
    d = os.path.join(root, "train", exp, "Plate%d" % plate)
    os.makedirs(d, exist_ok=True)
    wells = ["B02", "C03"] + ["W%03d" % i for i in range(n_wells)]
    for wi, well in enumerate(wells):
        for site in (1, 2):
            p = synth_planes(100 * (wi % 7) + site, n=1, H=S, W=S)[0]
            for ch in range(6):
                cv2.imwrite(os.path.join(d, "%s_s%d_w%d.jpeg" % (well, site, ch + 1)), cv2.GaussianBlur(p[ch], (0, 0), 1.5),
                            [cv2.IMWRITE_JPEG_QUALITY, 95])
        rec = {"id_code": "%s_%d_%s" % (exp, plate, well), "experiment": exp, "plate": plate, "well": well, "sirna": wi % 1108}
        if well == "B02":
            ctrl.append(dict(rec, well_type="negative_control"))
        elif well == "C03":
            ctrl.append(dict(rec, well_type="positive_control"))
        else:
            rows.append(rec)
    return pd.DataFrame(rows), pd.DataFrame(ctrl), exp


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=64)
    ap.add_argument("--samples", type=int, default=16, help="training samples (global)")
    ap.add_argument("--bs", type=int, default=8, help="global batch")
    ap.add_argument("--decode", default="host", choices=["host", "gpu"])
    ap.add_argument("--workers", type=int, default=0)
    args = ap.parse_args()
    from recursion_cellular_image_classification_b200 import parallel
    from recursion_cellular_image_classification_b200.cell_classifier import dataloader as dl
    from recursion_cellular_image_classification_b200.cell_classifier.models import TwoSitesNN
    from recursion_cellular_image_classification_b200.cell_classifier.train import train
    from test_gpu_shims import _write_tree
    rank, local_rank, world = parallel.init_from_env()
    root = os.path.join(tempfile.gettempdir(), "rxb_train_ranks_%d" % rank)      # every rank writes its own identical tree
    os.makedirs(root, exist_ok=True)
    os.chdir(root)
    S = args.size
    if S == 64 and args.samples == 16:
        df, dfc, _, exp = _write_tree(os.path.join(root, "data"), S=64)
        df = __import__("pandas").concat([df] * 4, ignore_index=True)            # 16 samples: two steps of 8 per epoch
    else:
        df, dfc, exp = _write_jpeg_tree(os.path.join(root, "data"), S, min(args.samples, 32))
        df = __import__("pandas").concat([df] * ((args.samples + len(df) - 1) // len(df)), ignore_index=True)[:args.samples]
    stats = {exp: {"mean": np.full(6, 0.08), "std": np.full(6, 0.06)}}
    kw = dict(verbose=False, decode=args.decode)
    ds_train = dl.ImagesDS(df, dfc, stats, os.path.join(root, "data"), "train", **kw)
    ds_val = dl.ImagesDS(df[:max(args.bs // world, 4)], dfc, stats, os.path.join(root, "data"), "val", **kw)
    model = TwoSitesNN(pretrained=False, nb_classes=1108)                        # lands on this rank's GPU
    opt = torch.optim.SGD(model.parameters(), lr=0.004, momentum=0.9, nesterov=True, weight_decay=3e-5)
    hp = {"bs": args.bs, "nb_epochs": 2, "scheduler": True, "lr": 0.004, "early_stopping": False, "patience": 10,
          "pretrained": False, "crop": S, "tensorboard": False}
    hist = train("ranks", ds_train, ds_val, model, opt, hp, num_workers=args.workers, device="cuda", debug=True)
    flat = model.flat.detach()
    same = True
    if world > 1:
        parts = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(parts, flat)
        same = all(torch.equal(parts[0], p) for p in parts)
    ok = bool(same and all(np.isfinite(h["val_loss"]) for h in hist) and len(hist) == 3 and
              (rank != 0 or os.path.exists("models/best_model_ranks.pth")))
    # throughput of the second epoch (the first one pays plan creation and worker start-up): all ranks' images over
    # the slowest rank's time
    t = torch.tensor([hist[-1]["train_seconds"], float(hist[-1]["train_images"])], dtype=torch.float64, device=flat.device)
    if world > 1:
        tm = t.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        t[0] = tm[0]
    if rank == 0:
        print(json.dumps({"world": world, "replicas_identical": bool(same), "history": hist, "ok": ok,
                          "device": str(flat.device), "image_size": S, "global_batch": args.bs, "decode": args.decode,
                          "workers_per_rank": args.workers,
                          "train_images_per_s_epoch2": float(t[1] / t[0]) if float(t[0]) > 0 else None}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
