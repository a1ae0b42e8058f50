#!/usr/bin/env python
"""train() itself on N GPUs (round 1 measured the N>1 step only through bench.py's own loop; train()'s multi-rank path
is covered on CPU by gloo tests with a stand-in model).  Launch:

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/train_two_ranks.py

Builds a small RxRx1-shaped tree (lossless PNG bytes under .jpeg names), trains DenseNet-121 for two epochs at 64x64
with a global batch of 8 and checks, on every rank: the replicas' parameters are bit-identical after training (gradient
all-reduce + replicated SGD), the loss is finite, and rank 0 wrote the checkpoint.  Prints one JSON line from rank 0.
Written at the end of round 1 without a GPU left to run it on: unmeasured."""
import json
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from recursion_cellular_image_classification_b200 import parallel
    from recursion_cellular_image_classification_b200.cell_classifier import dataloader as dl
    from recursion_cellular_image_classification_b200.cell_classifier.models import TwoSitesNN
    from recursion_cellular_image_classification_b200.cell_classifier.train import train
    from test_gpu_shims import _write_tree
    rank, local_rank, world = parallel.init_from_env()
    root = os.path.join(tempfile.gettempdir(), "rxb_train_ranks_%d" % rank)      # every rank writes its own identical tree
    os.makedirs(root, exist_ok=True)
    os.chdir(root)
    df, dfc, _, exp = _write_tree(os.path.join(root, "data"), S=64)
    df = __import__("pandas").concat([df] * 4, ignore_index=True)                # 16 samples: two steps of 8 per epoch
    stats = {exp: {"mean": np.full(6, 0.08), "std": np.full(6, 0.06)}}
    ds_train = dl.ImagesDS(df, dfc, stats, os.path.join(root, "data"), "train", verbose=False)
    ds_val = dl.ImagesDS(df, dfc, stats, os.path.join(root, "data"), "val", verbose=False)
    model = TwoSitesNN(pretrained=False, nb_classes=1108)                        # lands on this rank's GPU
    opt = torch.optim.SGD(model.parameters(), lr=0.004, momentum=0.9, nesterov=True, weight_decay=3e-5)
    hp = {"bs": 8, "nb_epochs": 2, "scheduler": True, "lr": 0.004, "early_stopping": False, "patience": 10,
          "pretrained": False, "crop": 64}
    hist = train("ranks", ds_train, ds_val, model, opt, hp, num_workers=0, device="cuda", debug=True)
    flat = model.flat.detach()
    same = True
    if world > 1:
        parts = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(parts, flat)
        same = all(torch.equal(parts[0], p) for p in parts)
    ok = bool(same and all(np.isfinite(h["val_loss"]) for h in hist) and len(hist) == 3 and
              (rank != 0 or os.path.exists("models/best_model_ranks.pth")))
    if rank == 0:
        print(json.dumps({"world": world, "replicas_identical": bool(same), "history": hist, "ok": ok,
                          "device": str(flat.device)}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
