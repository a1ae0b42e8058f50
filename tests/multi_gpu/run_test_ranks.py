#!/usr/bin/env python
"""test() itself on N GPUs (SURVEY 8e test/TTA row): wells sharded over the ranks, logits all-gathered over NCCL, the
same assignment on every rank.  Runs the seeded config-4 case of tests/c4_case.py (expected classes known from the
fp32 oracle) with 1 and 8 D4 views and prints one JSON line from rank 0.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu/run_test_ranks.py"""
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import c4_case as C
    from recursion_cellular_image_classification_b200 import parallel
    from recursion_cellular_image_classification_b200.cell_classifier import dataloader as dl
    from recursion_cellular_image_classification_b200.cell_classifier.models import DenseNet121
    from recursion_cellular_image_classification_b200.cell_classifier.test import test as rxb_test
    from recursion_cellular_image_classification_b200.synth import synth_plate_groups
    rank, local_rank, world = parallel.init_from_env()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    root = os.path.join(tempfile.gettempdir(), "rxb_test_ranks_%d" % rank)
    os.makedirs(root, exist_ok=True)
    pg = synth_plate_groups(3)
    df, dfc, planes = C.write_tree(root)
    ref, classes = C.build_oracle_model(planes, pg)               # CPU, identical on every rank (seeded)
    ds = dl.ImagesDS(df, dfc, {C.EXP: {"mean": C.MEAN, "std": C.STD}}, root, "test", verbose=False, device=str(dev))
    net = DenseNet121(nb_classes=1108, device=dev)
    net.load_state_dict(ref.state_dict())
    net.eval()
    out = {"world": world, "expected": classes}
    ok = True
    for views in (1, 8):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = rxb_test(df, ds, pg, C.EXPERIMENT_TYPE, net, bs=2, num_workers=0, device=str(dev), tta_views=views)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        mine = torch.tensor(res, device=dev)
        same = True
        if world > 1:
            parts = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine)
            same = all(torch.equal(parts[0], p) for p in parts)
        good = bool(same and list(res.astype(int)) == classes)
        ok = ok and good
        out["views_%d" % views] = {"assignment": res.astype(int).tolist(), "same_on_all_ranks": bool(same),
                                   "matches_fp32_oracle": good, "seconds": dt}
    if rank == 0:
        out["ok"] = ok
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
