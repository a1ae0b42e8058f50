#!/usr/bin/env python
"""The statistics SCRIPT itself on N GPUs (per-experiment statistics row): writes a small RxRx1-shaped data/
tree (lossless PNG bytes under the reference's .jpeg names), runs compute_stats_experiments.main() — experiments sharded
over the ranks, one GPU each, dictionaries merged, rank 0 writes stats_experiments.pickle and prints the verification
pass — and checks the pickle against a plain numpy float64 computation of compute_stats_experiments.py:13-23.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu/run_stats_ranks.py"""
import json
import os
import pickle
import sys
import tempfile

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import cv2
    from recursion_cellular_image_classification_b200 import compute_stats_experiments as cse
    from recursion_cellular_image_classification_b200 import parallel
    from recursion_cellular_image_classification_b200.synth import synth_planes
    rank, local_rank, world = parallel.init_from_env()
    root = os.path.join(tempfile.gettempdir(), "rxb_stats_ranks_%d" % rank)       # every rank writes the same tree
    os.makedirs(root, exist_ok=True)
    os.chdir(root)
    truth = {}
    for split, exps in (("train", ("HEPG2-01", "RPE-03", "U2OS-02")), ("test", ("HUVEC-17", "HEPG2-08"))):
        for ei, exp in enumerate(exps):
            d = os.path.join(root, "data", split, exp, "Plate1")
            os.makedirs(d, exist_ok=True)
            planes = synth_planes(len(exp) * 10 + ei, n=8, H=64, W=64)            # 4 wells x 2 sites
            for i in range(8):
                for ch in range(6):
                    with open(os.path.join(d, "W%02d_s%d_w%d.jpeg" % (i // 2, 1 + i % 2, ch + 1)), "wb") as f:
                        f.write(cv2.imencode(".png", planes[i, ch])[1].tobytes())
            x = planes.astype(np.float64) / 255.0                                   # compute_stats_experiments.py:15-23
            mean = x.mean(axis=(0, 2, 3))
            truth[exp] = (mean, np.sqrt((x ** 2).mean(axis=(0, 2, 3)) - mean ** 2))
    stats = cse.main(verify=True)
    ok = list(stats.keys()) == sorted(["HEPG2-01", "RPE-03", "U2OS-02"]) + sorted(["HUVEC-17", "HEPG2-08"])
    worst = 0.0
    for exp, (m, s) in truth.items():
        worst = max(worst, float(np.max(np.abs(stats[exp]["mean"] - m) / m)), float(np.max(np.abs(stats[exp]["std"] - s) / s)))
    ok = ok and worst < 1e-9
    if rank == 0:
        on_disk = pickle.load(open("stats_experiments.pickle", "rb"))
        ok = ok and all(np.array_equal(on_disk[e]["mean"], stats[e]["mean"]) for e in stats)
        print(json.dumps({"world": world, "experiments": len(stats), "max_rel_error_vs_numpy_f64": worst, "ok": bool(ok)}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
