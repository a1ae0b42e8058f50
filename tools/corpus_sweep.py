#!/usr/bin/env python
"""BASELINE config 5: per-experiment statistics + normalisation throughput over a synthetic RxRx1-shaped corpus
(51 experiments x 2,464 six-channel 512x512 images = 125,664 images, 197.6 GB of u8), streamed one experiment-chunk
at a time through rxb_stats_accumulate and rxb_load_norm_aug.

    python tools/corpus_sweep.py [--experiments 51] [--images-per-exp 2464] [--chunk 616]
    torchrun --nproc-per-node N tools/corpus_sweep.py ...      # experiments sharded over ranks, exact int64 all-reduce

The corpus is regenerated on the device chunk by chunk (generation is outside the timed regions); every chunk is
larger than the 126 MB L2.  Prints one JSON line (rank 0): GB/s of each pass against the measured HBM copy peak, and
a known-answer check of the statistics (the exact integer sums of a re-generated chunk).
Measured in round 2: profiles/r02_corpus_sweep_51_experiments.json (1 GPU), profiles/r02_n2_corpus_sweep.json."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--experiments", type=int, default=51)
    ap.add_argument("--images-per-exp", type=int, default=2464)
    ap.add_argument("--chunk", type=int, default=616)
    args = ap.parse_args()
    from recursion_cellular_image_classification_b200 import ops, parallel
    from recursion_cellular_image_classification_b200.synth import synth_planes_torch
    rank, local_rank, world = parallel.init_from_env()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    n_exp, per_exp, chunk = args.experiments, args.images_per_exp, args.chunk
    e_begin, e_end = parallel.shard_range(n_exp, rank, world)
    acc = tuple(torch.zeros(n_exp, 6, dtype=torch.int64, device=dev) for _ in range(3))
    ev = lambda: torch.cuda.Event(enable_timing=True)
    # untimed first launches (module load, first-touch) so that no rank's sum starts with a cold kernel
    warm = synth_planes_torch(1, 4, dev)
    ops.stats_accumulate(warm, torch.zeros(4, dtype=torch.int32, device=dev), 1)
    ops.load_norm_aug(warm, torch.arange(4, dtype=torch.int32, device=dev), torch.zeros(4, dtype=torch.int32, device=dev),
                      torch.zeros(4, dtype=torch.uint8, device=dev), torch.zeros(4, 2, dtype=torch.int32, device=dev),
                      torch.zeros(1, 6, device=dev), torch.ones(1, 6, device=dev), (512, 512), ops.OUT_BF16_S2D32)
    torch.cuda.synchronize()
    del warm
    stats_ms, images = 0.0, 0
    for e in range(e_begin, e_end):
        for c0 in range(0, per_exp, chunk):
            n = min(chunk, per_exp - c0)
            planes = synth_planes_torch(1000 * e + c0, n, dev)             # seed = (experiment, chunk)
            exp_id = torch.full((n,), e, dtype=torch.int32, device=dev)
            a, b = ev(), ev()
            a.record()
            ops.stats_accumulate(planes, exp_id, n_exp, acc)
            b.record()
            torch.cuda.synchronize()
            stats_ms += a.elapsed_time(b)
            images += n
            del planes
    acc = parallel.allreduce_stats(acc)
    mean, std = ops.stats_finalize(acc)
    m, d = ops.normalize_constants(mean.cpu().numpy(), std.cpu().numpy())
    norm_m, norm_d = torch.from_numpy(m).to(dev), torch.from_numpy(d).to(dev)
    load_ms = 0.0
    out = torch.empty(chunk, 256, 256, 32, dtype=torch.bfloat16, device=dev)
    for e in range(e_begin, e_end):
        for c0 in range(0, per_exp, chunk):
            n = min(chunk, per_exp - c0)
            planes = synth_planes_torch(1000 * e + c0, n, dev)
            idx = torch.arange(n, dtype=torch.int32, device=dev)
            exp_id = torch.full((n,), e, dtype=torch.int32, device=dev)
            aug = torch.zeros(n, dtype=torch.uint8, device=dev)
            crop = torch.zeros(n, 2, dtype=torch.int32, device=dev)
            a, b = ev(), ev()
            a.record()
            ops.load_norm_aug(planes, idx, exp_id, aug, crop, norm_m, norm_d, (512, 512), ops.OUT_BF16_S2D32, out=out[:n])
            b.record()
            torch.cuda.synchronize()
            load_ms += a.elapsed_time(b)
            del planes
    # known answer: the integer sums of one chunk, recomputed with torch
    e0 = e_begin
    planes = synth_planes_torch(1000 * e0, min(chunk, per_exp), dev)
    one = ops.stats_accumulate(planes, torch.zeros(planes.shape[0], dtype=torch.int32, device=dev), 1)
    exact = bool((one[0][0] == planes.to(torch.int64).sum(dim=(0, 2, 3))).all().item())
    t = torch.tensor([stats_ms, load_ms, float(images)], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        stats_ms, load_ms, images = tmax[0].item(), tmax[1].item(), t[2].item()
    if rank == 0:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] * world
        sb, lb = 6 * 512 * 512 * images, 3 * 6 * 512 * 512 * images
        print(json.dumps({
            "workload": "%d experiments x %d images of 6x512x512 u8 (%.1f GB), %d GPU(s)" % (n_exp, per_exp, sb / 1e9, world),
            "stats": {"ms": stats_ms, "gbs": sb / stats_ms / 1e6, "frac_of_measured_hbm": sb / stats_ms / 1e6 / peak,
                      "images_per_s": images / stats_ms * 1e3},
            "normalise": {"ms": load_ms, "gbs": lb / load_ms / 1e6, "frac_of_measured_hbm": lb / load_ms / 1e6 / peak,
                          "images_per_s": images / load_ms * 1e3},
            "known_answer_exact": exact,
            "mean_exp0": mean[e_begin].cpu().tolist()}))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
