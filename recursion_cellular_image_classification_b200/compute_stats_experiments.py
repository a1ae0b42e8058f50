"""Per-experiment channel statistics — drop-in for the reference's compute_stats_experiments.py.

`compute_mean_std(paths, mean=None, std=None)` keeps the reference signature and return value
(compute_stats_experiments.py:8-24: two float64[6] arrays); the per-pixel reduction runs in
rxb_stats_accumulate on the GPU (exact integer sums) instead of the Python/numpy loop at :13-20.
JPEG decode stays on the host by default (cv2, like the reference :15); `decode='gpu'` moves it to the device too
(rxb_jpeg_decode_gray, bit-identical to cv2 for the baseline grayscale files png_to_jpeg.py writes) — decode is 75 % of
the reference function's time.

Unlike the reference module, importing this file has no side effects; `python -m
recursion_cellular_image_classification_b200.compute_stats_experiments` reproduces the script
(:27-57): glob data/, write stats_experiments.pickle, print the verification pass — under torchrun with the
experiments sharded over the ranks (one GPU each), rank 0 writing the pickle.
"""
import glob
import os
import pickle

import numpy as np
import torch

from . import ops

NB_CHANNELS = 6
FILENAME = "stats_experiments.pickle"


def _channel_of(path):
    # compute_stats_experiments.py:14 parses `<well>_s<site>_w<channel>.jpeg` with path.split('_')[2][1]; the same
    # rule applied to the file name only, so directories containing '_' do not break it.
    return int(os.path.basename(path).split('_')[2][1]) - 1


def accumulate_paths(paths, device="cuda", chunk=384, decode="host"):
    """The exact integer accumulators (sum x, sum x^2, pixel count; int64 [6,1] each, on the device) of the files'
    pixels — everything compute_mean_std needs from the images, for the plain pass AND the verification pass."""
    if decode not in ("host", "gpu"):
        raise ValueError("decode must be 'host' or 'gpu'")
    dev = torch.device(device)
    acc = None
    for i in range(0, len(paths), chunk):
        part = paths[i:i + chunk]
        if decode == "gpu":
            bufs = []
            for p in part:
                with open(p, "rb") as f:
                    bufs.append(f.read())
            blob, offsets = ops.pack_jpeg_buffers(bufs)
            planes = ops.jpeg_decode_gray(blob.to(dev), offsets.to(dev), ops.jpeg_frame_size(bufs[0]))[:, None]
        else:
            import cv2
            ims = [cv2.imread(p, cv2.IMREAD_GRAYSCALE) for p in part]
            for p, im in zip(part, ims):
                if im is None:
                    raise FileNotFoundError(p)
            planes = torch.from_numpy(np.stack(ims)[:, None]).to(dev)             # [n,1,H,W] u8
        slot = torch.tensor([_channel_of(p) for p in part], dtype=torch.int32, device=dev)
        acc = ops.stats_accumulate(planes, slot, NB_CHANNELS, acc)                # one "experiment slot" per channel
    if acc is None:
        acc = tuple(torch.zeros(NB_CHANNELS, 1, dtype=torch.int64, device=dev) for _ in range(3))
    return acc


def mean_std_from_acc(acc, mean=None, std=None):
    """compute_stats_experiments.py:22-23 from the accumulators; with mean/std the statistics of (x/255-mean)/std
    (the reference's verification mode, :16-17) are derived from the SAME sums — no second pass over the files."""
    dev = acc[0].device
    pm = ps = None
    if (mean is not None) and (std is not None):
        pm = torch.as_tensor(np.asarray(mean, dtype=np.float64).reshape(NB_CHANNELS, 1), device=dev)
        ps = torch.as_tensor(np.asarray(std, dtype=np.float64).reshape(NB_CHANNELS, 1), device=dev)
    m, s = ops.stats_finalize(acc, pm, ps)
    return m.cpu().numpy().reshape(NB_CHANNELS), s.cpu().numpy().reshape(NB_CHANNELS)


def compute_mean_std(paths, mean=None, std=None, device="cuda", chunk=384, decode="host"):
    return mean_std_from_acc(accumulate_paths(paths, device=device, chunk=chunk, decode=decode), mean, std)


def stats_from_planes(planes, exp_id, n_exp, acc=None):
    """Device-native entry for decoded corpora: planes u8 [n,6,H,W] (cuda), exp_id int32 [n] ->
    accumulators; finish with `finalize`.  Experiments can be sharded across ranks and the int64
    accumulators all-reduced (parallel.allreduce_stats) — the sums are exact, so any split agrees."""
    return ops.stats_accumulate(planes, exp_id, n_exp, acc)


def finalize(acc, experiments):
    """-> the reference's pickle schema {exp: {'mean': f64[6], 'std': f64[6]}}."""
    m, s = ops.stats_finalize(acc)
    m, s = m.cpu().numpy(), s.cpu().numpy()
    return {e: {"mean": m[i].copy(), "std": s[i].copy()} for i, e in enumerate(experiments)}


def _gather_dicts(local):
    """Merge the per-rank {experiment: stats} dictionaries on every rank (tiny host objects)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(local)
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, local)
    merged = dict()
    for part in parts:
        merged.update(part)
    return merged


def main(decode="host", device=None, verify=True):
    """The reference script (:27-57).  Under torchrun the experiments are independent units sharded over the ranks
    (SURVEY 8e), each rank on its own GPU; rank 0 writes the pickle.  Returns the full dictionary on every rank."""
    from . import parallel
    rank, local_rank, world = parallel.init_from_env()
    if device is None:
        device = "cuda:%d" % local_rank
    experiments_train = [e.split('/')[-2] for e in glob.glob('data/train/*/', recursive=True)]
    experiments_test = [e.split('/')[-2] for e in glob.glob('data/test/*/', recursive=True)]
    experiments = sorted(experiments_train) + sorted(experiments_test)   # the same order on every rank
    begin, end = parallel.shard_range(len(experiments), rank, world)
    local, held = dict(), dict()
    for experiment in experiments[begin:end]:
        paths = glob.glob('data/*/' + experiment + '/*/*.jpeg', recursive=True)
        held[experiment] = accumulate_paths(paths, device=device, decode=decode)   # 18 integers per experiment
        mean, std = mean_std_from_acc(held[experiment])
        local[experiment] = {'mean': mean, 'std': std}
    stats_experiments = _gather_dicts(local)
    stats_experiments = {e: stats_experiments[e] for e in experiments}      # the reference's key order
    if rank == 0:
        with open(FILENAME, 'wb') as f:
            pickle.dump(stats_experiments, f)
    if verify:
        # the reference re-reads and re-decodes every file here (:51-57); the sums held from the first pass give the
        # same numbers (rxb_stats_finalize's pre_mean / pre_std mode)
        check = dict()
        for experiment in experiments[begin:end]:
            check[experiment] = mean_std_from_acc(held[experiment], mean=stats_experiments[experiment]['mean'],
                                                  std=stats_experiments[experiment]['std'])
        check = _gather_dicts(check)
        if rank == 0:
            print()
            print('Verification:')
            for experiment in experiments:
                print('mean=', check[experiment][0])
                print('std=', check[experiment][1])
    return stats_experiments


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--decode", default="host", choices=["host", "gpu"], help="where the JPEG files are decoded")
    ap.add_argument("--no-verify", action="store_true", help="skip the reference's verification pass")
    a = ap.parse_args()
    main(decode=a.decode, verify=not a.no_verify)
