"""Generate the golden fixtures by running the REFERENCE's own functions (imported from /root/reference)
on seeded synthetic inputs.  Run in the build container only (the reference does not travel):

    python tests/golden/make_golden.py

Fixtures written next to this file:
  stats_golden.npz   compute_stats_experiments.compute_mean_std (reference :8-24) on a synthetic tree of
                     lossless single-channel images (PNG bytes under the .jpeg names the reference globs),
                     normal mode and verification mode.
  assign_golden.npz  cell_classifier.test.test (reference test.py:9-58) driven by a seeded logits callable:
                     the masked+rescaled assignment for a 64-well case (inputs stored) and a full
                     1108-well experiment (inputs regenerated from the seed).
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


sys.path.insert(0, os.path.join(HERE, '..', '..'))
from recursion_cellular_image_classification_b200.synth import synth_planes, synth_logits, synth_plate_groups  # noqa: E402


def make_stats():
    import cv2
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp()
    os.chdir(tmp)  # the reference module is a script: it globs data/ and writes a pickle in the CWD
    try:
        os.makedirs("data/train", exist_ok=True)
        os.makedirs("data/test", exist_ok=True)
        sys.path.insert(0, REF)
        import compute_stats_experiments as cse
        seeds = [11, 12]
        means, stds, vmeans, vstds = [], [], [], []
        for e, seed in enumerate(seeds):
            planes = synth_planes(seed, n=4)
            d = os.path.join(tmp, "exp%d" % e, "Plate1")
            os.makedirs(d)
            paths = []
            for i in range(planes.shape[0]):
                for ch in range(6):
                    # <well>_s<site>_w<ch>.jpeg ; PNG bytes (lossless), cv2 sniffs the content
                    p = os.path.join(d, "B%02d_s%d_w%d.jpeg" % (2 + i // 2, 1 + i % 2, ch + 1))
                    ok, buf = cv2.imencode(".png", planes[i, ch])
                    assert ok
                    with open(p, "wb") as f:
                        f.write(buf.tobytes())
                    paths.append(p)
            assert all("_" not in os.path.dirname(p) for p in paths)
            m, s = cse.compute_mean_std(paths)
            vm, vs = cse.compute_mean_std(paths, mean=m, std=s)
            means.append(m); stds.append(s); vmeans.append(vm); vstds.append(vs)
        np.savez(os.path.join(HERE, "stats_golden.npz"), seeds=np.array(seeds), n_per_exp=4,
                 mean=np.array(means), std=np.array(stds), vmean=np.array(vmeans), vstd=np.array(vstds))
        print("stats golden:", np.array(means)[0], np.array(stds)[0])
    finally:
        os.chdir(cwd)


def make_assign():
    import pandas as pd
    import torch
    sys.path.insert(0, REF)
    from cell_classifier.test import test as ref_test

    def run(N, seed, et):
        logits = synth_logits(seed, N)
        pg = synth_plate_groups(seed + 1)
        plates = np.random.default_rng(seed + 2).integers(1, 5, size=N)
        df = pd.DataFrame({"plate": plates})

        class DS(torch.utils.data.Dataset):
            def __len__(self):
                return N

            def __getitem__(self, i):
                return torch.tensor([float(i)]), "id%d" % i

        def model(x):
            idx = x[:, 0].long().numpy()
            return torch.from_numpy(logits[idx])

        res = ref_test(df, DS(), pg, et, model, bs=16, num_workers=0, device="cpu")
        return logits, pg, plates, res

    l64, pg64, pl64, r64 = run(64, 101, 2)
    _, _, _, r1108 = run(1108, 202, 1)
    np.savez_compressed(os.path.join(HERE, "assign_golden.npz"), logits64=l64, pg64=pg64, plates64=pl64, et64=2,
                        res64=r64, seed1108=202, et1108=1, res1108=r1108)
    print("assign golden:", r64[:8], r1108[:8])


if __name__ == "__main__":
    make_stats()
    make_assign()
