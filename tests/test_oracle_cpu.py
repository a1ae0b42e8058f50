"""CPU tests: the oracle against the golden vectors produced by the reference itself, and against
numpy/OpenCV where it restates them."""
import os

import numpy as np
import pytest

from oracle import oracle_np as O
from recursion_cellular_image_classification_b200.synth import synth_logits, synth_plate_groups, synth_planes


def test_stats_oracle_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "stats_golden.npz"))
    for e, seed in enumerate(g["seeds"]):
        planes = synth_planes(int(seed), n=int(g["n_per_exp"]))
        m, s = O.compute_mean_std_arrays(planes)
        np.testing.assert_allclose(m, g["mean"][e], rtol=1e-13)
        np.testing.assert_allclose(s, g["std"][e], rtol=1e-12)
        vm, vs = O.compute_mean_std_arrays(planes, mean=m, std=s)
        np.testing.assert_allclose(vm, g["vmean"][e], atol=1e-12)
        np.testing.assert_allclose(vs, g["vstd"][e], rtol=1e-12)


def test_assign_oracle_matches_reference_golden_64(golden_dir):
    g = np.load(os.path.join(golden_dir, "assign_golden.npz"))
    import torch
    probs = torch.softmax(torch.from_numpy(g["logits64"]), 1).numpy()   # test.py:27
    pr = O.mask_rescale(probs, g["pg64"][:, int(g["et64"])], g["plates64"])
    res = O.greedy_assign(pr)
    np.testing.assert_array_equal(res, g["res64"])


@pytest.mark.timeout(120)
def test_assign_oracle_matches_reference_golden_1108(golden_dir):
    g = np.load(os.path.join(golden_dir, "assign_golden.npz"))
    import torch
    seed, N = int(g["seed1108"]), 1108
    logits = synth_logits(seed, N)
    pg = synth_plate_groups(seed + 1)
    plates = np.random.default_rng(seed + 2).integers(1, 5, size=N)
    probs = torch.softmax(torch.from_numpy(logits), 1).numpy()
    res = O.greedy_assign(O.mask_rescale(probs, pg[:, int(g["et1108"])], plates))
    np.testing.assert_array_equal(res, g["res1108"])


def test_pairwise_sum_restatement_is_numpys():
    """The kernel's row sum follows oracle.pairwise_sum_f32; that must be bit-identical to np.sum(axis=1)."""
    rng = np.random.default_rng(0)
    for n in (1, 5, 8, 9, 64, 127, 128, 129, 277, 1108, 2000):
        a = rng.random((7, n), dtype=np.float32)
        a[0, : n // 2] = 0
        ref = np.sum(a, axis=1)
        mine = np.array([O.pairwise_sum_f32(a[i]) for i in range(a.shape[0])], dtype=np.float32)
        np.testing.assert_array_equal(ref.view(np.uint32), mine.view(np.uint32))


def test_ref_compat_rotation_is_the_integer_gather_of_survey_a2():
    """cv2.warpAffine at multiples of 90 degrees == the gather of SURVEY §A.2 (what the CUDA loader does)."""
    rng = np.random.default_rng(1)
    S = 64
    img = rng.integers(0, 256, size=(S, S, 6), dtype=np.uint8)

    def r(i):
        return i if i <= S - 1 else 2 * (S - 1) - i

    for k in range(4):
        got = O.d4_augment(img, k=k, ref_compat=True)
        exp = np.empty_like(img)
        for y in range(S):
            for x in range(S):
                if k == 0:
                    exp[y, x] = img[y, x]
                elif k == 1:
                    exp[y, x] = img[x, r(S - y)]
                elif k == 2:
                    exp[y, x] = img[r(S - y), r(S - x)]
                else:
                    exp[y, x] = img[r(S - x), y]
        np.testing.assert_array_equal(got, exp)


def test_normalize_matches_f64_formula():
    rng = np.random.default_rng(2)
    img = rng.integers(0, 256, size=(32, 32, 6), dtype=np.uint8)
    mean = rng.random(6) * 0.2 + 0.05
    std = rng.random(6) * 0.1 + 0.05
    got = O.normalize(img, mean, std)
    ref = ((img / 255.0) - mean) / std
    np.testing.assert_allclose(got, ref, rtol=2e-5, atol=2e-5)


def test_s2d_layout():
    x = np.arange(4 * 6 * 8, dtype=np.float32).reshape(4, 6, 8)
    s = O.to_s2d32(x)
    assert s.shape == (2, 3, 32)
    for y in range(4):
        for xx in range(6):
            np.testing.assert_array_equal(s[y // 2, xx // 2, (y & 1) * 16 + (xx & 1) * 8:(y & 1) * 16 + (xx & 1) * 8 + 8],
                                          x[y, xx])
