"""CPU tests of the reference-facing shims' HOST logic with the device kernels replaced by numpy stand-ins (the
oracle): which files go in which launch, in what order, with what parameters.  The kernels themselves are tested on
the GPU (tests/test_gpu_*.py); nothing here exercises librxb."""
import numpy as np
import pytest
import torch

from oracle import oracle_np as O
from recursion_cellular_image_classification_b200 import compute_stats_experiments as cse
from recursion_cellular_image_classification_b200 import ops
from recursion_cellular_image_classification_b200.synth import synth_planes


@pytest.fixture
def fake_kernels(monkeypatch):
    """numpy stand-ins with the wrappers' signatures and return types."""
    calls = {"decode": 0, "accumulate": 0}

    def jpeg_decode_gray(blob, offsets, hw, select=None, out=None, check_status=True, parallel=True):
        calls["decode"] += 1
        b, o = blob.numpy(), offsets.numpy()
        idx = range(len(o) - 1) if select is None else select.tolist()
        planes = np.stack([O.jpeg_decode_gray(b[o[i]:o[i + 1]].tobytes()) for i in idx])
        assert planes.shape[1:] == tuple(hw)
        return torch.from_numpy(planes)

    def stats_accumulate(imgs, exp_id, n_exp, acc=None):
        calls["accumulate"] += 1
        assert imgs.dtype == torch.uint8 and imgs.dim() == 4 and exp_id.dtype == torch.int32
        if acc is None:
            acc = tuple(torch.zeros(n_exp, imgs.shape[1], dtype=torch.int64) for _ in range(3))
        x = imgs.to(torch.int64)
        for i, e in enumerate(exp_id.tolist()):
            acc[0][e] += x[i].sum(dim=(1, 2))
            acc[1][e] += (x[i] ** 2).sum(dim=(1, 2))
            acc[2][e] += imgs.shape[2] * imgs.shape[3]
        return acc

    def stats_finalize(acc, pre_mean=None, pre_std=None):
        s, q, c = (a.double() for a in acc)
        mean, ex2 = s / c / 255.0, q / c / 255.0 ** 2
        if pre_mean is not None:        # statistics of (x/255 - pm)/ps
            pm, ps = pre_mean, pre_std
            mean, ex2 = (mean - pm) / ps, (ex2 - 2 * pm * (s / c / 255.0) + pm ** 2) / ps ** 2
        return mean, torch.sqrt(ex2 - mean ** 2)

    monkeypatch.setattr(ops, "jpeg_decode_gray", jpeg_decode_gray)
    monkeypatch.setattr(ops, "stats_accumulate", stats_accumulate)
    monkeypatch.setattr(ops, "stats_finalize", stats_finalize)
    return calls


def _write_jpegs(root, planes):
    import cv2
    d = root / "exp0" / "Plate1"
    d.mkdir(parents=True)
    paths = []
    for i in range(planes.shape[0]):
        for ch in range(6):
            p = str(d / ("B%02d_s%d_w%d.jpeg" % (2 + i // 2, 1 + i % 2, ch + 1)))
            cv2.imwrite(p, planes[i, ch], [cv2.IMWRITE_JPEG_QUALITY, 95])
            paths.append(p)
    return paths


@pytest.mark.parametrize("decode", ["host", "gpu"])
def test_compute_mean_std_host_logic(tmp_path, fake_kernels, decode):
    """Chunking, channel parsing from the file name, the accumulator hand-over and the verification mode of
    compute_mean_std, for both decode paths, against the oracle on the decoded pixels."""
    import cv2
    planes = synth_planes(41, n=3, H=48, W=48)
    paths = _write_jpegs(tmp_path, planes)
    rng = np.random.default_rng(0)
    paths = [paths[i] for i in rng.permutation(len(paths))]            # channels arrive in any order
    decoded = np.zeros((3, 6, 48, 48), np.uint8)
    for p in paths:
        name = p.split("/")[-1]
        well, site, ch = int(name[1:3]) - 2, int(name[5]) - 1, int(name[8]) - 1
        decoded[well * 2 + site, ch] = cv2.imread(p, cv2.IMREAD_GRAYSCALE)
    om, os_ = O.compute_mean_std_arrays(decoded)
    m, s = cse.compute_mean_std(paths, device="cpu", chunk=7, decode=decode)
    assert m.dtype == np.float64 and m.shape == (6,) and s.shape == (6,)
    np.testing.assert_allclose(m, om, rtol=1e-12)
    np.testing.assert_allclose(s, os_, rtol=1e-10)
    assert fake_kernels["accumulate"] == 3 and fake_kernels["decode"] == (3 if decode == "gpu" else 0)
    vm, vs = cse.compute_mean_std(paths, mean=m, std=s, device="cpu", decode=decode)
    np.testing.assert_allclose(vm, 0, atol=1e-9)
    np.testing.assert_allclose(vs, 1, rtol=1e-9)
    with pytest.raises(ValueError):
        cse.compute_mean_std(paths, device="cpu", decode="nvjpeg")
    m0, s0 = cse.compute_mean_std([], device="cpu", decode=decode)      # empty path list: 0/0 like the reference
    assert m0.shape == (6,)
