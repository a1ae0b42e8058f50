"""GPU parity tests for the arbitrary-angle loader (rxb_load_norm_affine): the reference's full train
transform flips -> ShiftScaleRotate -> RandomCrop -> Normalize (dataloader.py:42-48, 128-139).  Bit-exact against the
oracle restatement, against OpenCV's warpAffine executed in the test, and against tests/golden/warp_golden.npz."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import oracle_np as O
from recursion_cellular_image_classification_b200 import _lib, ops
from recursion_cellular_image_classification_b200.synth import synth_planes

pytestmark = pytest.mark.gpu


def _run(cuda, src, src_idx, exp_id, flips, mats, crops, m, d, out_hw, fmt):
    t = lambda a, dt: torch.as_tensor(np.asarray(a), dtype=dt).to(cuda)
    out = ops.load_norm_affine(t(src, torch.uint8), t(src_idx, torch.int32), t(exp_id, torch.int32),
                               t(flips, torch.uint8), t(mats, torch.float64), t(crops, torch.int32),
                               t(m, torch.float32), t(d, torch.float32), out_hw, fmt)
    torch.cuda.synchronize()
    return out


def _identity_norm(n_exp=1):
    return np.zeros((n_exp, 6), np.float32), np.ones((n_exp, 6), np.float32)


def test_affine_loader_matches_opencv_golden(cuda, golden_dir):
    """u8 gather (identity normalisation) == cv2.warpAffine outputs stored by tests/golden/make_golden.py."""
    g = np.load(os.path.join(golden_dir, "warp_golden.npz"))
    n = len(g["angles"])
    small = np.random.default_rng(int(g["seed_small"])).integers(0, 256, size=(48, 48, 6), dtype=np.uint8)
    src = np.ascontiguousarray(np.moveaxis(small, 2, 0))[None]
    m, d = _identity_norm()
    out = _run(cuda, src, np.zeros(n), np.zeros(n), np.zeros(n), g["mats"], np.zeros((n, 2)), m, d, (48, 48),
               ops.OUT_F32_NCHW).cpu().numpy()
    np.testing.assert_array_equal(np.moveaxis(out, 1, 3).astype(np.uint8), g["out_small"])
    assert np.array_equal(out, np.rint(out))
    # the 512x512 digests: matrices from the host restatement of getRotationMatrix2D
    big = synth_planes(int(g["seed_big"]), n=1)
    mats = np.stack([ops.rotation_matrix(512, 512, float(a)) for a in g["angles"]])
    out = _run(cuda, big, np.zeros(n), np.zeros(n), np.zeros(n), mats, np.zeros((n, 2)), m, d, (512, 512),
               ops.OUT_F32_NCHW).cpu().numpy()
    for i in range(n):
        hwc = np.ascontiguousarray(np.moveaxis(out[i], 0, 2)).astype(np.uint8)
        assert hashlib.sha256(hwc.tobytes()).hexdigest() == str(g["sha256_big"][i]), float(g["angles"][i])


@pytest.mark.parametrize("S,out_hw", [(128, (128, 128)), (512, (364, 364))])
def test_affine_loader_f32_bit_exact_vs_oracle_and_opencv(cuda, S, out_hw):
    rng = np.random.default_rng(11)
    n_src, n_exp, B = 3, 2, 12
    src = rng.integers(0, 256, size=(n_src, 6, S, S), dtype=np.uint8)
    mean, std = rng.random((n_exp, 6)) * 0.2 + 0.05, rng.random((n_exp, 6)) * 0.1 + 0.05
    m, d = ops.normalize_constants(mean, std)
    angles = [0.0, 90.0, 180.0, -90.0, 1e-3] + list(rng.uniform(-180, 180, size=B - 5))
    flips = [i % 4 for i in range(B)]
    crops = [(int(rng.integers(0, S - out_hw[0] + 1)), int(rng.integers(0, S - out_hw[1] + 1))) for _ in range(B)]
    src_idx = [i % n_src for i in range(B)]
    exp_id = [i % n_exp for i in range(B)]
    mats = np.stack([ops.rotation_matrix(S, S, float(a)) for a in angles])
    out = _run(cuda, src, src_idx, exp_id, flips, mats, crops, m, d, out_hw, ops.OUT_F32_NCHW).cpu().numpy()
    for b in range(B):
        kw = dict(vflip=bool(flips[b] & 1), hflip=bool(flips[b] & 2), angle=float(angles[b]), crop_yx=crops[b],
                  out_hw=out_hw)
        ref = O.transform_affine(src[src_idx[b]], mean[exp_id[b]], std[exp_id[b]], **kw)
        np.testing.assert_array_equal(out[b].view(np.uint32), ref.view(np.uint32))
        if b % 3 == 0:   # and against the OpenCV call itself
            cvref = O.transform_affine(src[src_idx[b]], mean[exp_id[b]], std[exp_id[b]], use_cv2=True, **kw)
            np.testing.assert_array_equal(out[b].view(np.uint32), cvref.view(np.uint32))


def test_affine_loader_general_affine_non_square(cuda):
    """Scale/shear/shift matrices on a non-square image whose sides are not multiples of 16."""
    rng = np.random.default_rng(12)
    H, W, B = 72, 100, 8
    src = rng.integers(0, 256, size=(2, 6, H, W), dtype=np.uint8)
    mats = []
    for _ in range(B):
        M = O.rotation_matrix(W, H, float(rng.uniform(-180, 180)), scale=float(rng.uniform(0.5, 1.7)))
        M[:, 2] += rng.uniform(-9, 9, size=2)
        M[1, 0] += rng.uniform(-0.2, 0.2)
        mats.append(M)
    m, d = _identity_norm()
    out = _run(cuda, src, [i % 2 for i in range(B)], np.zeros(B), np.zeros(B), np.stack(mats), np.zeros((B, 2)), m, d,
               (H, W), ops.OUT_F32_NCHW).cpu().numpy()
    for b in range(B):
        ref = O.warp_affine_u8(np.ascontiguousarray(np.moveaxis(src[b % 2], 0, 2)), mats[b])
        np.testing.assert_array_equal(np.moveaxis(out[b], 0, 2).astype(np.uint8), ref)


@pytest.mark.parametrize("fmt", [ops.OUT_BF16_NHWC8, ops.OUT_BF16_S2D32])
def test_affine_loader_bf16_formats(cuda, fmt):
    rng = np.random.default_rng(13)
    S, out_hw, B = 128, (96, 64), 6
    src = rng.integers(0, 256, size=(2, 6, S, S), dtype=np.uint8)
    mean, std = rng.random((1, 6)) * 0.2 + 0.05, rng.random((1, 6)) * 0.1 + 0.05
    m, d = ops.normalize_constants(mean, std)
    angles = rng.uniform(-180, 180, size=B)
    mats = np.stack([ops.rotation_matrix(S, S, float(a)) for a in angles])
    out = _run(cuda, src, [i % 2 for i in range(B)], np.zeros(B), [i % 4 for i in range(B)], mats, [(3, 5)] * B, m, d,
               out_hw, fmt).float().cpu().numpy()
    for b in range(B):
        ref = O.transform_affine(src[b % 2], mean[0], std[0], vflip=bool(b & 1), hflip=bool(b & 2),
                                 angle=float(angles[b]), crop_yx=(3, 5), out_hw=out_hw)
        exp = O.to_nhwc8_bf16(ref)
        if fmt == ops.OUT_BF16_S2D32:
            exp = O.to_s2d32(exp)
        np.testing.assert_array_equal(out[b], exp)


def test_affine_loader_empty_batch_and_errors(cuda):
    src = torch.zeros(1, 6, 32, 32, dtype=torch.uint8, device=cuda)
    m, d = _identity_norm()
    out = _run(cuda, src.cpu().numpy(), np.zeros(0), np.zeros(0), np.zeros(0), np.zeros((0, 2, 3)), np.zeros((0, 2)),
               m, d, (32, 32), ops.OUT_F32_NCHW)
    assert tuple(out.shape) == (0, 6, 32, 32)
    with pytest.raises(_lib.RxbError):      # crop larger than the image
        _run(cuda, src.cpu().numpy(), [0], [0], [0], np.eye(2, 3)[None], [(0, 0)], m, d, (48, 48), ops.OUT_F32_NCHW)
    with pytest.raises(_lib.RxbError):      # S2D32 needs even output sizes
        _run(cuda, src.cpu().numpy(), [0], [0], [0], np.eye(2, 3)[None], [(0, 0)], m, d, (31, 31), ops.OUT_BF16_S2D32)
    with pytest.raises(_lib.RxbError):      # one matrix per image
        _run(cuda, src.cpu().numpy(), [0, 0], [0, 0], [0, 0], np.eye(2, 3)[None], [(0, 0)] * 2, m, d, (32, 32),
             ops.OUT_F32_NCHW)


def test_images_ds_rotate_augmentation_matches_oracle(cuda, tmp_path):
    """ImagesDS(augment='rotate'): the reference's full train transform with explicitly drawn parameters."""
    from test_gpu_shims import _write_tree
    from recursion_cellular_image_classification_b200.cell_classifier import dataloader as dl
    root = str(tmp_path)
    df, dfc, planes, exp = _write_tree(root)
    stats = {exp: {"mean": np.linspace(0.05, 0.1, 6), "std": np.linspace(0.04, 0.08, 6)}}
    ds = dl.ImagesDS(df, dfc, stats, root, "train", verbose=False, crop=40, augment="rotate")
    items = [ds.raw_item(i) for i in (0, 1, 2)]
    batch = dl.collate_raw(items)
    assert tuple(batch["mats"].shape) == (3, 3, 2, 3) and batch["mats"].dtype == torch.float64
    got = ds.device_batch(batch, cuda, out_format=ops.OUT_F32_NCHW).cpu().numpy().reshape(3, 3, 6, 40, 40)
    for i, item in enumerate(items):
        for g_ in range(3):
            c = int(item["codes"][g_])
            img = np.moveaxis(item["planes"][g_].numpy(), 0, 2)
            if c & 1:
                img = img[::-1]
            if c & 2:
                img = img[:, ::-1]
            y0, x0 = (int(v) for v in item["crops"][g_])
            w = O.warp_affine_u8(np.ascontiguousarray(img), item["mats"][g_].numpy())[y0:y0 + 40, x0:x0 + 40]
            ref = np.moveaxis(O.normalize(w, stats[exp]["mean"], stats[exp]["std"]), 2, 0)
            assert np.array_equal(got[i, g_].view(np.uint32), np.ascontiguousarray(ref).view(np.uint32))
    x, label = ds[0]                        # reference-compatible item: float32 [3,6,h,w], int label
    assert x.dtype == torch.float32 and tuple(x.shape) == (3, 6, 40, 40) and isinstance(label, int)
    with pytest.raises(ValueError):
        dl.ImagesDS(df, dfc, stats, root, "train", verbose=False, augment="shear")
