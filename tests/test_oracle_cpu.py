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


# ---------------------------------------------------------------- arbitrary-angle ShiftScaleRotate
def test_rotation_matrix_is_opencvs_bitwise():
    import cv2
    rng = np.random.default_rng(4)
    for ang in [0.0, 90.0, -90.0, 180.0, -180.0, 45.0] + list(rng.uniform(-180, 180, size=2000)):
        for (w, h) in ((512, 512), (96, 64)):
            np.testing.assert_array_equal(O.rotation_matrix(w, h, float(ang)),
                                          cv2.getRotationMatrix2D((w / 2, h / 2), float(ang), 1.0))


def test_warp_affine_restatement_matches_golden(golden_dir):
    import hashlib
    g = np.load(os.path.join(golden_dir, "warp_golden.npz"))
    small = np.random.default_rng(int(g["seed_small"])).integers(0, 256, size=(48, 48, 6), dtype=np.uint8)
    big = np.ascontiguousarray(np.moveaxis(synth_planes(int(g["seed_big"]), n=1)[0], 0, 2))
    for i, ang in enumerate(g["angles"]):
        np.testing.assert_array_equal(O.rotation_matrix(48, 48, float(ang)), g["mats"][i])
        np.testing.assert_array_equal(O.warp_affine_u8(small, g["mats"][i]), g["out_small"][i])
        got = O.shift_scale_rotate(big, float(ang))
        assert hashlib.sha256(got.tobytes()).hexdigest() == str(g["sha256_big"][i])


@pytest.mark.parametrize("shape", [(64, 64, 6), (40, 72, 6), (33, 17, 6), (96, 96), (1, 9, 6)])
def test_warp_affine_restatement_is_cv2_warpaffine(shape):
    """Bit-exact against the OpenCV call albumentations makes, for random angles and general affine maps
    (scale, shear, shift), square / non-square / single-channel / degenerate images."""
    import cv2
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, size=shape, dtype=np.uint8)
    h, w = shape[:2]
    mats = [O.rotation_matrix(w, h, float(a)) for a in rng.uniform(-180, 180, size=40)]
    for _ in range(20):
        M = O.rotation_matrix(w, h, float(rng.uniform(-180, 180)), scale=float(rng.uniform(0.6, 1.6)))
        M[:, 2] += rng.uniform(-7, 7, size=2)
        M[0, 1] += rng.uniform(-0.2, 0.2)
        mats.append(M)
    for M in mats:
        ref = cv2.warpAffine(img, M, (w, h), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)
        np.testing.assert_array_equal(O.warp_affine_u8(img, M), ref)


def test_warp_affine_at_right_angles_is_the_ref_compat_gather():
    """At multiples of 90 degrees the fixed-point warp degenerates to the integer gather of SURVEY §A.2."""
    rng = np.random.default_rng(6)
    img = rng.integers(0, 256, size=(64, 64, 6), dtype=np.uint8)
    for k in range(4):
        np.testing.assert_array_equal(O.shift_scale_rotate(img, 90.0 * k), O.d4_augment(img, k=k, ref_compat=True))


def test_transform_affine_restatement_vs_cv2_pipeline():
    rng = np.random.default_rng(8)
    img = rng.integers(0, 256, size=(6, 128, 128), dtype=np.uint8)
    mean, std = rng.random(6) * 0.2 + 0.05, rng.random(6) * 0.1 + 0.05
    for v, h, ang in [(0, 0, 17.0), (1, 0, -133.7), (0, 1, 90.0), (1, 1, 179.2)]:
        a = O.transform_affine(img, mean, std, bool(v), bool(h), ang, (5, 9), (92, 92))
        b = O.transform_affine(img, mean, std, bool(v), bool(h), ang, (5, 9), (92, 92), use_cv2=True)
        np.testing.assert_array_equal(a.view(np.uint32), b.view(np.uint32))


def test_kernel_warp_arithmetic_compiled_for_the_host_is_cv2_warpaffine(tmp_path):
    """The header the CUDA loader is built from (csrc/warp_fixed.cuh), compiled for the host with the kernel's
    per-pixel control flow (tests/warp_host.cpp), against cv2.warpAffine: flips, crops, rotations, general maps."""
    import ctypes
    import subprocess
    import cv2
    here = os.path.dirname(os.path.abspath(__file__))
    inc = os.path.join(here, "..", "recursion_cellular_image_classification_b200", "csrc")
    so = str(tmp_path / "libwarp_host.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c++", "-I", inc,
                    os.path.join(here, "warp_host.cpp"), "-o", so], check=True)
    lib = ctypes.CDLL(so)
    rng = np.random.default_rng(14)
    for (H, W) in ((64, 64), (72, 100), (512, 512)):
        src = rng.integers(0, 256, size=(6, H, W), dtype=np.uint8)
        for t in range(6 if H == 512 else 40):
            M = O.rotation_matrix(W, H, float(rng.uniform(-180, 180)) if t > 3 else 90.0 * t,
                                  scale=1.0 if t % 2 == 0 else float(rng.uniform(0.6, 1.5)))
            if t % 3 == 2:
                M[:, 2] += rng.uniform(-6, 6, size=2)
            vflip, hflip = int(rng.integers(2)), int(rng.integers(2))
            Ho, Wo = int(rng.integers(1, H + 1)), int(rng.integers(1, W + 1))
            y0, x0 = int(rng.integers(0, H - Ho + 1)), int(rng.integers(0, W - Wo + 1))
            dst = np.empty((6, Ho, Wo), dtype=np.uint8)
            lib.warp_host_planar_u8(src.ctypes.data_as(ctypes.c_void_p), H, W,
                                    np.ascontiguousarray(M).ctypes.data_as(ctypes.c_void_p), vflip, hflip, y0, x0, Ho,
                                    Wo, dst.ctypes.data_as(ctypes.c_void_p))
            img = np.moveaxis(src, 0, 2)
            if vflip:
                img = img[::-1]
            if hflip:
                img = img[:, ::-1]
            ref = cv2.warpAffine(np.ascontiguousarray(img), M, (W, H), flags=cv2.INTER_LINEAR,
                                 borderMode=cv2.BORDER_REFLECT_101)[y0:y0 + Ho, x0:x0 + Wo]
            np.testing.assert_array_equal(np.moveaxis(dst, 0, 2), ref)


# ---------------------------------------------------------------- baseline JPEG decode
def _jpeg_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "jpeg_golden.npz"))
    fo = np.concatenate([[0], np.cumsum(g["file_sizes"])])
    po = np.concatenate([[0], np.cumsum([h * w for h, w in g["shapes"]])])
    return [(g["files"][fo[i]:fo[i + 1]].tobytes(), g["planes"][po[i]:po[i + 1]].reshape(tuple(g["shapes"][i])))
            for i in range(len(g["shapes"]))]


def _jpeg_cases():
    """(bytes, cv2.imdecode output): sizes that are / are not multiples of 8, three qualities, default and optimised
    Huffman tables, restart intervals, fluorescence-like / noise / smooth content."""
    import cv2
    rng = np.random.default_rng(15)
    for (H, W) in [(64, 64), (8, 8), (37, 53), (1, 1), (96, 40)]:
        yy, xx = np.mgrid[0:H, 0:W]
        for img in (synth_planes(3, 1, C=1, H=H, W=W)[0, 0], rng.integers(0, 256, size=(H, W), dtype=np.uint8),
                    ((np.sin(yy / 7.0) + np.cos(xx / 5.0)) * 60 + 128).clip(0, 255).astype(np.uint8)):
            for q in (95, 40, 100):
                for extra in ([], [cv2.IMWRITE_JPEG_OPTIMIZE, 1], [cv2.IMWRITE_JPEG_RST_INTERVAL, 3]):
                    ok, buf = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, q] + extra)
                    assert ok
                    yield buf.tobytes(), cv2.imdecode(buf, -1)


def test_jpeg_oracle_matches_reference_golden(golden_dir):
    for buf, plane in _jpeg_golden(golden_dir):
        np.testing.assert_array_equal(O.jpeg_decode_gray(buf), plane)


def test_jpeg_oracle_is_cv2_imdecode():
    n = 0
    for buf, ref in _jpeg_cases():
        np.testing.assert_array_equal(O.jpeg_decode_gray(buf), ref)
        n += 1
    assert n == 135
    import cv2
    ok, buf = cv2.imencode(".jpg", np.zeros((16, 16), np.uint8), [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    with pytest.raises(O.JpegError):
        O.jpeg_decode_gray(buf.tobytes())
    with pytest.raises(O.JpegError):
        O.jpeg_decode_gray(b"not a jpeg at all")


def test_kernel_jpeg_code_compiled_for_the_host_is_cv2_imdecode(tmp_path, golden_dir):
    """The header the CUDA decoder is built from (csrc/jpeg_fixed.cuh), compiled for the host (tests/jpeg_host.cpp):
    bit-exact against the reference golden, cv2.imdecode at 512x512, and the status codes."""
    import ctypes
    import subprocess
    import cv2
    here = os.path.dirname(os.path.abspath(__file__))
    inc = os.path.join(here, "..", "recursion_cellular_image_classification_b200", "csrc")
    so = str(tmp_path / "libjpeg_host.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", "-I", inc, os.path.join(here, "jpeg_host.cpp"),
                    "-o", so], check=True)
    lib = ctypes.CDLL(so)

    def dec(buf, H, W):
        a = np.frombuffer(buf, dtype=np.uint8)
        dst = np.zeros((H, W), np.uint8)
        st = lib.jpeg_host_decode_gray(a.ctypes.data_as(ctypes.c_void_p), len(a), H, W,
                                       dst.ctypes.data_as(ctypes.c_void_p))
        # the speculative parallel path (warp emulated lane by lane) must agree wherever it applies
        dst2 = np.zeros((H, W), np.uint8)
        rounds = ctypes.c_int(0)
        st2 = lib.jpeg_host_decode_gray_parallel(a.ctypes.data_as(ctypes.c_void_p), len(a), H, W,
                                                 dst2.ctypes.data_as(ctypes.c_void_p), ctypes.byref(rounds))
        if st2 == -1:                                                # restart intervals: sequential path only
            assert b"\xff\xdd" in buf
        else:
            assert st2 == st and (st != 0 or np.array_equal(dst, dst2))
        return st, dst

    for buf, plane in _jpeg_golden(golden_dir):
        st, got = dec(buf, *plane.shape)
        assert st == 0
        np.testing.assert_array_equal(got, plane)
    for buf, ref in _jpeg_cases():
        st, got = dec(buf, *ref.shape)
        assert st == 0
        np.testing.assert_array_equal(got, ref)
    for seed in (3, 4):
        ok, buf = cv2.imencode(".jpg", synth_planes(seed, 1, C=1)[0, 0], [cv2.IMWRITE_JPEG_QUALITY, 95])
        st, got = dec(buf.tobytes(), 512, 512)
        assert st == 0
        np.testing.assert_array_equal(got, cv2.imdecode(buf, -1))
    assert dec(buf.tobytes(), 256, 256)[0] == 4                      # frame size != expected
    assert dec(b"not a jpeg at all....", 8, 8)[0] == 1
    assert dec(buf.tobytes()[:300], 512, 512)[0] == 1                # headers cut short
    assert dec(buf.tobytes()[:len(buf) // 2], 512, 512)[0] == 5      # scan cut short (both decode paths agree)
    assert dec(buf.tobytes()[:-4], 512, 512)[0] == 5                 # ... inside the last block
    ok, pbuf = cv2.imencode(".jpg", np.zeros((16, 16), np.uint8), [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    assert dec(pbuf.tobytes(), 16, 16)[0] == 2
    ok, cbuf = cv2.imencode(".jpg", np.zeros((16, 16, 3), np.uint8))
    assert dec(cbuf.tobytes(), 16, 16)[0] == 2                       # three components


# ---------------------------------------------------------------- model glue (reference models.py, M1/M2)
def test_stem_recipe_and_site_averaging_match_reference_golden(golden_dir):
    """The reference's own TwoSitesNN (tests/golden/make_golden.py): its 6-channel stem is the channel-mean of the
    3-channel kernel replicated six times (models.py:17-27) — the recipe the oracle nets and DenseNet121.
    reset_parameters use — and its forward averages sites per image / negative / positive third (:41-53)."""
    import torch
    g = np.load(os.path.join(golden_dir, "model_golden.npz"))
    rgb = torch.from_numpy(g["rgb_sample"])
    stem = torch.stack([torch.mean(rgb, 1)] * 6, dim=1).numpy()              # the oracle's recipe
    np.testing.assert_array_equal(stem, g["stem6_sample"])
    net = O.densenet121_6ch(16, seed=3)
    w = net.features.conv0.weight.detach().numpy()
    assert w.shape == (64, 6, 7, 7)
    for c in range(1, 6):
        np.testing.assert_array_equal(w[:, 0], w[:, c])
    for G in (3, 6):
        x = torch.from_numpy(g["x%d" % G])
        feats = (x.reshape(-1, 6, 8, 8).mean(dim=(2, 3)) * torch.arange(1, 7, dtype=x.dtype))   # the golden's stub trunk
        np.testing.assert_array_equal(O.two_sites_features(feats, x.shape[0]).numpy(), g["y%d" % G])
        # the product's grouping rule (models.sample_group): the sample's own sites are the FIRST third of the item,
        # and their feature mean is the first block of what the reference's forward concatenates (models.py:46-50)
        from recursion_cellular_image_classification_b200.cell_classifier.models import sample_group
        own = sample_group(G)
        assert own == G // 3
        F = feats.shape[1]
        np.testing.assert_array_equal(feats.reshape(x.shape[0], G, F)[:, :own].mean(1).numpy(), g["y%d" % G][:, :F])
    assert [sample_group(G) for G in (1, 2, 3, 6, 4)] == [1, 2, 1, 2, 4]     # items without controls: every image is a site


# ---------------------------------------------------------------- product host helpers (no GPU needed)
def test_product_host_helpers_agree_with_oracle_and_opencv():
    import cv2
    from recursion_cellular_image_classification_b200 import ops
    rng = np.random.default_rng(17)
    for ang in rng.uniform(-180, 180, size=200):
        np.testing.assert_array_equal(ops.rotation_matrix(512, 512, float(ang)),
                                      cv2.getRotationMatrix2D((256.0, 256.0), float(ang), 1.0))
    mean, std = rng.random((3, 6)) * 0.2 + 0.05, rng.random((3, 6)) * 0.1 + 0.05
    m, d = ops.normalize_constants(mean, std)
    om, od = O.normalize_constants(mean, std)
    np.testing.assert_array_equal(m, om)
    np.testing.assert_array_equal(d, od)
    imgs = [rng.integers(0, 256, size=hw, dtype=np.uint8) for hw in ((40, 72), (8, 8), (512, 512))]
    bufs = [cv2.imencode(".jpg", im, [cv2.IMWRITE_JPEG_QUALITY, 95])[1].tobytes() for im in imgs]
    for im, b in zip(imgs, bufs):
        assert ops.jpeg_frame_size(b) == im.shape
    blob, offsets = ops.pack_jpeg_buffers(bufs)
    assert offsets.tolist() == [0] + list(np.cumsum([len(b) for b in bufs]))
    for i, b in enumerate(bufs):
        assert blob[offsets[i]:offsets[i + 1]].numpy().tobytes() == b
    blob0, offsets0 = ops.pack_jpeg_buffers([])
    assert blob0.numel() == 0 and offsets0.tolist() == [0]
    assert [ops.aug_code(v, h, k, r) for v, h, k, r in ((1, 0, 0, 0), (0, 1, 0, 0), (0, 0, 3, 0), (1, 1, 2, 1))] == \
        [1, 2, 12, 27]
    from recursion_cellular_image_classification_b200 import _lib
    with pytest.raises(_lib.RxbError):
        ops.jpeg_frame_size(b"\xff\xd8 nothing useful here")


def test_two_sites_resnet50_oracle_matches_reference_golden(golden_dir):
    """The oracle's restatement of the reference's real model (ResNet-50 trunk + MLP head, SURVEY §8f-3, the next
    model row) gives the reference's own eval-mode logits for the same seed and input."""
    import torch
    g = np.load(os.path.join(golden_dir, "model_golden.npz"))
    net = O.two_sites_resnet50(seed=int(g["full_seed"]))
    net.eval()
    x = torch.randn(2, 3, 6, 64, 64, generator=torch.Generator().manual_seed(6))
    with torch.no_grad():
        y = net(x).numpy()
    np.testing.assert_allclose(y, g["full_logits"], rtol=1e-5, atol=1e-6)


# ---------------------------------------------------------------- memory safety of the kernels' shared headers
@pytest.mark.timeout(300)
def test_sanitizer_fuzz_of_the_kernel_headers():
    """tools/fuzz_jpeg.sh: mutated / truncated / marker-injected JPEG files through both decode flows, and degenerate
    and non-finite matrices through the warp arithmetic, under AddressSanitizer + UBSan with exact-size buffers.  A
    short run here; round 1 ran 11.8 M JPEG cases (after fixing the three defects the fuzzer found: an over-read of
    up to 124 bytes behind the last file by idle lanes of the unstuffing step, a look-up-table overflow on an
    over-subscribed Huffman table, a negative shift on a DC category >= 16)."""
    import shutil
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    probe = subprocess.run("echo 'int main(){return 0;}' | g++ -x c++ -fsanitize=address,undefined - -o /dev/null",
                           shell=True, capture_output=True)
    if probe.returncode != 0 or shutil.which("bash") is None:
        pytest.skip("sanitizer runtime not available")
    p = subprocess.run(["bash", os.path.join(here, "..", "tools", "fuzz_jpeg.sh"), "6", "2"], capture_output=True,
                       text=True)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    assert "warp cases" in p.stdout and p.stdout.count("statuses") == 2


@pytest.mark.skipif(not os.path.isdir("/root/reference/cell_classifier"), reason="the reference tree is only in the build container")
def test_assignment_oracle_vs_the_reference_run_live_on_edge_cases():
    """cell_classifier.test.test of the reference itself (imported from /root/reference — build container only) against
    the oracle on seeded edge cases: wells on a plate no class belongs to (all-zero rows: the loop then writes
    results[0] = 0, test.py:50-53), identical wells (exact ties), more wells on a plate than it has classes, one well."""
    import sys
    import pandas as pd
    import torch
    sys.path.insert(0, "/root/reference")
    try:
        from cell_classifier.test import test as ref_test
    finally:
        sys.path.remove("/root/reference")
    rng = np.random.default_rng(21)
    for case in range(24):
        N = int(rng.choice([1, 2, 7, 19, 40]))
        logits = (rng.standard_normal((N, 1108)) * rng.choice([0.5, 3.0, 9.0])).astype(np.float32)
        pg = synth_plate_groups(100 + case)
        plates = rng.integers(1, 5, size=N)
        kind = case % 4
        if kind == 1 and N > 2:
            plates[rng.integers(N)] = 7                                   # a plate without classes
            logits[1] = logits[0]                                         # two identical wells
            plates[1] = plates[0]
        elif kind == 2:
            pg[:, :] = 2                                                  # every class on plate 2 ...
            pg[:3, :] = 1                                                 # ... except three: plate 1 has 3 classes
            plates[:] = 1                                                 # and N wells
        elif kind == 3:
            logits[:] = 0.0                                               # uniform probabilities everywhere
        et = int(rng.integers(4))

        class DS(torch.utils.data.Dataset):
            def __len__(self):
                return N

            def __getitem__(self, i):
                return torch.tensor([float(i)]), "id%d" % i

        ref = ref_test(pd.DataFrame({"plate": plates}), DS(), pg, et,
                       lambda x: torch.from_numpy(logits[x[:, 0].long().numpy()]), bs=16, num_workers=0, device="cpu")
        probs = torch.softmax(torch.from_numpy(logits), 1).numpy()
        mine = O.greedy_assign(O.mask_rescale(probs, pg[:, et], plates))
        np.testing.assert_array_equal(mine, ref)


@pytest.mark.timeout(300)
def test_baseline_config0_runs_on_the_oracle():
    """BASELINE config 0, the reference's own CPU-runnable case, end to end on the oracle: per-experiment statistics
    and normalisation of a batch of 8 synthetic 6x256x256 images (two experiments), then a ResNet-18-style 6-channel
    network (reference stem recipe) forward / CrossEntropy / backward / nesterov SGD, 1108 classes."""
    import torch
    torch.manual_seed(0)
    exp_of = [0, 0, 0, 0, 1, 1, 1, 1]
    planes = np.concatenate([synth_planes(11, n=4, H=256, W=256), synth_planes(12, n=4, H=256, W=256)])
    stats = [O.compute_mean_std_arrays(planes[:4]), O.compute_mean_std_arrays(planes[4:])]
    assert abs(stats[0][0][0] - stats[1][0][0]) > 1e-3                    # the experiments differ (scale factor)
    x = np.stack([O.transform(planes[i], *stats[exp_of[i]]) for i in range(8)])
    for e in (0, 1):                                                      # normalised: zero mean, unit std per channel
        sel = x[[i for i in range(8) if exp_of[i] == e]]
        np.testing.assert_allclose(sel.mean(axis=(0, 2, 3)), 0, atol=2e-5)
        np.testing.assert_allclose(sel.std(axis=(0, 2, 3)), 1, rtol=1e-4)
    net = O.resnet18_6ch(1108, seed=0)
    net.train()
    opt = O.sgd_reference(net.parameters(), lr=0.0005 * 8)                # main.py:71
    y = torch.arange(8) * 137 % 1108
    xt = torch.from_numpy(x)
    losses = []
    for _ in range(3):
        opt.zero_grad()
        loss = torch.nn.CrossEntropyLoss()(net(xt), y)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]
    assert tuple(net.conv1.weight.shape) == (64, 6, 7, 7)
