"""GPU parity tests for the device JPEG decoder (rxb_jpeg_decode_gray): bit-exact against
cv2.imdecode(buf, -1) — the reference's call (dataloader.py:141-146) — executed in the test, against the oracle
restatement, and against tests/golden/jpeg_golden.npz (files written by the reference's png_to_jpeg converter)."""
import numpy as np
import pytest
import torch

from oracle import oracle_np as O
from recursion_cellular_image_classification_b200 import _lib, ops
from recursion_cellular_image_classification_b200.synth import synth_planes
from test_oracle_cpu import _jpeg_cases, _jpeg_golden

pytestmark = pytest.mark.gpu


def _decode(cuda, buffers, hw, **kw):
    blob, offsets = ops.pack_jpeg_buffers(buffers)
    out = ops.jpeg_decode_gray(blob.to(cuda), offsets.to(cuda), hw, **kw)
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("parallel", [True, False])
def test_jpeg_decode_matches_reference_golden(cuda, golden_dir, parallel):
    for buf, plane in _jpeg_golden(golden_dir):
        got = _decode(cuda, [buf], plane.shape, parallel=parallel)
        np.testing.assert_array_equal(got[0].cpu().numpy(), plane)
        np.testing.assert_array_equal(O.jpeg_decode_gray(buf), plane)


@pytest.mark.parametrize("parallel", [True, False])
def test_jpeg_decode_is_cv2_imdecode_over_formats(cuda, parallel):
    """Sizes that are not multiples of 8, qualities, optimised Huffman tables, restart intervals (which the parallel
    kernel hands to the single-lane kernel); files of one size are decoded in one launch (several warps per CTA, a
    partial last CTA).  parallel: all lanes decode speculative subsequences / one lane per file."""
    by_shape = {}
    for buf, ref in _jpeg_cases():
        by_shape.setdefault(ref.shape, []).append((buf, ref))
    assert sum(len(v) for v in by_shape.values()) == 135
    for hw, cases in by_shape.items():
        got = _decode(cuda, [c[0] for c in cases], hw, parallel=parallel).cpu().numpy()
        for i, (_, ref) in enumerate(cases):
            np.testing.assert_array_equal(got[i], ref)


def test_jpeg_decode_full_size_batch_feeds_stats_and_loader(cuda):
    """Twelve 512x512 q95 files = two six-channel images, decoded straight into the planar layout the statistics and
    loader kernels read."""
    import cv2
    planes = synth_planes(6, n=2)
    bufs, refs = [], []
    for i in range(2):
        for c in range(6):
            ok, b = cv2.imencode(".jpg", planes[i, c], [cv2.IMWRITE_JPEG_QUALITY, 95])
            bufs.append(b.tobytes())
            refs.append(cv2.imdecode(b, -1))
    assert ops.jpeg_frame_size(bufs[0]) == (512, 512)
    got = _decode(cuda, bufs, (512, 512)).view(2, 6, 512, 512)
    ref = np.stack(refs).reshape(2, 6, 512, 512)
    np.testing.assert_array_equal(got.cpu().numpy(), ref)
    np.testing.assert_array_equal(_decode(cuda, bufs, (512, 512), parallel=False).view(2, 6, 512, 512).cpu().numpy(), ref)
    smooth = [cv2.GaussianBlur(planes[0, c], (0, 0), 2.0) for c in range(6)]     # long zero runs, few bits per block
    sb = [cv2.imencode(".jpg", im, [cv2.IMWRITE_JPEG_QUALITY, 95])[1].tobytes() for im in smooth]
    flat = cv2.imencode(".jpg", np.full((512, 512), 7, np.uint8))[1].tobytes()    # hundreds of blocks per subsequence
    gs = _decode(cuda, sb + [flat], (512, 512)).cpu().numpy()
    for i, b_ in enumerate(sb + [flat]):
        np.testing.assert_array_equal(gs[i], cv2.imdecode(np.frombuffer(b_, np.uint8), -1))
    acc = ops.stats_accumulate(got, torch.zeros(2, dtype=torch.int32, device=cuda), 1)
    mean, std = ops.stats_finalize(acc)
    om, os_ = O.compute_mean_std_arrays(ref)
    np.testing.assert_allclose(mean.cpu().numpy()[0], om, rtol=1e-12)
    np.testing.assert_allclose(std.cpu().numpy()[0], os_, rtol=1e-10)


def test_jpeg_decode_status_codes_and_empty(cuda):
    import cv2
    ok, good = cv2.imencode(".jpg", synth_planes(1, 1, C=1, H=32, W=32)[0, 0], [cv2.IMWRITE_JPEG_QUALITY, 95])
    ok, prog = cv2.imencode(".jpg", np.zeros((32, 32), np.uint8), [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    ok, rgb = cv2.imencode(".jpg", np.zeros((32, 32, 3), np.uint8))
    ok, other = cv2.imencode(".jpg", np.zeros((16, 16), np.uint8))
    bufs = [good.tobytes(), b"not a jpeg at all....", prog.tobytes(), rgb.tobytes(), other.tobytes(),
            good.tobytes()[:60], good.tobytes(), good.tobytes()[:len(good) * 3 // 4], good.tobytes()[:-4]]
    ref = cv2.imdecode(good, -1)
    for parallel in (True, False):
        out, status = _decode(cuda, bufs, (32, 32), check_status=False, parallel=parallel)
        assert status.cpu().tolist() == [0, 1, 2, 2, 4, 1, 0, 5, 5]
        np.testing.assert_array_equal(out[0].cpu().numpy(), ref)
        np.testing.assert_array_equal(out[6].cpu().numpy(), ref)
    with pytest.raises(_lib.RxbError):
        _decode(cuda, bufs, (32, 32))
    empty = _decode(cuda, [], (32, 32))
    assert tuple(empty.shape) == (0, 32, 32)


def test_jpeg_decode_subset_selection(cuda):
    import cv2
    imgs = [synth_planes(30 + i, 1, C=1, H=48, W=48)[0, 0] for i in range(5)]
    bufs = [cv2.imencode(".jpg", im, [cv2.IMWRITE_JPEG_QUALITY, 95])[1].tobytes() for im in imgs]
    blob, offsets = ops.pack_jpeg_buffers(bufs)
    sel = torch.tensor([4, 0, 2], dtype=torch.int64, device=cuda)
    got = ops.jpeg_decode_gray(blob.to(cuda), offsets.to(cuda), (48, 48), select=sel).cpu().numpy()
    for j, i in enumerate([4, 0, 2]):
        np.testing.assert_array_equal(got[j], cv2.imdecode(np.frombuffer(bufs[i], np.uint8), -1))


def test_images_ds_gpu_decode_equals_host_decode(cuda, tmp_path):
    """ImagesDS(decode='gpu') — JPEG bytes travel to the device and are decoded there — returns exactly what the
    host-decode path (cv2.imdecode, dataloader.py:141-146) returns for the same draws, in every mode."""
    import glob
    import random
    import cv2
    from test_gpu_shims import _write_tree
    from recursion_cellular_image_classification_b200.cell_classifier import dataloader as dl
    root = str(tmp_path)
    df, dfc, _, exp = _write_tree(root)
    for path in glob.glob(root + "/**/*.jpeg", recursive=True):       # real JPEGs like png_to_jpeg.py writes
        img = cv2.imdecode(np.frombuffer(open(path, "rb").read(), np.uint8), -1)
        ok, buf = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, 95])
        open(path, "wb").write(buf.tobytes())
    stats = {exp: {"mean": np.linspace(0.05, 0.1, 6), "std": np.linspace(0.04, 0.08, 6)}}
    for mode, kw in (("train", dict(crop=40)), ("train", dict(crop=40, augment="rotate")), ("val", dict(crop=32)),
                     ("test", dict())):
        host = dl.ImagesDS(df, dfc, stats, root, mode, verbose=False, **kw)
        gpu = dl.ImagesDS(df, dfc, stats, root, mode, verbose=False, decode="gpu", **kw)
        for first_only in (False, True):
            random.seed(5)
            bh = dl.collate_raw([host.raw_item(i) for i in (0, 2, 3)])
            random.seed(5)
            bg = dl.collate_raw([gpu.raw_item(i) for i in (0, 2, 3)])
            assert "planes" not in bg and bg["jpeg_offsets"].numel() == 3 * bh["planes"].shape[1] * 6 + 1
            xh = host.device_batch(bh, cuda, out_format=ops.OUT_F32_NCHW, first_only=first_only)
            xg = gpu.device_batch(bg, cuda, out_format=ops.OUT_F32_NCHW, first_only=first_only)
            assert xh.shape == xg.shape and torch.equal(xh, xg)
    x, label = gpu[0]
    assert x.dtype == torch.float32 and tuple(x.shape) == (6, 6, 64, 64)
    with pytest.raises(ValueError):
        dl.ImagesDS(df, dfc, stats, root, "val", verbose=False, decode="nvjpeg")
