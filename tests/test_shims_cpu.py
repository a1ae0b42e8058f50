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


# ---------------------------------------------------------------- ImagesDS / test(): host logic with stand-in kernels
def _tree(root, S=32, jpeg=False):
    """Two experiments x one plate x (B02 negative control, C03 positive control, three sample wells) x two sites."""
    import cv2
    import pandas as pd
    rows, ctrl, planes = [], [], {}
    for ei, exp in enumerate(("HEPG2-01", "U2OS-02")):
        for split in ("train", "test"):
            d = root / split / exp / "Plate1"
            d.mkdir(parents=True)
            for wi, well in enumerate(("B02", "C03", "D04", "E05", "F06")):
                for site in (1, 2):
                    p = synth_planes(1000 * ei + 10 * wi + site, n=1, H=S, W=S)[0]
                    for ch in range(6):
                        path = str(d / ("%s_s%d_w%d.jpeg" % (well, site, ch + 1)))
                        if jpeg:
                            cv2.imwrite(path, p[ch], [cv2.IMWRITE_JPEG_QUALITY, 95])
                            p[ch] = cv2.imread(path, cv2.IMREAD_GRAYSCALE)
                        else:                                   # lossless PNG bytes under the .jpeg name
                            open(path, "wb").write(cv2.imencode(".png", p[ch])[1].tobytes())
                    planes[(split, exp, well, site)] = p
        for wi, well in enumerate(("B02", "C03", "D04", "E05", "F06")):
            rec = {"id_code": "%s_1_%s" % (exp, well), "experiment": exp, "plate": 1, "well": well, "sirna": 10 * ei + wi}
            if well == "B02":
                ctrl.append(dict(rec, well_type="negative_control"))
            elif well == "C03":
                ctrl.append(dict(rec, well_type="positive_control"))
            else:
                rows.append(rec)
    return pd.DataFrame(rows), pd.DataFrame(ctrl), planes


@pytest.fixture
def fake_loader(monkeypatch, fake_kernels):
    def load_norm_aug(src, src_idx, exp_id, aug, crop_yx, norm_m, norm_d, out_hw, out_format, out=None):
        assert out_format == ops.OUT_F32_NCHW
        outs = []
        for b in range(src_idx.numel()):
            c, e = int(aug[b]), int(exp_id[b])
            img = O.d4_augment(np.moveaxis(src[int(src_idx[b])].numpy(), 0, 2), bool(c & 1), bool(c & 2), (c >> 2) & 3,
                               bool(c & 16))
            y0, x0 = (int(v) for v in crop_yx[b])
            x = img[y0:y0 + out_hw[0], x0:x0 + out_hw[1]].astype(np.float32)
            outs.append(np.moveaxis((x - norm_m[e].numpy()) * norm_d[e].numpy(), 2, 0))
        return torch.from_numpy(np.stack(outs))

    def load_norm_affine(src, src_idx, exp_id, flips, M, crop_yx, norm_m, norm_d, out_hw, out_format, out=None):
        outs = []
        for b in range(src_idx.numel()):
            c, e = int(flips[b]), int(exp_id[b])
            img = np.moveaxis(src[int(src_idx[b])].numpy(), 0, 2)
            img = img[::-1] if c & 1 else img
            img = img[:, ::-1] if c & 2 else img
            y0, x0 = (int(v) for v in crop_yx[b])
            x = O.warp_affine_u8(np.ascontiguousarray(img), M[b].numpy())[y0:y0 + out_hw[0], x0:x0 + out_hw[1]]
            outs.append(np.moveaxis((x.astype(np.float32) - norm_m[e].numpy()) * norm_d[e].numpy(), 2, 0))
        return torch.from_numpy(np.stack(outs))

    monkeypatch.setattr(ops, "load_norm_aug", load_norm_aug)
    monkeypatch.setattr(ops, "load_norm_affine", load_norm_affine)


@pytest.mark.parametrize("decode", ["host", "gpu"])
def test_images_ds_host_logic(tmp_path, fake_loader, decode):
    """Per-experiment statistics reach the right images, the image / negative / positive thirds keep their order,
    crops and modes follow dataloader.py:128-209 — for both decode paths and both augmentations."""
    from recursion_cellular_image_classification_b200.cell_classifier import dataloader as dl
    df, dfc, planes = _tree(tmp_path, jpeg=decode == "gpu")
    stats = {"U2OS-02": {"mean": np.linspace(0.2, 0.3, 6), "std": np.linspace(0.1, 0.2, 6)},
             "HEPG2-01": {"mean": np.linspace(0.05, 0.1, 6), "std": np.linspace(0.04, 0.08, 6)}}
    kw = dict(verbose=False, device="cpu", decode=decode)

    def expect(split, exp, well, site, **t):
        return O.transform(planes[(split, exp, well, site)], stats[exp]["mean"], stats[exp]["std"], **t)

    def is_one_of(x, cands):
        return any(np.array_equal(x.numpy().view(np.uint32), c.view(np.uint32)) for c in cands)

    val = dl.ImagesDS(df, dfc, stats, str(tmp_path), "val", crop=20, **kw)
    assert len(val) == 6
    for i in (0, 4):                                            # one sample of each experiment
        x, label = val[i]
        exp, well = df.iloc[i].experiment, df.iloc[i].well
        assert tuple(x.shape) == (3, 6, 20, 20) and label == int(df.iloc[i].sirna)
        t = dict(crop_yx=(6, 6), out_hw=(20, 20))
        assert is_one_of(x[0], [expect("train", exp, well, s, **t) for s in (1, 2)])
        assert is_one_of(x[1], [expect("train", exp, "B02", s, **t) for s in (1, 2)])
        assert is_one_of(x[2], [expect("train", exp, "C03", s, **t) for s in (1, 2)])
    tst = dl.ImagesDS(df, dfc, stats, str(tmp_path), "test", **kw)
    x, idc = tst[5]
    exp, well = df.iloc[5].experiment, df.iloc[5].well
    assert tuple(x.shape) == (6, 6, 32, 32) and idc == df.iloc[5].id_code
    for j, (w, s) in enumerate([(well, 1), (well, 2), ("B02", 1), ("B02", 2), ("C03", 1), ("C03", 2)]):
        assert is_one_of(x[j], [expect("test", exp, w, s)])
    for augment in ("d4", "rotate"):
        trn = dl.ImagesDS(df, dfc, stats, str(tmp_path), "train", crop=20, augment=augment, **kw)
        items = [trn.raw_item(i) for i in (1, 3, 4)]
        batch = dl.collate_raw(items)
        got = trn.device_batch(batch, torch.device("cpu"), out_format=ops.OUT_F32_NCHW).numpy().reshape(3, 3, 6, 20, 20)
        first = trn.device_batch(batch, torch.device("cpu"), out_format=ops.OUT_F32_NCHW, first_only=True).numpy()
        for bi, (i, item) in enumerate(zip((1, 3, 4), items)):
            exp = df.iloc[i].experiment
            src = (batch["planes"][bi].numpy() if decode == "host" else
                   np.stack([O.jpeg_decode_gray(b) for b in item["jpeg"]]).reshape(3, 6, 32, 32))
            for g_ in range(3):
                c, crop = int(item["codes"][g_]), tuple(int(v) for v in item["crops"][g_])
                if augment == "d4":
                    ref = O.transform(src[g_], stats[exp]["mean"], stats[exp]["std"], vflip=bool(c & 1),
                                      hflip=bool(c & 2), k=(c >> 2) & 3, crop_yx=crop, out_hw=(20, 20))
                else:
                    img = np.moveaxis(src[g_], 0, 2)
                    img = img[::-1] if c & 1 else img
                    img = img[:, ::-1] if c & 2 else img
                    w = O.warp_affine_u8(np.ascontiguousarray(img), item["mats"][g_].numpy())
                    w = w[crop[0]:crop[0] + 20, crop[1]:crop[1] + 20]
                    ref = np.ascontiguousarray(np.moveaxis(O.normalize(w, stats[exp]["mean"], stats[exp]["std"]), 2, 0))
                assert np.array_equal(got[bi, g_].view(np.uint32), ref.view(np.uint32))
            assert np.array_equal(first[bi], got[bi, 0])
    # the fast path of train() / evaluate() / test(): items without the control wells (what DenseNet's single linear
    # head can use) — one random site in train/val, both sites in test, same transforms
    import random as _random
    for mode, G in (("val", 1), ("test", 2)):
        ds = dl.ImagesDS(df, dfc, stats, str(tmp_path), mode, crop=20, **kw)
        _random.seed(5)
        item = ds.raw_item(4, controls=False)
        assert item["codes"].shape == (G,) and item["crops"].shape == (G, 2)
        x = ds.device_batch(dl.collate_raw([item]), torch.device("cpu"), out_format=ops.OUT_F32_NCHW)
        exp, well = df.iloc[4].experiment, df.iloc[4].well
        split = "train" if mode == "val" else "test"
        t = dict(crop_yx=(6, 6), out_hw=(20, 20)) if mode == "val" else {}
        assert x.shape[0] == G
        if mode == "val":
            assert is_one_of(x[0], [expect(split, exp, well, s, **t) for s in (1, 2)])
        else:
            for s in (1, 2):
                assert is_one_of(x[s - 1], [expect(split, exp, well, s)])


def test_test_shim_host_logic_matches_reference_golden(golden_dir, monkeypatch):
    """test() with stand-ins for the two kernels (oracle softmax/mask/rescale + greedy): batching, view stacking, the
    plate / plate-group column plumbing and the float64 return type, against the reference's own output."""
    import os
    import pandas as pd
    from recursion_cellular_image_classification_b200.cell_classifier import test as shim
    g = np.load(os.path.join(golden_dir, "assign_golden.npz"))

    def tta_softmax_avg_mask(logits, plate=None, group_col=None):
        probs = np.mean([O.softmax(v) for v in logits.numpy()], axis=0).astype(np.float32)
        return torch.from_numpy(O.mask_rescale(probs, group_col.numpy(), plate.numpy()))

    monkeypatch.setattr(ops, "tta_softmax_avg_mask", tta_softmax_avg_mask)
    monkeypatch.setattr(ops, "greedy_assign", lambda p: torch.from_numpy(O.greedy_assign(p.numpy()).astype(np.int32)))
    logits = g["logits64"]
    N = logits.shape[0]

    class DS(torch.utils.data.Dataset):
        def __len__(self):
            return N

        def __getitem__(self, i):
            return torch.tensor([float(i)]), "id%d" % i

    res = shim.test(pd.DataFrame({"plate": g["plates64"]}), DS(), g["pg64"], int(g["et64"]),
                    lambda x: torch.from_numpy(logits[x[:, 0].long().numpy()]), bs=16, num_workers=0, device="cpu")
    assert res.dtype == np.float64
    np.testing.assert_array_equal(res, g["res64"])


def test_test_shim_with_dataset_and_eight_d4_views(tmp_path, fake_loader, monkeypatch):
    """test() over an ImagesDS with tta_views=8 (BASELINE config 4: 2 sites x 8 D4 views): every view re-normalises
    and re-augments the well's OWN two sites (the first third of the reference's item, models.py:46-49 — control wells
    never reach a single-image linear head and are not loaded), logits are averaged over the sites, probabilities over
    the views, then mask + rescale + greedy assignment — against the same computation spelled out with the oracle."""
    from recursion_cellular_image_classification_b200.cell_classifier import dataloader as dl
    from recursion_cellular_image_classification_b200.cell_classifier import test as shim
    df, dfc, planes = _tree(tmp_path)
    stats = {"U2OS-02": {"mean": np.linspace(0.2, 0.3, 6), "std": np.linspace(0.1, 0.2, 6)},
             "HEPG2-01": {"mean": np.linspace(0.05, 0.1, 6), "std": np.linspace(0.04, 0.08, 6)}}
    C = 8
    proj = torch.from_numpy(np.random.default_rng(3).standard_normal((6 * 4, C)).astype(np.float32))

    def fold(x):                              # [n,6,H,W] float32 -> [n,24]: channel means of the four image quadrants
        n, c, H, W = x.shape                  # (sensitive to flips and rotations)
        q = x.reshape(n, c, 2, H // 2, 2, W // 2).mean(dim=(3, 5))
        return q.reshape(n, -1)

    def model(x):
        return (fold(x) @ proj) * 3.0

    def tta_softmax_avg_mask(logits, plate=None, group_col=None):
        probs = np.mean([O.softmax(v) for v in logits.numpy()], axis=0).astype(np.float32)
        return torch.from_numpy(O.mask_rescale(probs, group_col.numpy(), plate.numpy()))

    monkeypatch.setattr(ops, "tta_softmax_avg_mask", tta_softmax_avg_mask)
    monkeypatch.setattr(ops, "greedy_assign", lambda p: torch.from_numpy(O.greedy_assign(p.numpy()).astype(np.int32)))
    ds = dl.ImagesDS(df, dfc, stats, str(tmp_path), "test", verbose=False, device="cpu")
    real = dl.ImagesDS.device_batch         # test() asks for the stem layout; the stand-in loader produces fp32 NCHW
    monkeypatch.setattr(dl.ImagesDS, "device_batch",
                        lambda self, batch, dev, out_format=None, first_only=False:
                        real(self, batch, dev, ops.OUT_F32_NCHW, first_only))
    pg = np.stack([np.array([1, 1, 1, 1, 2, 2, 2, 2])] * 4, axis=1)        # classes 0-3 on plate 1, 4-7 on plate 2
    res = shim.test(df, ds, pg, 1, model, bs=4, num_workers=0, device="cpu", tta_views=8)
    # the same thing, spelled out
    codes = [ops.aug_code(v, False, k) for v in (False, True) for k in range(4)]
    assert len(set(codes)) == 8
    views = []
    for c in codes:
        rows = []
        for i in range(len(df)):
            exp, well = df.iloc[i].experiment, df.iloc[i].well
            imgs = [planes[("test", exp, well, s)] for s in (1, 2)]
            x = np.stack([O.transform(im, stats[exp]["mean"], stats[exp]["std"], vflip=bool(c & 1), hflip=bool(c & 2),
                                      k=(c >> 2) & 3) for im in imgs])
            rows.append(model(torch.from_numpy(x)).mean(0).numpy())
        views.append(np.stack(rows))
    probs = np.mean([O.softmax(v) for v in views], axis=0).astype(np.float32)
    ref = O.greedy_assign(O.mask_rescale(probs, pg[:, 1], df.plate.values))
    np.testing.assert_array_equal(res, ref.astype(np.float64))
    assert set(res.astype(int)) <= {0, 1, 2, 3}                              # every well sits on plate 1


def test_densenet_parameter_layout_is_torchvisions(tmp_path):
    """The flat parameter / buffer layout of the executor (host side of DenseNet121, no GPU): names, order and shapes
    are torchvision densenet121's (6-channel stem), the element counts are what librxb plans for, checkpoints load
    with and without DataParallel's `module.` prefix (main.py:147), and the stem follows the reference recipe."""
    import ctypes
    from recursion_cellular_image_classification_b200 import _lib
    from recursion_cellular_image_classification_b200.cell_classifier.models import DenseNet121, TwoSitesNN
    ref = O.densenet121_6ch(1108, seed=4)
    net = DenseNet121(1108, device="cpu", seed=1)
    named = [(k, tuple(v.shape)) for k, v in ref.named_parameters()]
    assert [(k, tuple(s)) for k, s in net.specs] == named
    bufs = [(k, tuple(v.shape)) for k, v in ref.named_buffers() if not k.endswith("num_batches_tracked")]
    assert [(k, tuple(s)) for k, s in net.buf_specs] == bufs
    cfg = _lib.Dn121Config(8, 256, 256, 1108, 1e-5, 0.1)
    lib = _lib.load()
    assert lib.rxb_dn121_param_count(ctypes.byref(cfg)) == net.flat.numel() == 8098964
    assert lib.rxb_dn121_buffer_count(ctypes.byref(cfg)) == net.bn_buffers.numel()
    # fresh initialisation: BatchNorm at identity, stem = one kernel replicated over the six channels (models.py:24-26)
    w = net.view("features.conv0.weight")
    assert all(torch.equal(w[:, 0], w[:, c]) for c in range(1, 6))
    assert float(net.view("features.norm0.weight").min()) == 1.0 and float(net.buffer_view("features.norm0.running_var").min()) == 1.0
    # checkpoints
    net.load_state_dict(ref.state_dict())
    sd = net.state_dict()
    for k, v in ref.state_dict().items():
        if not k.endswith("num_batches_tracked"):
            assert torch.equal(sd[k], v), k
    net2 = TwoSitesNN(pretrained=False, nb_classes=1108, device="cpu")
    net2.load_state_dict({"module." + k: v for k, v in sd.items()})
    assert torch.equal(net2.flat.data, net.flat.data) and torch.equal(net2.bn_buffers, net.bn_buffers)
    with pytest.raises(KeyError):
        net2.load_state_dict({k: v for k, v in sd.items() if k != "classifier.bias"})
    # pretrained=True (what main.py:43 asks for on a GPU box): weights from the file RXB_PRETRAINED_DENSENET121 names go
    # through the stem surgery and keep our 1108-way head; without any source it warns and stays at random init
    import torchvision
    tv = torchvision.models.densenet121(weights=None)
    ckpt = str(tmp_path / "densenet121_imagenet_like.pth")
    torch.save(tv.state_dict(), ckpt)
    import os
    os.environ["RXB_PRETRAINED_DENSENET121"] = ckpt
    try:
        net3 = TwoSitesNN(pretrained=True, nb_classes=1108, device="cpu")
    finally:
        del os.environ["RXB_PRETRAINED_DENSENET121"]
    assert net3.pretrained_loaded
    assert torch.equal(net3.view("features.conv0.weight"),
                       torch.stack([tv.features.conv0.weight.detach().mean(1)] * 6, dim=1))
    assert torch.equal(net3.view("features.denseblock3.denselayer7.conv2.weight"),
                       tv.features.denseblock3.denselayer7.conv2.weight.detach())
    assert tuple(net3.view("classifier.weight").shape) == (1108, 1024)
    from recursion_cellular_image_classification_b200.cell_classifier import models as M
    real = M._pretrained_densenet121_state
    M._pretrained_densenet121_state = lambda: None                  # "no network, no file"
    try:
        with pytest.warns(UserWarning, match="not available"):
            net4 = TwoSitesNN(pretrained=True, nb_classes=1108, device="cpu")
    finally:
        M._pretrained_densenet121_state = real
    assert not net4.pretrained_loaded
    # torch.optim.SGD as main.py:89-93 builds it sees one flat parameter
    assert [p.numel() for p in net.parameters()] == [8098964]
    if not torch.cuda.is_available():
        with pytest.raises(_lib.RxbError):                   # compute needs the GPU: no CPU path
            net(torch.zeros(1, 6, 32, 32))


def test_bench_host_helpers():
    """bench.py's host-side pieces: the nvidia-smi sample parser behind the `clocks` key, the peaks file reader, and
    the stdout claim (only the JSON line may reach stdout)."""
    import importlib.util
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    s = bench.ClockSampler(0)
    s.samples = [["1965", "1965", "700.1", "Not Active", "Not Active", "Not Active", "Active"],
                 ["1950", "1965", "710.0", "Not Active", "Not Active", "Not Active", "Not Active"],
                 ["garbage"], ["1935", "1965", "690.0", "Not Active", "Not Active", "Not Active", "Not Active"]]
    c = s.finish()
    assert c == {"sm_mhz": 1950.0, "sm_max_mhz": 1965.0, "reasons": ["sw_power_cap"], "samples": 3}
    p = bench.measured_peaks()
    assert p["hbm_gbs"] > 1000 and p["bf16_tflops_sustained"] <= p["bf16_tflops"] * 1.001
    # a child that writes to fd 1 from C level (like NCCL's banner) before the line: stdout must hold the line alone
    code = ("import os, sys, json; sys.path.insert(0, %r); import bench; bench.claim_stdout(); os.write(1, b'NCCL version x\\n');"
            "print('python print'); print(json.dumps({'ok': 1}), file=bench.OUT, flush=True)" % root)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, check=True)
    assert json.loads(r.stdout) == {"ok": 1} and "NCCL version x" in r.stderr and "python print" in r.stderr


def test_bench_traffic_model_follows_the_fused_backward(monkeypatch):
    """bench.py's algorithmic-HBM-traffic model of one training step (the denominator of `hbm_frac_step`): the data flow
    before the dense layers' weight gradients and gradient fix-ups moved into the data-gradient kernels moved 148.7 GB
    per 128-image step, the shipped one 120.6 GB; the debug switches that restore the separate kernels
    restore their bytes."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod2", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for k in ("RXB_DBG_NO_WGFUSE", "RXB_DBG_NO_WGFUSE3", "RXB_DBG_NO_FIXFOLD", "RXB_DBG_NO_BNTAIL"):
        monkeypatch.delenv(k, raising=False)
    t = bench.step_traffic_model(128)
    assert t["conv_wgrad_1x1"] == 0 and t["conv_wgrad_3x3"] == 0
    assert abs(sum(t.values()) / 1e9 - 120.6) < 0.1
    assert abs(t["conv_dgrad_1x1"] / 1e9 - 42.26) < 0.01            # M*(128 + 3*Cin)*2 B over the 58 dense layers
    for k in ("RXB_DBG_NO_WGFUSE", "RXB_DBG_NO_WGFUSE3"):
        monkeypatch.setenv(k, "1")
    t0 = bench.step_traffic_model(128)
    assert abs(sum(t0.values()) / 1e9 - 148.66) < 0.1
    assert abs(t0["conv_wgrad_1x1"] / 1e9 - 17.93) < 0.01 and abs(t0["grad_fixup"] / 1e9 - 5.84) < 0.01
    # weak scaling: the model is linear in the batch
    assert abs(sum(bench.step_traffic_model(64).values()) * 2 - sum(t0.values())) < 1e-3 * sum(t0.values())


def test_test_shim_unwraps_dataparallel_and_default_device(monkeypatch, golden_dir):
    """main.py:94 hands test() a torch.nn.DataParallel wrapper: the shim calls the wrapped module itself (one process
    drives one GPU; the wrapper's scatter/replicate must not run).  Models built without a device follow the rank."""
    import os
    import pandas as pd
    from recursion_cellular_image_classification_b200 import parallel
    from recursion_cellular_image_classification_b200.cell_classifier import test as shim
    g = np.load(os.path.join(golden_dir, "assign_golden.npz"))
    logits = torch.from_numpy(g["logits64"])

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.zeros(1))
            self.calls = 0

        def forward(self, x):
            self.calls += 1
            return logits[x[:, 0].long()]

    def tta_softmax_avg_mask(lg, plate=None, group_col=None):
        probs = np.mean([O.softmax(v) for v in lg.numpy()], axis=0).astype(np.float32)
        return torch.from_numpy(O.mask_rescale(probs, group_col.numpy(), plate.numpy()))

    monkeypatch.setattr(ops, "tta_softmax_avg_mask", tta_softmax_avg_mask)
    monkeypatch.setattr(ops, "greedy_assign", lambda p: torch.from_numpy(O.greedy_assign(p.numpy()).astype(np.int32)))
    net = Net()
    wrapper = torch.nn.DataParallel(net)
    monkeypatch.setattr(wrapper, "forward", lambda *a, **k: (_ for _ in ()).throw(AssertionError("wrapper forward used")))

    class DS(torch.utils.data.Dataset):
        def __len__(self):
            return 64

        def __getitem__(self, i):
            return torch.tensor([float(i)]), "id%d" % i

    res = shim.test(pd.DataFrame({"plate": g["plates64"]}), DS(), g["pg64"], int(g["et64"]), wrapper, bs=16, num_workers=0,
                    device="cpu")
    np.testing.assert_array_equal(res, g["res64"])
    assert net.calls == 4
    if not torch.cuda.is_available():
        assert parallel.default_device() == "cpu"
        monkeypatch.setenv("WORLD_SIZE", "4")
        monkeypatch.setenv("LOCAL_RANK", "3")
        assert parallel.default_device() == "cpu"            # no GPU: the host-side surface only


def test_pretrained_runs_freeze_everything_but_the_head_for_two_epochs(tmp_path, monkeypatch):
    """train.py:46-67: with hyperparams['pretrained'] only the head ('mlp' / 'classifier' children) learns during
    epochs 1-2, everything from epoch 3.  The loop asks the model for head-only updates accordingly; a stock 3-channel
    torchvision checkpoint (what `pretrained=True` would download) loads through the reference's stem surgery."""
    from test_parallel_cpu import _FeatureDS, _LinearNet
    from recursion_cellular_image_classification_b200.cell_classifier import train as T
    from recursion_cellular_image_classification_b200.cell_classifier.models import DenseNet121
    monkeypatch.chdir(tmp_path)
    monkeypatch.delenv("WORLD_SIZE", raising=False)

    def softmax_ce(logits, target, grad_scale=None):
        lp = torch.log_softmax(logits.double(), dim=1)
        return -lp[torch.arange(len(target)), target], None

    monkeypatch.setattr(ops, "softmax_ce", softmax_ce)

    class Net(_LinearNet):
        def __init__(self):
            super().__init__()
            self.flags = []

        def sgd_step(self, B, H, W, lr, head_only=False, **kw):
            self.flags.append(head_only)
            super().sgd_step(B, H, W, lr, **kw)

    for pretrained, expect in ((True, [True, True, False, False]), (False, [False] * 4)):
        net = Net()
        opt = torch.optim.SGD([net.flat], lr=0.05, momentum=0.9, nesterov=True, weight_decay=3e-5)
        hp = {"bs": 8, "nb_epochs": 4, "scheduler": True, "lr": 0.05, "early_stopping": False, "patience": 10,
              "pretrained": pretrained}
        T.train("fz", _FeatureDS(8, 1), _FeatureDS(8, 1), net, opt, hp, num_workers=0, device="cpu", debug=True)
        assert net.flags == expect
    # the head of the real model is the tail of its flat buffer; an ImageNet-shaped checkpoint loads non-strictly
    from torchvision import models as tvm
    torch.manual_seed(0)
    tv = tvm.densenet121(weights=None)
    net = DenseNet121(1108, device="cpu")
    head = net.view("classifier.weight").clone()
    with pytest.raises(RuntimeError):
        net.load_state_dict(tv.state_dict())                           # 1000-class head, strict
    net.load_state_dict(tv.state_dict(), strict=False)
    assert torch.equal(net.view("features.conv0.weight"),
                       torch.stack([tv.features.conv0.weight.detach().mean(1)] * 6, dim=1))      # models.py:24-26
    assert torch.equal(net.view("classifier.weight"), head)
    b, e = net.head_range()
    assert e == net.flat.numel() and e - b == 1024 * 1108 + 1108
    assert net.flat.data[b:b + 5].tolist() == net.view("classifier.weight").reshape(-1)[:5].tolist()


def test_train_writes_the_references_tensorboard_scalars(tmp_path, monkeypatch):
    """board/<id> holds 'training/loss' and 'lr/group_0' per iteration and 'validation/accuracy' / 'validation/loss'
    per epoch (train.py:114-135), with the cosine learning rate of each epoch."""
    from tensorboard.backend.event_processing.event_accumulator import EventAccumulator
    from test_parallel_cpu import _FeatureDS, _LinearNet
    from recursion_cellular_image_classification_b200.cell_classifier import train as T
    monkeypatch.chdir(tmp_path)
    monkeypatch.delenv("WORLD_SIZE", raising=False)

    def softmax_ce(logits, target, grad_scale=None):
        lp = torch.log_softmax(logits.double(), dim=1)
        return -lp[torch.arange(len(target)), target], None

    monkeypatch.setattr(ops, "softmax_ce", softmax_ce)
    net = _LinearNet()
    opt = torch.optim.SGD([net.flat], lr=0.05, momentum=0.9, nesterov=True, weight_decay=3e-5)
    hp = {"bs": 4, "nb_epochs": 3, "scheduler": True, "lr": 0.05, "early_stopping": False, "patience": 10,
          "pretrained": False}
    hist = T.train("tb", _FeatureDS(8, 1), _FeatureDS(8, 1), net, opt, hp, num_workers=0, device="cpu", debug=True)
    acc = EventAccumulator(str(tmp_path / "board" / "tb"))
    acc.Reload()
    tags = set(acc.Tags()["scalars"])
    assert tags == {"training/loss", "lr/group_0", "validation/accuracy", "validation/loss"}
    tl = acc.Scalars("training/loss")
    assert [e.step for e in tl] == [1, 2, 3, 4, 5, 6] and tl[-1].value < tl[0].value          # 2 iterations x 3 epochs
    lrs = [e.value for e in acc.Scalars("lr/group_0")]
    np.testing.assert_allclose(lrs, [T.cosine_lr(0.05, e, 3) for e in (0, 0, 1, 1, 2, 2)], rtol=1e-6)
    vl = acc.Scalars("validation/loss")
    assert [e.step for e in vl] == [0, 1, 2, 3]
    np.testing.assert_allclose([e.value for e in vl], [h["val_loss"] for h in hist], rtol=1e-6)


def test_train_early_stopping_and_best_checkpoint(tmp_path, monkeypatch):
    """train.py:74-80,88-96: with early_stopping the run ends `patience` epochs after the last improvement of the
    validation accuracy; the checkpoint on disk is the best epoch's, not the last one's."""
    from test_parallel_cpu import _FeatureDS, _LinearNet
    from recursion_cellular_image_classification_b200.cell_classifier import train as T
    monkeypatch.chdir(tmp_path)
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    accs = iter([0.10, 0.30, 0.20, 0.25, 0.30, 0.90, 0.95])           # epochs 0..: best at epoch 1, then no improvement
    monkeypatch.setattr(T, "evaluate", lambda model, ds, bs, nw, dev: (next(accs), 1.0))
    net = _LinearNet()
    snapshots = []
    real_step = net.sgd_step

    def sgd_step(*a, **k):
        real_step(*a, **k)
        snapshots.append(net.flat.data.clone())

    net.sgd_step = sgd_step
    opt = torch.optim.SGD([net.flat], lr=0.05, momentum=0.9, nesterov=True, weight_decay=3e-5)
    hp = {"bs": 8, "nb_epochs": 10, "scheduler": False, "lr": 0.05, "early_stopping": True, "patience": 3,
          "pretrained": False, "tensorboard": False}
    hist = T.train("es", _FeatureDS(8, 1), _FeatureDS(8, 1), net, opt, hp, num_workers=0, device="cpu", debug=True)
    assert [h["epoch"] for h in hist] == [0, 1, 2, 3, 4]              # stopped 3 epochs after the best (epoch 1)
    assert not (tmp_path / "board").exists()
    sd = torch.load("models/best_model_es.pth")
    assert torch.equal(sd["module.flat"], snapshots[0])               # weights as they were after epoch 1


def test_config4_case_is_well_separated(tmp_path):
    """The seeded config-4 case the GPU parity test runs (tests/c4_case.py), oracle side only: the expected
    assignment (two collisions resolved by the greedy loop) does not move under logit noise three times the 2e-2
    tolerance, so a GPU/oracle disagreement there is an error, not a tie."""
    import c4_case as C
    from recursion_cellular_image_classification_b200.synth import synth_plate_groups
    pg = synth_plate_groups(3)
    df, dfc, planes = C.write_tree(str(tmp_path))
    net, classes = C.build_oracle_model(planes, pg)
    L = C.oracle_logits(net, planes, 8)
    _, res = C.oracle_assign(L, pg, df.plate.values)
    assert list(res.astype(int)) == classes
    assert len(set(classes)) == 6 and all(pg[c, C.EXPERIMENT_TYPE] == p for c, p in zip(classes, C.PLATES))
    rng = np.random.default_rng(0)
    for _ in range(10):
        noise = rng.standard_normal(L.shape).astype(np.float32)
        noise *= 0.06 * np.linalg.norm(L) / np.linalg.norm(noise)
        np.testing.assert_array_equal(C.oracle_assign(L + noise, pg, df.plate.values)[1], res)


def test_two_sites_resnet50_layout_is_the_references(tmp_path):
    """Host side of TwoSitesResNet50 (no GPU): parameter / buffer names, order and shapes are the reference model's
    own (through the oracle restatement that is pinned against the reference's golden logits), a seed reproduces the
    reference constructor's initial weights, `module.`-prefixed checkpoints load, and TwoSitesNN(trunk=...) selects it."""
    import ctypes
    from oracle import oracle_np as O
    from recursion_cellular_image_classification_b200 import _lib
    from recursion_cellular_image_classification_b200.cell_classifier.models import (TwoSitesNN, TwoSitesResNet50,
                                                                                      two_sites_resnet50_param_specs)
    ref = O.two_sites_resnet50(seed=5)
    specs, bufs = two_sites_resnet50_param_specs()
    assert [(n, tuple(p.shape)) for n, p in ref.named_parameters()] == [(n, tuple(s)) for n, s in specs]
    assert [n for n, _ in ref.named_buffers() if "num_batches" not in n] == [n for n, _ in bufs]
    net = TwoSitesResNet50(device="cpu", seed=5)
    sd = net.state_dict()
    for k, v in ref.state_dict().items():
        if "num_batches" not in k:
            assert torch.equal(sd[k], v), k
    cfg = _lib.Rn50Config(2, 3, 64, 64, 1108, 1024, 1e-5)
    lib = _lib.load()
    assert lib.rxb_rn50_param_count(ctypes.byref(cfg)) == net.flat.numel() == sum(p.numel() for p in ref.parameters())
    assert lib.rxb_rn50_buffer_count(ctypes.byref(cfg)) == net.bn_buffers.numel()
    net2 = TwoSitesNN(pretrained=False, nb_classes=1108, trunk="resnet50", device="cpu")
    assert isinstance(net2, TwoSitesResNet50) and net2.wants_controls
    net2.load_state_dict({"module." + k: v for k, v in sd.items()})
    assert torch.equal(net2.flat.data, net.flat.data)
    net2.train()
    with pytest.raises(_lib.RxbError):
        net2(torch.zeros(1, 3, 6, 32, 32))
    assert not isinstance(TwoSitesNN(pretrained=False, device="cpu"), TwoSitesResNet50)
