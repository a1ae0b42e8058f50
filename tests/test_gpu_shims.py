"""GPU tests of the reference-facing Python surface: compute_mean_std, ImagesDS, train(), test()."""
import os

import numpy as np
import pandas as pd
import pytest
import torch

from oracle import oracle_np as O
from recursion_cellular_image_classification_b200 import compute_stats_experiments as cse
from recursion_cellular_image_classification_b200 import ops
from recursion_cellular_image_classification_b200.cell_classifier import dataloader as dl
from recursion_cellular_image_classification_b200.cell_classifier.models import DenseNet121, DummyClassifier
from recursion_cellular_image_classification_b200.cell_classifier.test import test as rxb_test
from recursion_cellular_image_classification_b200.cell_classifier.train import train as rxb_train
from recursion_cellular_image_classification_b200.synth import synth_logits, synth_plate_groups, synth_planes

pytestmark = pytest.mark.gpu


def _write_tree(root, S=64, n_wells=6):
    """A tiny RxRx1-shaped tree (lossless PNG bytes under the .jpeg names the reference uses)."""
    import cv2
    rows, ctrl = [], []
    planes = {}
    exp, plate = "HEPG2-01", 1
    wells = ["B02", "C03", "D04", "E05", "F06", "G07"][:n_wells]
    for split in ("train", "test"):
        d = os.path.join(root, split, exp, "Plate%d" % plate)
        os.makedirs(d, exist_ok=True)
        for wi, well in enumerate(wells):
            for site in (1, 2):
                p = synth_planes(100 * wi + site, n=1, H=S, W=S)[0]
                planes[(split, well, site)] = p
                for ch in range(6):
                    ok, buf = cv2.imencode(".png", p[ch])
                    with open(os.path.join(d, "%s_s%d_w%d.jpeg" % (well, site, ch + 1)), "wb") as f:
                        f.write(buf.tobytes())
    for wi, well in enumerate(wells):
        rec = {"id_code": "%s_%d_%s" % (exp, plate, well), "experiment": exp, "plate": plate, "well": well, "sirna": wi}
        if well == "B02":
            ctrl.append(dict(rec, well_type="negative_control"))
        elif well == "C03":
            ctrl.append(dict(rec, well_type="positive_control"))
        else:
            rows.append(rec)
    return pd.DataFrame(rows), pd.DataFrame(ctrl), planes, exp


def test_compute_mean_std_signature_and_values(cuda, tmp_path, golden_dir):
    """Same call as the reference (paths in, two float64[6] out), same numbers as its golden output."""
    import cv2
    g = np.load(os.path.join(golden_dir, "stats_golden.npz"))
    planes = synth_planes(int(g["seeds"][0]), n=int(g["n_per_exp"]))
    d = tmp_path / "exp0" / "Plate1"
    d.mkdir(parents=True)
    paths = []
    for i in range(planes.shape[0]):
        for ch in range(6):
            p = str(d / ("B%02d_s%d_w%d.jpeg" % (2 + i // 2, 1 + i % 2, ch + 1)))
            ok, buf = cv2.imencode(".png", planes[i, ch])
            open(p, "wb").write(buf.tobytes())
            paths.append(p)
    mean, std = cse.compute_mean_std(paths)
    assert mean.dtype == np.float64 and mean.shape == (6,) and std.shape == (6,)
    np.testing.assert_allclose(mean, g["mean"][0], rtol=1e-5)
    np.testing.assert_allclose(std, g["std"][0], rtol=1e-5)
    vm, vs = cse.compute_mean_std(paths, mean=mean, std=std)
    np.testing.assert_allclose(vm, 0, atol=1e-9)
    np.testing.assert_allclose(vs, 1, rtol=1e-9)


def test_images_ds_items_match_oracle(cuda, tmp_path):
    root = str(tmp_path)
    df, dfc, planes, exp = _write_tree(root)
    stats = {exp: {"mean": np.linspace(0.05, 0.1, 6), "std": np.linspace(0.04, 0.08, 6)}}
    ds = dl.ImagesDS(df, dfc, stats, root, "val", verbose=False, crop=32)
    x, label = ds[1]
    assert x.dtype == torch.float32 and tuple(x.shape) == (3, 6, 32, 32) and isinstance(label, int)
    # val mode: centre crop, no flips; negative control is well B02, positive control C03 (sites random)
    well = df.iloc[1].well
    cands = [O.transform(planes[("train", well, s)], stats[exp]["mean"], stats[exp]["std"], crop_yx=(16, 16),
                         out_hw=(32, 32)) for s in (1, 2)]
    assert any(np.array_equal(x[0].numpy().view(np.uint32), c.view(np.uint32)) for c in cands)
    negs = [O.transform(planes[("train", "B02", s)], stats[exp]["mean"], stats[exp]["std"], crop_yx=(16, 16),
                        out_hw=(32, 32)) for s in (1, 2)]
    assert any(np.array_equal(x[1].numpy().view(np.uint32), c.view(np.uint32)) for c in negs)
    # test mode: both sites of image / negative / positive control, no crop, id_code label (dataloader.py:182-209)
    dst = dl.ImagesDS(df, dfc, stats, root, "test", verbose=False)
    xt, idc = dst[0]
    assert tuple(xt.shape) == (6, 6, 64, 64) and idc == df.iloc[0].id_code
    w0 = df.iloc[0].well
    for s in (1, 2):
        ref = O.transform(planes[("test", w0, s)], stats[exp]["mean"], stats[exp]["std"])
        assert np.array_equal(xt[s - 1].numpy().view(np.uint32), ref.view(np.uint32))
    # train mode: D4 augmentation parameters are drawn explicitly and reproduced by the oracle
    dtr = dl.ImagesDS(df, dfc, stats, root, "train", verbose=False, crop=32)
    item = dtr.raw_item(2)
    batch = dl.collate_raw([item])
    got = dtr.device_batch(batch, cuda, out_format=ops.OUT_F32_NCHW).cpu().numpy()
    for g_ in range(3):
        c = int(item["codes"][g_])
        ref = O.transform(item["planes"][g_].numpy(), stats[exp]["mean"], stats[exp]["std"], vflip=bool(c & 1),
                          hflip=bool(c & 2), k=(c >> 2) & 3, crop_yx=tuple(int(v) for v in item["crops"][g_]),
                          out_hw=(32, 32))
        assert np.array_equal(got[g_].view(np.uint32), ref.view(np.uint32))


def test_train_runs_and_saves_reference_style_checkpoint(cuda, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    root = str(tmp_path / "data")
    df, dfc, _, exp = _write_tree(root, S=64)
    stats = {exp: {"mean": np.full(6, 0.08), "std": np.full(6, 0.06)}}
    ds_train = dl.ImagesDS(df, dfc, stats, root, "train", verbose=False)
    ds_val = dl.ImagesDS(df, dfc, stats, root, "val", verbose=False)
    model = DenseNet121(nb_classes=1108, device=cuda)
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, nesterov=True, weight_decay=3e-5)  # main.py:89-93
    hp = {"bs": 4, "nb_epochs": 2, "scheduler": True, "lr": 0.01, "early_stopping": False, "patience": 10,
          "pretrained": False, "crop": 64}
    w0 = model.flat.detach().clone()
    hist = rxb_train("unit", ds_train, ds_val, model, opt, hp, num_workers=0, device="cuda", debug=True)
    assert len(hist) == 3 and all(np.isfinite(h["val_loss"]) for h in hist)
    assert not torch.equal(w0, model.flat.detach())
    sd = torch.load("models/best_model_unit.pth")
    assert all(k.startswith("module.") for k in sd)                      # DataParallel-style keys (main.py:147)
    m2 = DenseNet121(nb_classes=1108, device=cuda)
    m2.load_state_dict(sd)


def test_test_matches_reference_golden_through_the_shim(cuda, golden_dir):
    """test() end to end with a logits callable, like the reference is driven with DummyClassifier: same class ids
    as the reference produced (tests/golden/assign_golden.npz)."""
    g = np.load(os.path.join(golden_dir, "assign_golden.npz"))
    logits = g["logits64"]
    N = logits.shape[0]
    df = pd.DataFrame({"plate": g["plates64"]})

    class DS(torch.utils.data.Dataset):
        def __len__(self):
            return N

        def __getitem__(self, i):
            return torch.tensor([float(i)]), "id%d" % i

    def model(x):
        return torch.from_numpy(logits[x[:, 0].long().cpu().numpy()])

    res = rxb_test(df, DS(), g["pg64"], int(g["et64"]), model, bs=16, num_workers=0, device="cuda")
    assert res.dtype == np.float64 and res.shape == (N,)
    np.testing.assert_array_equal(res, g["res64"])
    # the reference's fake backend also runs through it
    res2 = rxb_test(df, DS(), g["pg64"], 0, DummyClassifier(1108), bs=16, num_workers=0, device="cuda")
    assert res2.shape == (N,)


@pytest.mark.parametrize("tta_views", [1, 8])
def test_test_on_images_ds_matches_the_fp32_oracle(cuda, tmp_path, tta_views):
    """BASELINE config 4 end to end on the device: ImagesDS(test) -> fused loader (8 D4 views) -> the native
    DenseNet-121 -> softmax / view mean / plate-group mask / rescale -> greedy assignment, against the same computation
    spelled out with the fp32 oracle (torchvision densenet121 + the numpy loader + test.py's loop) on the seeded
    well-separated case of tests/c4_case.py: identical assignments, probabilities and logits within 2e-2."""
    import c4_case as C
    pg = synth_plate_groups(3)
    df, dfc, planes = C.write_tree(str(tmp_path))
    ref, classes = C.build_oracle_model(planes, pg)
    want_logits = C.oracle_logits(ref, planes, tta_views)
    want_probs, want = C.oracle_assign(want_logits, pg, df.plate.values)
    assert list(want.astype(int)) == [classes[i] for i in (0, 1, 2, 3, 4, 5)]     # collisions resolved by the greedy loop
    stats = {C.EXP: {"mean": C.MEAN, "std": C.STD}}
    ds = dl.ImagesDS(df, dfc, stats, str(tmp_path), "test", verbose=False)
    net = DenseNet121(nb_classes=1108, device=cuda)
    net.load_state_dict(ref.state_dict())
    net.eval()
    from recursion_cellular_image_classification_b200.cell_classifier.test import predict_probs
    got_probs = predict_probs(df, ds, pg, C.EXPERIMENT_TYPE, net, bs=4, num_workers=0, device="cuda",
                              tta_views=tta_views).cpu().numpy()
    got = rxb_test(df, ds, pg, C.EXPERIMENT_TYPE, net, bs=4, num_workers=0, device="cuda", tta_views=tta_views)
    rel_p = np.linalg.norm(got_probs - want_probs) / np.linalg.norm(want_probs)
    # the reference-layout item [6,6,H,W] (both sites of image / negative / positive control) through forward():
    # only the first third reaches DenseNet's head (models.py:46-49), averaged — compared with the oracle's logits
    x6 = torch.stack([ds[i][0] for i in range(len(df))]).to(cuda)                  # [N,6,6,H,W] float32
    got_logits = net(x6).cpu().numpy()
    rel_l = np.linalg.norm(got_logits - want_logits[0]) / np.linalg.norm(want_logits[0])
    print("\nconfig 4, %d views: probabilities rel %.4g, identity-view logits rel %.4g, top probabilities %s" %
          (tta_views, rel_p, rel_l, np.round(got_probs.max(1), 4)))
    np.testing.assert_array_equal(got, want)
    assert rel_l < 2e-2, rel_l
    assert rel_p < 2e-2, rel_p
