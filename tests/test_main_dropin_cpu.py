"""The reference's main.py, byte for byte, driven through this package's call surface (north star: "main.py drives it
as a drop-in").  CPU only: the device kernels are replaced by numpy stand-ins (the oracle) and the DenseNet executor by
a tiny stand-in model, so what is exercised is everything between main.py and the C ABI — imports, constructor
signatures, torch.optim.SGD on model.parameters(), the DataParallel wrapper, train()'s side effects
(models/best_model_<id>.pth), ImagesDS in all three modes, plate_groups / experiment_type plumbing, test()'s return
type, the submission file.  Needs /root/reference (build container only)."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_MAIN = "/root/reference/main.py"

ALIAS = {   # the maintainer's "import change" as an alias package: main.py itself stays untouched
    "__init__": "",
    "dataloader": "from recursion_cellular_image_classification_b200.cell_classifier.dataloader import *  # noqa\n"
                  "from recursion_cellular_image_classification_b200.cell_classifier.dataloader import train_test_split, ImagesDS\n",
    "models": "from recursion_cellular_image_classification_b200.cell_classifier.models import TwoSitesNN, DummyClassifier\n",
    "train": "from recursion_cellular_image_classification_b200.cell_classifier.train import train\n",
    "test": "from recursion_cellular_image_classification_b200.cell_classifier.test import test\n",
}

STANDINS = textwrap.dedent('''
    """numpy / torch-CPU stand-ins for librxb (test infrastructure): installed before main.py runs."""
    import sys
    import numpy as np
    import torch
    sys.path.insert(0, %(root)r)
    from oracle import oracle_np as O
    from recursion_cellular_image_classification_b200 import ops
    from recursion_cellular_image_classification_b200.cell_classifier import models, dataloader

    def load_norm_aug(src, src_idx, exp_id, aug, crop_yx, norm_m, norm_d, out_hw, out_format, out=None):
        outs = []
        for b in range(src_idx.numel()):
            c, e = int(aug[b]), int(exp_id[b])
            img = O.d4_augment(np.moveaxis(src[int(src_idx[b])].numpy(), 0, 2), bool(c & 1), bool(c & 2), (c >> 2) & 3)
            y0, x0 = (int(v) for v in crop_yx[b])
            x = img[y0:y0 + out_hw[0], x0:x0 + out_hw[1]].astype(np.float32)
            outs.append(np.moveaxis((x - norm_m[e].numpy()) * norm_d[e].numpy(), 2, 0))
        return torch.from_numpy(np.stack(outs))

    def tta_softmax_avg_mask(logits, plate=None, group_col=None):
        probs = np.mean([O.softmax(v) for v in logits.numpy()], axis=0).astype(np.float32)
        return torch.from_numpy(O.mask_rescale(probs, group_col.numpy(), plate.numpy()))

    def softmax_ce(logits, target, grad_scale=None):
        lp = torch.log_softmax(logits.double(), dim=1)
        return -lp[torch.arange(len(target)), target], None

    ops.load_norm_aug = load_norm_aug
    ops.tta_softmax_avg_mask = tta_softmax_avg_mask
    ops.greedy_assign = lambda p: torch.from_numpy(O.greedy_assign(p.numpy()).astype(np.int32))
    ops.softmax_ce = softmax_ce

    # the executor: logits = (per-channel mean of the batch) @ first rows of the classifier weight
    D = models.DenseNet121
    LOG = []

    def forward(self, x):
        x = x.float()
        feat = x.reshape(x.shape[0], 6, -1).mean(2) if x.dim() == 4 else x.reshape(x.shape[0], -1)[:, :6]
        return feat @ self.view("classifier.weight")[:, :6].t() + self.view("classifier.bias")

    def train_step(self, xs, target, global_batch=None, phase=-1, loss_out=None):
        if phase in (-1, 0):
            w = self.view("classifier.weight")
            w.requires_grad_(False)
            logits = forward(self, xs)
            p = torch.softmax(logits, 1)
            loss_out[0] = -(torch.log(p[torch.arange(len(target)), target]).sum() / (global_batch or len(target)))
            LOG.append(float(loss_out[0]))
            self.flat.grad.zero_()
            p[torch.arange(len(target)), target] -= 1
            self.grad_view("classifier.bias").copy_(p.sum(0) / (global_batch or len(target)))
        return loss_out

    def sgd_step(self, B, H, W, lr, momentum=0.9, weight_decay=3e-5, nesterov=True, grad_scale=1.0):
        self.flat.data.add_(self.flat.grad, alpha=-lr)
        self._weights_dirty = True

    def phase_grad_range(self, B, H, W, phase):
        n, L = 5, self.flat.numel()
        return (L * (n - 1 - phase)) // n, (L * (n - phase)) // n

    D.forward, D.train_step, D.sgd_step, D.phase_grad_range = forward, train_step, sgd_step, phase_grad_range
    real_batch = dataloader.ImagesDS.device_batch
    dataloader.ImagesDS.device_batch = (lambda self, batch, dev, out_format=None, first_only=False:
                                        real_batch(self, batch, dev, ops.OUT_F32_NCHW, first_only))
''')


def _write_world(root):
    import pickle
    import cv2
    import pandas as pd
    S = 512                                                       # train() crops 512 unless HYPERPARAMS says otherwise
    rng = np.random.default_rng(0)

    def write_well(split, exp, plate, well, value):
        d = os.path.join(root, "data", split, exp, "Plate%d" % plate)
        os.makedirs(d, exist_ok=True)
        for site in (1, 2):
            for ch in range(6):
                img = np.full((S, S), (value + 7 * ch + site) % 256, np.uint8)
                img[:S // 2, :S // 3] += 9                        # not symmetric: flips and rotations matter
                with open(os.path.join(d, "%s_s%d_w%d.jpeg" % (well, site, ch + 1)), "wb") as f:
                    f.write(cv2.imencode(".png", img)[1].tobytes())

    rows, ctrl = [], []
    wells = ["C%02d" % i for i in range(3, 23)]
    for wi, well in enumerate(wells):
        write_well("train", "HEPG2-01", 1, well, 10 * wi)
        rows.append({"id_code": "HEPG2-01_1_" + well, "experiment": "HEPG2-01", "plate": 1, "well": well,
                     "sirna": int(rng.integers(0, 1108))})
    for well, kind in (("B02", "negative_control"), ("B03", "positive_control")):
        write_well("train", "HEPG2-01", 1, well, 200)
        ctrl.append({"id_code": "HEPG2-01_1_" + well, "experiment": "HEPG2-01", "plate": 1, "well": well, "sirna": 1108,
                     "well_type": kind})
    os.makedirs(os.path.join(root, "data", "metadata"))
    os.makedirs(os.path.join(root, "data", "full_metadata"))
    pd.DataFrame(rows).to_csv(os.path.join(root, "data", "metadata", "train.csv"), index=False)
    pd.DataFrame(ctrl).to_csv(os.path.join(root, "data", "metadata", "train_controls.csv"), index=False)
    trows, tctrl = [], []
    for exp, plate in (("HEPG2-08", 2), ("U2OS-04", 3)):
        for wi, well in enumerate(("D04", "E05", "F06")):
            write_well("test", exp, plate, well, 30 * wi + plate)
            trows.append({"id_code": "%s_%d_%s" % (exp, plate, well), "experiment": exp, "plate": plate, "well": well})
        for well, kind in (("B02", "negative_control"), ("B03", "positive_control")):
            write_well("test", exp, plate, well, 150)
            tctrl.append({"id_code": "%s_%d_%s" % (exp, plate, well), "experiment": exp, "plate": plate, "well": well,
                          "sirna": 1108, "well_type": kind})
    pd.DataFrame(trows).to_csv(os.path.join(root, "data", "metadata", "test.csv"), index=False)
    pd.DataFrame(tctrl).to_csv(os.path.join(root, "data", "metadata", "test_controls.csv"), index=False)
    # main.py:157-166 derives the plate groups from the training metadata: every siRNA on exactly three plates
    full = [{"sirna": s, "plate": p} for s in range(1108) for p in range(1, 5) if p != s % 4 + 1]
    pd.DataFrame(full).to_csv(os.path.join(root, "data", "full_metadata", "train.csv"), index=False)
    stats = {e: {"mean": np.linspace(0.1, 0.2, 6), "std": np.linspace(0.05, 0.1, 6)} for e in ("HEPG2-01", "HEPG2-08", "U2OS-04")}
    with open(os.path.join(root, "stats_experiments.pickle"), "wb") as f:
        pickle.dump(stats, f)
    return trows


@pytest.mark.skipif(not os.path.exists(REF_MAIN), reason="the reference tree is only in the build container")
@pytest.mark.timeout(600)
def test_reference_main_py_runs_unchanged_on_this_package(tmp_path):
    import pandas as pd
    import torch
    if torch.cuda.is_available():
        pytest.skip("main.py's local (CPU) branch is what this test drives")
    root = str(tmp_path)
    trows = _write_world(root)
    pkg = os.path.join(root, "cell_classifier")
    os.makedirs(pkg)
    for name, body in ALIAS.items():
        with open(os.path.join(pkg, name + ".py"), "w") as f:
            f.write(body)
    with open(os.path.join(root, "standins.py"), "w") as f:
        f.write(STANDINS % {"root": ROOT})
    driver = ("import sys, runpy; sys.path.insert(0, %r); sys.path.insert(0, %r); import standins; "
              "sys.argv = ['main.py', '--debug', '--experiment_id', 'dry']; runpy.run_path(%r, run_name='__main__'); "
              "print('TRAIN_LOSSES', standins.LOG)" % (ROOT, root, REF_MAIN))
    p = subprocess.run([sys.executable, "-c", driver], cwd=root, capture_output=True, text=True)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "########## TRAINING ##########" in p.stdout and "########## TEST ##########" in p.stdout
    # train(): 5 epochs of one 2-sample batch (main.py's local hyper-parameters), a checkpoint in the reference's format
    losses = eval(p.stdout.split("TRAIN_LOSSES", 1)[1].strip())
    assert len(losses) == 5 and losses[-1] < losses[0]
    sd = torch.load(os.path.join(root, "models", "best_model_dry.pth"))
    assert all(k.startswith("module.") for k in sd) and "module.classifier.bias" in sd
    assert p.stdout.count("Validation Results") == 6
    # test(): one class id per test well, on the well's plate group; the submission the reference writes
    sub = pd.read_csv(os.path.join(root, "submission_dry.csv"))
    assert list(sub.columns) == ["id_code", "sirna"] and list(sub.id_code) == [r["id_code"] for r in trows]
    experiment_types = [3, 1]                                     # main.py:168, first two entries
    for i, r in enumerate(trows):
        s, et = int(sub.sirna[i]), experiment_types[i // 3]
        plates_of_s = [q for q in range(1, 5) if q != s % 4 + 1]
        group = plates_of_s + [10 - sum(plates_of_s)]             # main.py:163-165
        assert group[et] == r["plate"] or s == 0
    assert len(set(sub.sirna[:3])) == 3 and len(set(sub.sirna[3:])) == 3      # one class per well within an experiment


@pytest.mark.skipif(not os.path.exists("/root/reference/compute_stats_experiments.py"),
                    reason="the reference tree is only in the build container")
@pytest.mark.timeout(600)
def test_reference_stats_script_and_ours_write_the_same_pickle(tmp_path):
    """The reference's compute_stats_experiments.py run as the script it is (it globs data/, writes
    stats_experiments.pickle, prints its verification pass) against this package's module run the same way, on one
    synthetic tree: same experiments, same float64 means and standard deviations (kernels replaced by stand-ins)."""
    import pickle
    import cv2
    from recursion_cellular_image_classification_b200.synth import synth_planes
    trees = {}
    for who in ("reference", "ours"):
        root = tmp_path / who
        for split, exps in (("train", ("HEPG2-01", "RPE-03")), ("test", ("HUVEC-17",))):
            for ei, exp in enumerate(exps):
                d = root / "data" / split / exp / "Plate1"
                d.mkdir(parents=True)
                planes = synth_planes(len(exp) + ei, n=2, H=512, W=512)
                for site in (1, 2):
                    for ch in range(6):
                        (d / ("B02_s%d_w%d.jpeg" % (site, ch + 1))).write_bytes(
                            cv2.imencode(".png", planes[site - 1, ch])[1].tobytes())     # lossless bytes, .jpeg name
        trees[who] = str(root)
    r = subprocess.run([sys.executable, "/root/reference/compute_stats_experiments.py"], cwd=trees["reference"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    code = textwrap.dedent('''
        import sys
        sys.path.insert(0, %r)
        sys.path.insert(0, %r)
        import torch
        from test_parallel_cpu import _fake_stat_kernels
        from recursion_cellular_image_classification_b200 import compute_stats_experiments as cse
        _fake_stat_kernels()
        cse.main(device="cpu")
    ''') % (ROOT, os.path.join(ROOT, "tests"))
    o = subprocess.run([sys.executable, "-c", code], cwd=trees["ours"], capture_output=True, text=True)
    assert o.returncode == 0, o.stderr[-2000:]
    with open(os.path.join(trees["reference"], "stats_experiments.pickle"), "rb") as f:
        ref = pickle.load(f)
    with open(os.path.join(trees["ours"], "stats_experiments.pickle"), "rb") as f:
        ours = pickle.load(f)
    assert set(ref) == set(ours) == {"HEPG2-01", "RPE-03", "HUVEC-17"}
    for e in ref:
        assert ours[e]["mean"].dtype == np.float64 and ours[e]["mean"].shape == (6,)
        np.testing.assert_allclose(ours[e]["mean"], ref[e]["mean"], rtol=1e-12)
        np.testing.assert_allclose(ours[e]["std"], ref[e]["std"], rtol=1e-10)
    assert "Verification:" in r.stdout and "Verification:" in o.stdout       # both print the mean ~ 0 / std = 1 pass
