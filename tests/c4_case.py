"""The seeded BASELINE-config-4 case shared by the CPU and GPU tests (test infrastructure): a tiny RxRx1-shaped test
experiment — two plates x three sample wells x two sites, plus the B02 / C03 control wells — whose images differ in
spatial spectrum (D4-symmetric patterns, so the eight views agree), a torchvision DenseNet-121 with the reference's
stem and non-trivial running statistics, and a classifier whose six designated rows are solved from the fp32 features
so that the oracle's logits are WELL SEPARATED: every well has a clear favourite class, two pairs of wells collide on a
favourite (the greedy loop of test.py:48-56 must send the weaker well to its second choice)."""
import os

import numpy as np

S = 64
WELLS = ["D04", "E05", "F06", "G07", "H08", "I09"]
PLATES = [1, 1, 1, 2, 2, 2]
EXP = "HUVEC-18"
EXPERIMENT_TYPE = 1
MEAN, STD = np.full(6, 0.3), np.full(6, 0.25)


def well_planes(w, site):
    rng = np.random.default_rng(77 * w + site)
    yy, xx = np.mgrid[0:S, 0:S]
    cy, cx = (yy + 0.5) / S - 0.5, (xx + 0.5) / S - 0.5
    r = np.sqrt(cy * cy + cx * cx) / 0.7071
    pats = [rng.random((S, S)) * 255, 255 * np.exp(-8 * r * r), 255.0 * ((yy + xx) % 2),
            255.0 * (((yy // 8) + (xx // 8)) % 2), 127 * (1 + np.cos(10 * np.pi * r)), 40 + 0 * r,
            200 * r, 90 + 0 * r]
    img = pats[w][None] * np.ones((6, 1, 1)) + rng.normal(0, 3 + site, (6, S, S))
    return np.clip(img, 0, 255).astype(np.uint8)


def write_tree(root):
    """Files (lossless PNG bytes under the reference's .jpeg names) + the two data frames ImagesDS takes."""
    import cv2
    import pandas as pd
    rows, ctrl, planes = [], [], {}
    for plate in (1, 2):
        d = os.path.join(root, "test", EXP, "Plate%d" % plate)
        os.makedirs(d, exist_ok=True)
        wells = [(w, WELLS[w]) for w in range(6) if PLATES[w] == plate] + [(6, "B02"), (7, "C03")]
        for w, well in wells:
            for site in (1, 2):
                p = well_planes(w, site)
                planes[(plate, well, site)] = p
                for ch in range(6):
                    with open(os.path.join(d, "%s_s%d_w%d.jpeg" % (well, site, ch + 1)), "wb") as f:
                        f.write(cv2.imencode(".png", p[ch])[1].tobytes())
            rec = {"id_code": "%s_%d_%s" % (EXP, plate, well), "experiment": EXP, "plate": plate, "well": well}
            if well == "B02":
                ctrl.append(dict(rec, well_type="negative_control", sirna=1108))
            elif well == "C03":
                ctrl.append(dict(rec, well_type="positive_control", sirna=1109))
            else:
                rows.append(rec)
    return pd.DataFrame(rows), pd.DataFrame(ctrl), planes


def view_codes(n_views):
    from recursion_cellular_image_classification_b200 import ops
    return [ops.aug_code(v, False, k) for v in (False, True) for k in range(4)][:n_views]


def build_oracle_model(planes, plate_groups):
    """(fp32 torchvision net in eval mode, the six designated classes)."""
    import torch
    from oracle import oracle_np as O
    net = O.densenet121_6ch(1108, seed=7).float()
    g = torch.Generator().manual_seed(11)
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.1, generator=g)
                m.running_var.uniform_(0.5, 1.5, generator=g)
    net.eval()
    feats = []
    with torch.no_grad():
        for c in view_codes(8):
            x = np.stack([O.transform(planes[(PLATES[w], WELLS[w], s)], MEAN, STD, vflip=bool(c & 1), hflip=bool(c & 2),
                                      k=(c >> 2) & 3) for w in range(6) for s in (1, 2)])
            f = torch.relu(net.features(torch.from_numpy(x))).mean((2, 3))
            feats.append(f.view(6, 2, -1).mean(1).double())
    F = torch.stack(feats).mean(0)                                     # [6 wells, 1024]
    col = plate_groups[:, EXPERIMENT_TYPE]
    p1, p2 = np.flatnonzero(col == 1)[[5, 50, 200]], np.flatnonzero(col == 2)[[7, 70, 170]]
    classes = [int(c) for c in list(p1) + list(p2)]                    # A1 B1 C1 A2 B2 C2
    T = torch.zeros(6, 6, dtype=torch.float64)
    T[0, 0] = 2.0                       # plate 1: well 0 -> A1
    T[1, 0], T[1, 1] = 1.4, 0.8         #          well 1 also prefers A1, falls back to B1
    T[2, 2] = 2.0                       #          well 2 -> C1
    T[3, 3] = 2.0                       # plate 2: well 3 -> A2
    T[4, 4] = 1.6                       #          well 4 -> B2
    T[5, 4], T[5, 5] = 1.0, 0.6         #          well 5 also prefers B2, falls back to C2
    Wsel = (torch.linalg.pinv(F) @ T).T                                 # [6, 1024]
    with torch.no_grad():
        net.classifier.weight.normal_(0, 2e-4, generator=g)
        net.classifier.bias.zero_()
        for i, c in enumerate(classes):
            net.classifier.weight[c] = Wsel[i].float()
    return net, classes


def oracle_logits(net, planes, n_views):
    """[V, 6, 1108] fp32: per view, the mean over a well's two sites of the net's logits (first third of the
    reference's item, models.py:46-49; linear head)."""
    import torch
    from oracle import oracle_np as O
    out = []
    with torch.no_grad():
        for c in view_codes(n_views):
            x = np.stack([O.transform(planes[(PLATES[w], WELLS[w], s)], MEAN, STD, vflip=bool(c & 1), hflip=bool(c & 2),
                                      k=(c >> 2) & 3) for w in range(6) for s in (1, 2)])
            out.append(net(torch.from_numpy(x)).view(6, 2, -1).mean(1).numpy())
    return np.stack(out)


def oracle_assign(logits, plate_groups, plates):
    """(masked + rescaled probabilities f32 [N,C], assignment) from [V,N,C] logits — test.py:27-56 via the oracle."""
    from oracle import oracle_np as O
    probs = np.mean([O.softmax(v) for v in logits], axis=0).astype(np.float32)
    probs = O.mask_rescale(probs, plate_groups[:, EXPERIMENT_TYPE], np.asarray(plates))
    return probs, O.greedy_assign(probs.copy())
