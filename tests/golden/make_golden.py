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
  model_golden.npz   cell_classifier.models.TwoSitesNN (reference models.py:8-57) built with pretrained=False under a
                     seed: the 6-channel stem it derives from the 3-channel kernel (:17-27), and its forward's
                     site/control averaging (:41-53) run with a stub trunk and an identity head on seeded input.
  jpeg_golden.npz    the reference's own PNG->JPEG converter (png_to_jpeg.convert_png_to_jpeg, :11-15: PIL 'L', quality 95)
                     run on seeded synthetic planes, and cv2.imdecode(buf, -1) of the result — the reference's decode
                     call (cell_classifier/dataloader.py:141-146): file bytes + decoded planes.
  warp_golden.npz    cv2.getRotationMatrix2D + cv2.warpAffine(INTER_LINEAR, BORDER_REFLECT_101) — the call
                     albumentations 0.3.0 ShiftScaleRotate makes for dataloader.py:45-46 (albumentations itself is
                     not installed) — on seeded 6-channel u8 images: full outputs at 48x48, SHA-256 digests at 512x512.
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


WARP_ANGLES = [-180.0, -137.3, -90.0, -45.0, -0.37, 12.5, 90.0, 151.9]


def make_warp():
    import hashlib
    import cv2

    def warp(img, ang):
        h, w = img.shape[:2]
        M = cv2.getRotationMatrix2D((w / 2, h / 2), ang, 1.0)
        return M, cv2.warpAffine(img, M, (w, h), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)

    small = np.random.default_rng(7).integers(0, 256, size=(48, 48, 6), dtype=np.uint8)
    big = np.ascontiguousarray(np.moveaxis(synth_planes(5, n=1)[0], 0, 2))      # [512,512,6]
    mats, outs, digests = [], [], []
    for ang in WARP_ANGLES:
        M, o = warp(small, ang)
        mats.append(M)
        outs.append(o)
        digests.append(hashlib.sha256(warp(big, ang)[1].tobytes()).hexdigest())
    np.savez_compressed(os.path.join(HERE, "warp_golden.npz"), angles=np.array(WARP_ANGLES), seed_small=7,
                        seed_big=5, mats=np.array(mats), out_small=np.array(outs), sha256_big=np.array(digests),
                        cv2_version=cv2.__version__)
    print("warp golden:", cv2.__version__, digests[1][:16])


def make_model():
    import torch
    sys.path.insert(0, REF)
    from cell_classifier.models import TwoSitesNN
    from torchvision import models
    torch.manual_seed(123)
    net = TwoSitesNN(pretrained=False, nb_classes=1108)
    torch.manual_seed(123)
    rgb = models.resnet50(weights=None).conv1.weight.detach().numpy()        # the kernel the surgery started from
    stem6 = net.base_nn.conv1.weight.detach().numpy()

    class Trunk(torch.nn.Module):                                            # features = per-image channel means
        def forward(self, x):
            return x.mean(dim=(2, 3)) * torch.arange(1, 7, dtype=x.dtype)

    net.base_nn = Trunk()
    net.mlp = torch.nn.Identity()
    net.eval()
    g = torch.Generator().manual_seed(5)
    outs = {}
    for G in (3, 6):                                                         # train/val: 3 images, test: 6 (two sites)
        x = torch.randn(4, G, 6, 8, 8, generator=g)
        outs["x%d" % G] = x.numpy()
        outs["y%d" % G] = net(x).detach().numpy()
    # the whole reference model (ResNet-50 trunk + MLP head) in eval mode on seeded input
    torch.manual_seed(321)
    full = TwoSitesNN(pretrained=False, nb_classes=1108)
    full.eval()
    xf = torch.randn(2, 3, 6, 64, 64, generator=torch.Generator().manual_seed(6))
    with torch.no_grad():
        outs["full_logits"] = full(xf).numpy()
    outs["full_seed"] = 321
    np.savez_compressed(os.path.join(HERE, "model_golden.npz"), rgb_sample=rgb[:4], stem6_sample=stem6[:4],
                        stem6_sum=np.float64(stem6.astype(np.float64).sum()), **outs)
    print("model golden:", stem6.shape, outs["y3"].shape, outs["y6"].shape)


def make_jpeg():
    import cv2
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp()
    os.chdir(tmp)       # png_to_jpeg.py is a script: it globs data/**/*.png in the CWD (nothing there) when imported
    try:
        sys.path.insert(0, REF)
        import png_to_jpeg
        files, planes, shapes = [], [], []
        for i, (seed, H, W) in enumerate([(21, 64, 64), (22, 64, 64), (23, 40, 72), (24, 256, 256)]):
            img = synth_planes(seed, n=1, C=1, H=H, W=W)[0, 0]
            path = os.path.join(tmp, "w%d.png" % i)
            cv2.imwrite(path, img)
            png_to_jpeg.convert_png_to_jpeg(path)
            buf = open(os.path.join(tmp, "w%d.jpeg" % i), "rb").read()
            dec = cv2.imdecode(np.frombuffer(buf, dtype=np.uint8), -1)
            assert dec.shape == (H, W) and dec.dtype == np.uint8
            files.append(np.frombuffer(buf, dtype=np.uint8))
            planes.append(dec.reshape(-1))
            shapes.append((H, W))
        np.savez_compressed(os.path.join(HERE, "jpeg_golden.npz"), shapes=np.array(shapes),
                            file_sizes=np.array([len(f) for f in files]), files=np.concatenate(files),
                            planes=np.concatenate(planes), cv2_version=cv2.__version__)
        print("jpeg golden:", [len(f) for f in files])
    finally:
        os.chdir(cwd)


if __name__ == "__main__":
    if "--model-only" in sys.argv:
        make_model()
        sys.exit(0)
    if "--jpeg-only" in sys.argv:
        make_jpeg()
        sys.exit(0)
    make_model()
    make_jpeg()
    make_warp()
    if "--warp-only" in sys.argv:
        sys.exit(0)
    make_stats()
    make_assign()
