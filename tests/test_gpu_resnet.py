"""GPU parity of the reference's real model served natively: TwoSitesResNet50 — torchvision ResNet-50
trunk, image / negative / positive feature thirds, MLP head (reference cell_classifier/models.py:7-57), evaluation
mode — against the reference's own output (tests/golden/model_golden.npz) and the fp32 oracle restatement.
Tolerance: the north star's 2e-2 relative for bf16 logits."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle_np as O
from recursion_cellular_image_classification_b200._lib import RxbError
from recursion_cellular_image_classification_b200.cell_classifier.models import TwoSitesNN, TwoSitesResNet50

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return (a - b).norm().item() / max(b.norm().item(), 1e-12)


def _randomise_running_stats(ref, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in ref.modules():
            if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.BatchNorm1d)):
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
                m.weight.copy_(torch.rand(m.weight.shape, generator=g) * 0.5 + 0.75)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)


def test_matches_the_references_own_logits(cuda, golden_dir):
    """The reference's TwoSitesNN itself (seed 321, eval mode, x[2,3,6,64,64]) produced full_logits; the native
    executor with the same weights gives them within 2e-2."""
    g = np.load(os.path.join(golden_dir, "model_golden.npz"))
    net = TwoSitesNN(pretrained=False, nb_classes=1108, trunk="resnet50", device=cuda)
    assert isinstance(net, TwoSitesResNet50)
    net.load_state_dict(O.two_sites_resnet50(seed=int(g["full_seed"])).state_dict(), strict=False)
    x = torch.randn(2, 3, 6, 64, 64, generator=torch.Generator().manual_seed(6))
    got = net(x.to(cuda)).cpu()
    want = torch.from_numpy(g["full_logits"])
    print("\nvs reference golden: rel %.4g" % _rel(got, want))
    assert _rel(got, want) < 2e-2
    net.train()
    with pytest.raises(RxbError):                      # no native training of this trunk, and no fallback
        net(x.to(cuda))


@pytest.mark.parametrize("B,G,S", [(2, 6, 128), (1, 3, 364), (3, 3, 96)])
def test_matches_fp32_oracle_with_trained_like_statistics(cuda, B, G, S):
    """Non-trivial running statistics and BatchNorm weights, the reference's two item layouts (G=3 train/val, G=6 test)
    and its training crop 364 (91 -> 46 -> 23 -> 12 pixel maps: odd sizes through the stride-2 convolutions)."""
    ref = O.two_sites_resnet50(seed=9)
    _randomise_running_stats(ref, 10)
    ref.eval()
    net = TwoSitesResNet50(device=cuda)
    net.load_state_dict(ref.state_dict(), strict=False)
    x = torch.randn(B, G, 6, S, S, generator=torch.Generator().manual_seed(B + S)).to(torch.bfloat16).float()
    ref = ref.to(cuda)
    with torch.no_grad():
        want = ref(x.to(cuda))
    got = net(x.to(cuda))
    print("\nB=%d G=%d S=%d: logits rel %.4g" % (B, G, S, _rel(got, want)))
    assert _rel(got, want) < 2e-2
    # the thirds matter: swapping the control images changes the logits (they reach the MLP, models.py:50-53)
    perm = list(range(G // 3, G)) + list(range(G // 3))
    assert _rel(net(x[:, perm].to(cuda)), got) > 1e-3


def test_test_shim_serves_the_reference_model(cuda, tmp_path):
    """test() with the reference's model: items keep their control wells (G = 6), logits come from the concatenated
    thirds, and the assignment equals the oracle's on the same items."""
    import c4_case as C
    from recursion_cellular_image_classification_b200.cell_classifier import dataloader as dl
    from recursion_cellular_image_classification_b200.cell_classifier.test import predict_probs
    from recursion_cellular_image_classification_b200.synth import synth_plate_groups
    pg = synth_plate_groups(3)
    df, dfc, planes = C.write_tree(str(tmp_path))
    ref = O.two_sites_resnet50(seed=4)
    _randomise_running_stats(ref, 5)
    ref.eval()
    net = TwoSitesResNet50(device=cuda)
    net.load_state_dict(ref.state_dict(), strict=False)
    ds = dl.ImagesDS(df, dfc, {C.EXP: {"mean": C.MEAN, "std": C.STD}}, str(tmp_path), "test", verbose=False)
    got = predict_probs(df, ds, pg, C.EXPERIMENT_TYPE, net, bs=2, num_workers=0, device="cuda").cpu().numpy()
    # oracle: the reference's test item (both sites of the well, of B02, of C03), fp32
    rows = []
    with torch.no_grad():
        for w in range(6):
            imgs = [planes[(C.PLATES[w], well, s)] for well in (C.WELLS[w], "B02", "C03") for s in (1, 2)]
            x = np.stack([O.transform(im, C.MEAN, C.STD) for im in imgs])[None]
            rows.append(ref(torch.from_numpy(x)).numpy()[0])
    want, _ = C.oracle_assign(np.stack(rows)[None], pg, df.plate.values)
    rel = np.linalg.norm(got - want) / np.linalg.norm(want)
    print("\nTwoSitesResNet50 through test(): probabilities rel %.4g" % rel)
    assert rel < 2e-2
