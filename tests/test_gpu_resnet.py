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


def _oracle_train_forward(ref, x, m0, m1):
    """models.py:41-57 in training mode with the Dropout draws replaced by the given masks."""
    bs = x.shape[0]
    feats = ref.base_nn(x.reshape([-1, x.shape[2], x.shape[3], x.shape[4]]))
    cat = O.two_sites_features(feats, bs)
    m = ref.mlp
    h1 = m[2](m[0](cat) * m0)
    return m[6](m[4](torch.relu(h1)) * m1)


@pytest.mark.parametrize("B,G,S", [(4, 3, 128), (2, 3, 364)])
def test_train_step_matches_fp32_oracle(cuda, B, G, S):
    """One training step of the reference's model (train.py:37,44): BatchNorm batch statistics in the trunk and the
    head, Dropout with explicit masks, CrossEntropy, full backward — loss and training-mode logits within 2e-2 of
    torch fp32, every parameter gradient as close to fp32 as PyTorch's own bf16 autocast path, running statistics
    updated like torch's."""
    import copy
    ref = O.two_sites_resnet50(seed=11).to(cuda)
    _randomise_running_stats(ref, 12)
    net = TwoSitesResNet50(device=cuda)
    net.load_state_dict(ref.state_dict(), strict=False)
    g = torch.Generator().manual_seed(S + B)
    x = torch.randn(B, G, 6, S, S, generator=g).to(torch.bfloat16).float().to(cuda)
    y = torch.randint(0, 1108, (B,), generator=g).to(cuda)
    keep = 0.7
    m0 = ((torch.rand(B, 6144, generator=g) < keep).float() / keep).to(cuda)
    m1 = ((torch.rand(B, 1024, generator=g) < keep).float() / keep).to(cuda)
    # yardstick: torch bf16 autocast against its own fp32
    cal = copy.deepcopy(ref)
    cal.train()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        lc = torch.nn.CrossEntropyLoss()(_oracle_train_forward(cal, x, m0, m1).float(), y)
    lc.backward()
    ref.train()
    out = _oracle_train_forward(ref, x, m0, m1)
    loss = torch.nn.CrossEntropyLoss()(out, y)
    loss.backward()
    logits = torch.empty(B, 1108, device=cuda)
    net.train()
    my_loss = net.train_step(x, y, masks=(m0, m1), logits_out=logits).item()
    torch.cuda.synchronize()
    rel_logits = _rel(logits, out.detach())
    flat_ref = torch.cat([p.grad.flatten() for _, p in ref.named_parameters()])
    flat_cal = torch.cat([p.grad.flatten() for _, p in cal.named_parameters()])
    cos = torch.nn.functional.cosine_similarity(flat_ref, net.flat.grad, dim=0).item()
    cos_cal = torch.nn.functional.cosine_similarity(flat_ref, flat_cal, dim=0).item()
    worst = []
    for (name, p), (_, pc) in zip(ref.named_parameters(), cal.named_parameters()):
        g_ref, g_my, g_cal = p.grad.flatten(), net.grad_view(name).flatten(), pc.grad.flatten()
        assert torch.isfinite(g_my).all(), name
        e_my, e_cal, n_ref = (g_my - g_ref).norm().item(), (g_cal - g_ref).norm().item(), g_ref.norm().item()
        worst.append((e_my / (2.0 * e_cal + 0.02 * n_ref + 1e-12), name, e_my, e_cal, n_ref))
    worst.sort(reverse=True)
    print("\nResNet-50 TwoSitesNN train step B=%d G=%d S=%d: loss ours %.5f fp32 %.5f | logits rel %.4f | grad cos %.4f "
          "(torch-bf16 %.4f) | worst (our err)/(2 x torch-bf16 err + 2%%): %s" %
          (B, G, S, my_loss, loss.item(), rel_logits, cos, cos_cal, [(round(w[0], 2), w[1]) for w in worst[:4]]))
    assert abs(my_loss - loss.item()) < 2e-2 * abs(loss.item())
    assert rel_logits < 2e-2, rel_logits
    assert cos > cos_cal - 0.03, (cos, cos_cal)
    assert worst[0][0] < 1.0, worst[:5]
    for name, buf in ref.named_buffers():
        if name.endswith("running_mean") or name.endswith("running_var"):
            assert _rel(net.buffer_view(name), buf) < 2e-2, name


def test_sgd_steps_learn_and_head_only_freezes_the_trunk(cuda):
    B, G, S = 4, 3, 64
    net = TwoSitesResNet50(device=cuda, seed=3)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, G, 6, S, S, generator=g).to(cuda)
    y = torch.randint(0, 1108, (B,), generator=g).to(cuda)
    masks = tuple(torch.ones_like(m) for m in net.dropout_masks(B))
    net.train()
    trunk_before = net.flat.data[:net.head_range()[0]].clone()
    losses = []
    for i in range(6):
        losses.append(net.train_step(x, y, masks=masks).item())
        net.sgd_step(B, G, S, S, lr=0.01, head_only=i < 2)
        if i == 1:
            assert torch.equal(net.flat.data[:net.head_range()[0]], trunk_before)      # train.py:46-58: trunk frozen
    assert not torch.equal(net.flat.data[:net.head_range()[0]], trunk_before)
    assert losses[-1] < losses[0], losses
    net.eval()
    assert torch.isfinite(net(x)).all()
