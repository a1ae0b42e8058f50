"""GPU parity of the native DenseNet-121 executor against torchvision's densenet121 (fp32) with the
reference's 6-channel stem: logits and loss within 2e-2 relative (north star), gradients compared by
direction and norm."""
import numpy as np
import pytest
import torch

from oracle import oracle_np as O
from recursion_cellular_image_classification_b200.cell_classifier.models import DenseNet121

pytestmark = pytest.mark.gpu


def _setup(cuda, B, S, seed=0):
    ref = O.densenet121_6ch(num_classes=1108, seed=seed).to(cuda).float()
    net = DenseNet121(nb_classes=1108, device=cuda)
    net.load_state_dict(ref.state_dict(), strict=True)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(B, 6, S, S, generator=g).to(cuda)
    x = x.to(torch.bfloat16).float()          # both sides see the same bf16-representable input
    y = torch.randint(0, 1108, (B,), generator=g).to(cuda)
    return ref, net, x, y


def _rel(a, b):
    return (a - b).norm().item() / max(b.norm().item(), 1e-12)


def test_state_dict_roundtrip(cuda):
    ref = O.densenet121_6ch(num_classes=1108, seed=3)
    net = DenseNet121(nb_classes=1108, device=cuda)
    net.load_state_dict(ref.state_dict())
    sd = net.state_dict()
    for k, v in ref.state_dict().items():
        if k.endswith("num_batches_tracked"):
            continue
        assert torch.equal(sd[k].cpu(), v), k


@pytest.mark.parametrize("B,S", [(4, 128), (2, 64)])
def test_forward_eval_matches_torchvision(cuda, B, S):
    ref, net, x, _ = _setup(cuda, B, S)
    # non-trivial running statistics
    with torch.no_grad():
        for m in ref.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.1)
                m.running_var.uniform_(0.5, 1.5)
    net.load_state_dict(ref.state_dict())
    ref.eval()
    net.eval()
    with torch.no_grad():
        want = ref(x)
    got = net(x)
    assert _rel(got, want) < 2e-2, _rel(got, want)


def test_train_step_matches_torchvision(cuda):
    B, S = 4, 128
    ref, net, x, y = _setup(cuda, B, S)
    ref.train()
    net.train()
    out = ref(x)
    loss = torch.nn.CrossEntropyLoss()(out, y)
    loss.backward()
    my_loss = net.train_step(x, y)
    torch.cuda.synchronize()
    assert abs(my_loss.item() - loss.item()) < 2e-2 * abs(loss.item()), (my_loss.item(), loss.item())
    # logits of the training-mode forward (batch statistics)
    net.train()
    got = net(x)
    assert _rel(got, out.detach()) < 2e-2, _rel(got, out.detach())
    # gradients: direction and norm per parameter tensor
    worst = (1.0, None)
    for name, p in ref.named_parameters():
        g_ref = p.grad.flatten()
        g_my = net.grad_view(name).flatten()
        if g_ref.norm().item() < 1e-8:
            continue
        cos = torch.nn.functional.cosine_similarity(g_ref, g_my, dim=0).item()
        ratio = g_my.norm().item() / g_ref.norm().item()
        if cos < worst[0]:
            worst = (cos, name)
        assert cos > 0.95, (name, cos, ratio)
        assert 0.8 < ratio < 1.25, (name, cos, ratio)
    # the whole gradient vector
    flat_ref = torch.cat([p.grad.flatten() for _, p in ref.named_parameters()])
    cos_all = torch.nn.functional.cosine_similarity(flat_ref, net.flat.grad, dim=0).item()
    assert cos_all > 0.99, (cos_all, worst)
    # running statistics were updated like torch's
    for name, buf in ref.named_buffers():
        if name.endswith("running_mean") or name.endswith("running_var"):
            got_b = net.buffer_view(name)
            assert _rel(got_b, buf) < 3e-2, (name, _rel(got_b, buf))


def test_sgd_step_and_loss_decreases(cuda):
    B, S = 4, 64
    _, net, x, y = _setup(cuda, B, S, seed=5)
    net.train()
    losses = []
    for _ in range(6):
        l = net.train_step(x, y)
        net.sgd_step(B, S, S, lr=0.02, momentum=0.9, weight_decay=3e-5, nesterov=True)
        losses.append(l.item())
    assert losses[-1] < losses[0], losses
