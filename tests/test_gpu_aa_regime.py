"""GPU parity of the native DenseNet-121 executor IN THE REGIME THE BENCHMARK RUNS IN: 512x512
inputs, a batch large enough that every persistent conv CTA walks many tiles (so pipeline stages, TMEM accumulator
stages and the in-place staging buffers all wrap), and a multi-step trajectory.  Oracle: torchvision densenet121 with
the reference's stem in genuine fp32 (TF32 off, tests/conftest.py).  Tolerances are the north star's: 2e-2 relative
for bf16 logits and loss, written below without relaxation.  Collected before the other GPU files on purpose."""
import numpy as np
import pytest
import torch

from oracle import oracle_np as O
from recursion_cellular_image_classification_b200.cell_classifier.models import DenseNet121

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return (a - b).norm().item() / max(b.norm().item(), 1e-12)


def _pair(cuda, seed):
    ref = O.densenet121_6ch(num_classes=1108, seed=seed).to(cuda).float()
    net = DenseNet121(nb_classes=1108, device=cuda)
    net.load_state_dict(ref.state_dict(), strict=True)
    return ref, net


def test_train_step_512_batch16_matches_fp32_oracle(cuda):
    """One full training step at the benchmark's image size, 16 images (conv CTAs walk up to ~14 tiles in dense
    block 1, ~55 in the stem): loss, training-mode logits (batch statistics), evaluation logits after the step's
    running-statistics update, per-tensor gradients."""
    B, S = 16, 512
    ref, net = _pair(cuda, seed=0)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, 6, S, S, generator=g).to(torch.bfloat16).float().to(cuda)
    y = torch.randint(0, 1108, (B,), generator=g).to(cuda)
    ref.train()
    net.train()
    # yardstick for the GRADIENT only: PyTorch's own bf16 autocast against its fp32 on the same weights and input
    # (bf16 through 121 BatchNorm'd layers at random initialisation is noisy; there is no stated tolerance for gradients)
    import copy
    cal = copy.deepcopy(ref)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        torch.nn.CrossEntropyLoss()(cal(x).float(), y).backward()
    cal_grad = torch.cat([p.grad.flatten() for _, p in cal.named_parameters()])
    del cal
    out = ref(x)
    loss = torch.nn.CrossEntropyLoss()(out, y)
    loss.backward()
    my_loss = net.train_step(x, y).item()
    torch.cuda.synchronize()
    # running statistics after exactly one training step (momentum 0.1, unbiased variance), before anything else runs
    for name, buf in ref.named_buffers():
        if name.endswith("running_mean") or name.endswith("running_var"):
            assert _rel(net.buffer_view(name), buf) < 2e-2, name
    got_train = net(x)                                   # training-mode forward: batch statistics again ...
    with torch.no_grad():
        ref(x)                                           # ... on both sides, so the running statistics stay in step
    rel_train = _rel(got_train, out.detach())
    flat_ref = torch.cat([p.grad.flatten() for _, p in ref.named_parameters()])
    cos = torch.nn.functional.cosine_similarity(flat_ref, net.flat.grad, dim=0).item()
    cos_cal = torch.nn.functional.cosine_similarity(flat_ref, cal_grad, dim=0).item()
    rel_grad = _rel(net.flat.grad, flat_ref)
    ref.eval()
    net.eval()
    with torch.no_grad():
        want_eval = ref(x)
    rel_eval = _rel(net(x), want_eval)
    print("\n512x512 B=16: loss ours %.5f fp32 %.5f | train logits rel %.4f | eval logits rel %.4f | grad cos %.4f "
          "(torch-bf16 %.4f) rel %.4f (torch-bf16 %.4f)"
          % (my_loss, loss.item(), rel_train, rel_eval, cos, cos_cal, rel_grad, _rel(cal_grad, flat_ref)))
    assert abs(my_loss - loss.item()) < 2e-2 * abs(loss.item())
    assert rel_eval < 2e-2, rel_eval
    assert rel_train < 2e-2, rel_train
    assert cos > cos_cal - 0.03, (cos, cos_cal)         # as close to the fp32 gradient as PyTorch's bf16 path is


def test_twenty_step_loss_trajectory_matches_fp32_oracle(cuda):
    """20 SGD steps (the reference's optimizer: momentum .9, nesterov, wd 3e-5 — main.py:89-93) on
    one fixed batch, natively in bf16 and with torch fp32: the loss curves stay within 2e-2 relative of each other at
    every step while the loss falls from 7.0 to below 60 % of that.  (The learning rate keeps the 20 steps in the
    descent.  The bf16 run trails the fp32 one slightly and the gap grows with the distance travelled: measured on the
    B200, lr 0.003 ends at 28 % of the initial loss with 2.7 % between the curves, lr 0.01 memorises the 16 images -
    loss 0.2 - with 10 %; gpurun logs of round 2.)"""
    B, S, lr = 16, 256, 0.002
    ref, net = _pair(cuda, seed=2)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, 6, S, S, generator=g).to(torch.bfloat16).float().to(cuda)
    y = torch.randint(0, 1108, (B,), generator=g).to(cuda)
    opt = O.sgd_reference(ref.parameters(), lr=lr)
    lossf = torch.nn.CrossEntropyLoss()
    ref.train()
    net.train()
    mine, theirs = [], []
    for _ in range(20):
        opt.zero_grad()
        l = lossf(ref(x), y)
        l.backward()
        opt.step()
        theirs.append(l.item())
        mine.append(net.train_step(x, y).item())
        net.sgd_step(B, S, S, lr=lr, momentum=0.9, weight_decay=3e-5, nesterov=True)
    mine, theirs = np.array(mine), np.array(theirs)
    dev = np.abs(mine - theirs) / theirs
    print("\nloss trajectory ours  :", np.round(mine, 4).tolist())
    print("loss trajectory fp32  :", np.round(theirs, 4).tolist())
    print("max relative deviation %.4f at step %d" % (dev.max(), int(dev.argmax())))
    assert theirs[-1] < 0.6 * theirs[0], theirs                 # the run actually trains
    assert dev.max() < 2e-2, dev.tolist()
    # the weights themselves after 20 steps
    flat_ref = torch.cat([p.detach().flatten() for _, p in ref.named_parameters()])
    assert _rel(net.flat.detach(), flat_ref) < 2e-2


def test_bn_backward_with_small_and_negative_gammas(cuda):
    """BatchNorm weights as a trained / weight-decayed checkpoint has them — tiny, zero and negative gammas, non-zero
    betas: the parameter gradients of every BatchNorm and of the convolutions behind them still follow the
    fp32 oracle.  (The W.dW identity that recovers sum(dy*x) divides by gamma*rstd; degenerate channels take a direct
    reduction in the data-gradient epilogue instead.)"""
    B, S = 8, 128
    ref, net = _pair(cuda, seed=4)
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for name, m in ref.named_modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                w = torch.rand(m.weight.shape, generator=g) * 1.0 + 0.5
                pick = torch.rand(m.weight.shape, generator=g)
                w[pick < 0.10] *= 1e-3                            # nearly dead channels
                w[(pick >= 0.10) & (pick < 0.13)] = 0.0           # dead channels
                w[(pick >= 0.13) & (pick < 0.25)] *= -1.0         # negative gammas
                m.weight.copy_(w.to(m.weight.device))
                m.bias.copy_((torch.randn(m.bias.shape, generator=g) * 0.2).to(m.bias.device))
    net.load_state_dict(ref.state_dict(), strict=True)
    x = torch.randn(B, 6, S, S, generator=g).to(torch.bfloat16).float().to(cuda)
    y = torch.randint(0, 1108, (B,), generator=g).to(cuda)
    ref.train()
    net.train()
    loss = torch.nn.CrossEntropyLoss()(ref(x), y)
    loss.backward()
    my_loss = net.train_step(x, y).item()
    torch.cuda.synchronize()
    assert abs(my_loss - loss.item()) < 2e-2 * abs(loss.item()), (my_loss, loss.item())
    # same-input yardstick: what PyTorch's own bf16 autocast does against its fp32 on these weights
    import copy
    cal = copy.deepcopy(ref)
    cal.zero_grad()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        lc = torch.nn.CrossEntropyLoss()(cal(x).float(), y)
    lc.backward()
    worst, noise = [], []
    for (name, p), (_, pc) in zip(ref.named_parameters(), cal.named_parameters()):
        if "norm" not in name:
            continue
        if name.startswith("features.norm0."):
            # the stem BatchNorm feeds ReLU / max-pool and then only BatchNorms: its true gradient is ~0 by scale
            # invariance (|g_fp32| 0.02 here) and both bf16 paths carry rounding noise of the 16.7 M-term sums instead -
            # reported, not gated (tests/test_gpu_densenet.py treats it the same way)
            noise.append((name, (net.grad_view(name).flatten() - p.grad.flatten()).norm().item(),
                          (pc.grad.flatten() - p.grad.flatten()).norm().item(), p.grad.norm().item()))
            continue
        g_ref, g_my, g_cal = p.grad.flatten(), net.grad_view(name).flatten(), pc.grad.flatten()
        assert torch.isfinite(g_my).all(), name
        e_my, e_cal, n_ref = (g_my - g_ref).norm().item(), (g_cal - g_ref).norm().item(), g_ref.norm().item()
        worst.append((e_my / (2.0 * e_cal + 0.02 * n_ref + 1e-12), name, e_my, e_cal, n_ref))
    worst.sort(reverse=True)
    print("\nnoise-only (name, our error, torch-bf16 error, |g_fp32|):", noise)
    print("BatchNorm gradients, degenerate gammas: (our error)/(2 x torch-bf16 error + 2%) worst:",
          [(round(w[0], 3), w[1]) for w in worst[:5]])
    assert worst[0][0] < 1.0, worst[:5]


def test_repeated_runs_forward_bit_identical_training_noise_bounded(cuda):
    """Run-to-run behaviour at 512x512, 16 images.
    Evaluation-mode forward has no cross-CTA accumulation (running statistics), so repeated runs must be BIT-IDENTICAL:
    a stale, recycled or raced tile in the stem, the 1x1 / x-merged 3x3 forward kernels, the transitions or the head
    would break that (companion of the bit-exact repeat-launch stress of the data-gradient kernel, test_gpu_conv.py).
    A training step is not bit-reproducible: the per-channel BatchNorm sums and the weight gradients are accumulated
    with fp32 atomics / L2 reduce-adds whose order varies (~1e-7), and a randomly initialised 121-layer network on noise
    images amplifies that chaotically through the bf16 roundings - run to run the loss moves by ~1e-4 and gradient
    tensors deep in the backward pass by tens of percent, the same size as the bf16-vs-fp32 gradient error of ANY bf16
    path here (PyTorch autocast: 55 % norm-wise, cosine 0.85; tools/repeat_step_diag.py prints the per-tensor table).
    What must hold: the loss agrees to 1e-3, and the gradient of the classifier bias - one GEMM away from the loss, not
    amplified - to 1e-3."""
    B, S = 16, 512
    _, net = _pair(cuda, seed=6)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(B, 6, S, S, generator=g).to(torch.bfloat16).float().to(cuda)
    y = torch.randint(0, 1108, (B,), generator=g).to(cuda)
    with torch.no_grad():
        net.bn_buffers[:] = torch.rand(net.bn_buffers.shape, generator=g).to(cuda) * 0.5 + 0.75    # non-trivial statistics
    net.eval()
    first = net(x).clone()
    assert torch.isfinite(first).all()
    for _ in range(4):
        assert torch.equal(net(x), first)
    net.train()
    buffers0 = net.bn_buffers.clone()
    runs = []
    for _ in range(3):
        net.bn_buffers.copy_(buffers0)
        loss = net.train_step(x, y).item()
        runs.append((loss, net.grad_view("classifier.bias").clone()))
    for loss, gb in runs[1:]:
        assert abs(loss - runs[0][0]) < 1e-3 * abs(runs[0][0])
        assert (gb - runs[0][1]).norm().item() < 1e-3 * runs[0][1].norm().item()
