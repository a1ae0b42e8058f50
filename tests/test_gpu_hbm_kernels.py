"""GPU parity tests (through the C ABI) for the HBM-bound kernel families against the oracle and the
golden vectors: statistics, loader, TTA/mask/assignment, CE, SGD."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle_np as O
from recursion_cellular_image_classification_b200 import ops
from recursion_cellular_image_classification_b200.synth import synth_logits, synth_plate_groups, synth_planes

pytestmark = pytest.mark.gpu


# ---------------------------------------------------------------- statistics
def test_stats_matches_reference_golden(cuda, golden_dir):
    g = np.load(os.path.join(golden_dir, "stats_golden.npz"))
    n = int(g["n_per_exp"])
    planes = np.concatenate([synth_planes(int(s), n=n) for s in g["seeds"]])
    exp_id = np.repeat(np.arange(len(g["seeds"]), dtype=np.int32), n)
    acc = ops.stats_accumulate(torch.from_numpy(planes).to(cuda), torch.from_numpy(exp_id).to(cuda), len(g["seeds"]))
    mean, std = ops.stats_finalize(acc)
    np.testing.assert_allclose(mean.cpu().numpy(), g["mean"], rtol=1e-5)   # north star: 1e-5 relative
    np.testing.assert_allclose(std.cpu().numpy(), g["std"], rtol=1e-5)
    # far tighter in practice (exact integer sums)
    np.testing.assert_allclose(mean.cpu().numpy(), g["mean"], rtol=1e-12)
    np.testing.assert_allclose(std.cpu().numpy(), g["std"], rtol=1e-10)
    vm, vs = ops.stats_finalize(acc, mean, std)   # verification pass: mean ~ 0, std = 1
    np.testing.assert_allclose(vm.cpu().numpy(), g["vmean"], atol=1e-9)
    np.testing.assert_allclose(vs.cpu().numpy(), g["vstd"], rtol=1e-9)


@pytest.mark.parametrize("hw", [(16, 16), (48, 64), (512, 512)])
def test_stats_exact_integer_sums(cuda, hw):
    rng = np.random.default_rng(3)
    n, C = 5, 6
    planes = rng.integers(0, 256, size=(n, C) + hw, dtype=np.uint8)
    planes[0] = 255   # maximum values
    planes[1] = 0
    exp_id = np.array([0, 1, 0, 2, 1], dtype=np.int32)
    s, q, cnt = ops.stats_accumulate(torch.from_numpy(planes).to(cuda), torch.from_numpy(exp_id).to(cuda), 3)
    p64 = planes.astype(np.int64)
    for e in range(3):
        sel = p64[exp_id == e]
        np.testing.assert_array_equal(s[e].cpu().numpy(), sel.sum(axis=(0, 2, 3)))
        np.testing.assert_array_equal(q[e].cpu().numpy(), (sel ** 2).sum(axis=(0, 2, 3)))
        np.testing.assert_array_equal(cnt[e].cpu().numpy(), np.full(C, sel.shape[0] * hw[0] * hw[1]))


def test_stats_chunked_accumulation_and_empty(cuda):
    planes = synth_planes(5, n=4, H=64, W=64)
    t = torch.from_numpy(planes).to(cuda)
    e = torch.zeros(4, dtype=torch.int32, device=cuda)
    whole = ops.stats_accumulate(t, e, 1)
    acc = ops.stats_accumulate(t[:1], e[:1], 1)
    acc = ops.stats_accumulate(t[1:], e[1:], 1, acc)
    acc = ops.stats_accumulate(t[:0], e[:0], 1, acc)   # empty chunk is a no-op
    for a, b in zip(whole, acc):
        assert torch.equal(a, b)
    m, s = ops.stats_finalize(acc)
    om, os_ = O.compute_mean_std_arrays(planes)
    np.testing.assert_allclose(m.cpu().numpy()[0], om, rtol=1e-12)
    np.testing.assert_allclose(s.cpu().numpy()[0], os_, rtol=1e-10)


# ---------------------------------------------------------------- loader
def _loader_case(cuda, S, out_hw, codes, crops, fmt):
    rng = np.random.default_rng(7)
    n_src, n_exp = 3, 2
    src = rng.integers(0, 256, size=(n_src, 6, S, S), dtype=np.uint8)
    mean = rng.random((n_exp, 6)) * 0.2 + 0.05
    std = rng.random((n_exp, 6)) * 0.1 + 0.05
    B = len(codes)
    src_idx = np.array([i % n_src for i in range(B)], dtype=np.int32)
    exp_id = np.array([i % n_exp for i in range(B)], dtype=np.int32)
    m, d = ops.normalize_constants(mean, std)
    out = ops.load_norm_aug(torch.from_numpy(src).to(cuda), torch.from_numpy(src_idx).to(cuda),
                            torch.from_numpy(exp_id).to(cuda), torch.tensor(codes, dtype=torch.uint8, device=cuda),
                            torch.tensor(crops, dtype=torch.int32, device=cuda), torch.from_numpy(m).to(cuda),
                            torch.from_numpy(d).to(cuda), out_hw, fmt)
    torch.cuda.synchronize()
    refs = []
    for b in range(B):
        c = codes[b]
        refs.append(O.transform(src[src_idx[b]], mean[exp_id[b]], std[exp_id[b]], vflip=bool(c & 1), hflip=bool(c & 2),
                                k=(c >> 2) & 3, crop_yx=crops[b], out_hw=out_hw, ref_compat=bool(c & 16)))
    return out, np.stack(refs)


@pytest.mark.parametrize("S,out_hw", [(128, (128, 128)), (128, (92, 92)), (512, (364, 364)), (512, (512, 512))])
def test_loader_f32_bit_exact_all_canonical_d4(cuda, S, out_hw):
    codes = [ops.aug_code(v, h, k) for v in (0, 1) for h in (0, 1) for k in range(4)]
    rng = np.random.default_rng(9)
    crops = [(int(rng.integers(0, S - out_hw[0] + 1)), int(rng.integers(0, S - out_hw[1] + 1))) for _ in codes]
    out, ref = _loader_case(cuda, S, out_hw, codes, crops, ops.OUT_F32_NCHW)
    # bit-exact: same gather index AND same float32 arithmetic as the numpy restatement
    np.testing.assert_array_equal(out.cpu().numpy().view(np.uint32), ref.view(np.uint32))


def test_loader_reference_compatible_rotation_bit_exact(cuda):
    S, out_hw = 512, (364, 364)
    codes = [ops.aug_code(v, h, k, ref_compat=True) for v in (0, 1) for h in (0, 1) for k in range(4)]
    crops = [(0, 0), (148, 148), (74, 74), (1, 147)] * 4
    out, ref = _loader_case(cuda, S, out_hw, codes, crops, ops.OUT_F32_NCHW)
    np.testing.assert_array_equal(out.cpu().numpy().view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("fmt", [ops.OUT_BF16_NHWC8, ops.OUT_BF16_S2D32])
def test_loader_bf16_formats(cuda, fmt):
    S, out_hw = 128, (96, 64)
    codes = [ops.aug_code(v, h, k) for v in (0, 1) for h in (0, 1) for k in range(4)]
    crops = [(3, 5)] * len(codes)
    out, ref = _loader_case(cuda, S, out_hw, codes, crops, fmt)
    got = out.float().cpu().numpy()
    for b in range(len(codes)):
        exp = O.to_nhwc8_bf16(ref[b])
        if fmt == ops.OUT_BF16_S2D32:
            exp = O.to_s2d32(exp)
        np.testing.assert_array_equal(got[b], exp)   # same value rounded to bf16 (RN) -> identical


# ---------------------------------------------------------------- TTA / mask / assignment
def test_assign_matches_reference_golden_64(cuda, golden_dir):
    g = np.load(os.path.join(golden_dir, "assign_golden.npz"))
    probs = torch.softmax(torch.from_numpy(g["logits64"]), 1)   # test.py:27 on the host, like the reference
    col = torch.from_numpy(g["pg64"][:, int(g["et64"])].astype(np.int32)).to(cuda)
    plates = torch.from_numpy(g["plates64"].astype(np.int32)).to(cuda)
    pr = ops.mask_rescale_(probs.to(cuda).clone(), plates, col)
    ref_pr = O.mask_rescale(probs.numpy(), g["pg64"][:, int(g["et64"])], g["plates64"])
    np.testing.assert_array_equal(pr.cpu().numpy().view(np.uint32), ref_pr.view(np.uint32))  # bit-exact rescale
    res = ops.greedy_assign(pr)
    np.testing.assert_array_equal(res.cpu().numpy(), g["res64"].astype(np.int32))


def test_assign_matches_reference_golden_1108(cuda, golden_dir):
    g = np.load(os.path.join(golden_dir, "assign_golden.npz"))
    seed, N = int(g["seed1108"]), 1108
    logits = synth_logits(seed, N)
    pg = synth_plate_groups(seed + 1)
    plates = np.random.default_rng(seed + 2).integers(1, 5, size=N)
    probs = torch.softmax(torch.from_numpy(logits), 1).to(cuda)
    pr = ops.mask_rescale_(probs, torch.from_numpy(plates.astype(np.int32)).to(cuda),
                           torch.from_numpy(pg[:, int(g["et1108"])].astype(np.int32)).to(cuda))
    res = ops.greedy_assign(pr).cpu().numpy()
    np.testing.assert_array_equal(res, g["res1108"].astype(np.int32))
    # size-independent property: a class is never handed to two wells (class 0 is also the "no pick" value)
    picked = res[res != 0]
    assert len(set(picked.tolist())) == len(picked)


@pytest.mark.parametrize("N,C", [(1, 8), (5, 7), (37, 300), (200, 1108)])
def test_assign_vs_oracle_with_ties_and_zero_rows(cuda, N, C):
    rng = np.random.default_rng(N * 1000 + C)
    p = rng.random((N, C), dtype=np.float32)
    p[rng.random((N, C)) < 0.5] = 0       # many exact zeros
    if N > 3:
        p[2] = 0                          # an all-zero row
        p[3] = p[1]                       # duplicated row -> exact ties between rows
    p = O.rescale(p)
    ref = O.greedy_assign(p)
    got = ops.greedy_assign(torch.from_numpy(p).to(cuda)).cpu().numpy()
    np.testing.assert_array_equal(got, ref.astype(np.int32))


def test_tta_softmax_average_mask(cuda):
    V, N, C = 16, 33, 1108
    rng = np.random.default_rng(4)
    logits = (rng.standard_normal((V, N, C)) * 2).astype(np.float32)
    pg = synth_plate_groups(5)[:, 1]
    plates = rng.integers(1, 5, size=N)
    got = ops.tta_softmax_avg_mask(torch.from_numpy(logits).to(cuda), torch.from_numpy(plates.astype(np.int32)).to(cuda),
                                   torch.from_numpy(pg.astype(np.int32)).to(cuda)).cpu().numpy()
    probs = np.mean([O.softmax(logits[v].astype(np.float64)) for v in range(V)], axis=0)
    ref = O.mask_rescale(probs, pg, plates)
    np.testing.assert_allclose(got, ref, rtol=2e-5, atol=1e-9)
    # mask is bit-exact: exactly the disallowed classes are zero
    np.testing.assert_array_equal(got == 0, ref == 0)
    # V = 1 and no mask == plain softmax (the reference's test.py:27 with identity views)
    one = ops.tta_softmax_avg_mask(torch.from_numpy(logits[:1]).to(cuda)).cpu().numpy()
    np.testing.assert_allclose(one, O.softmax(logits[0].astype(np.float64)), rtol=2e-5, atol=1e-9)


# ---------------------------------------------------------------- CE and SGD
def test_softmax_ce_matches_torch(cuda):
    B, C = 24, 1108
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(B, C, generator=g) * 3
    target = torch.randint(0, C, (B,), generator=g)
    ref_logits = logits.clone().requires_grad_(True)
    ref_loss = torch.nn.CrossEntropyLoss()(ref_logits, target)
    ref_loss.backward()
    loss_rows, d = ops.softmax_ce(logits.to(cuda), target.to(cuda), grad_scale=1.0 / B)
    assert abs(loss_rows.mean().item() - ref_loss.item()) < 1e-5 * abs(ref_loss.item())
    np.testing.assert_allclose(d.cpu().numpy(), ref_logits.grad.numpy(), rtol=1e-4, atol=1e-8)


def test_sgd_nesterov_matches_torch(cuda):
    n = 10007
    g = torch.Generator().manual_seed(1)
    p0 = torch.randn(n, generator=g)
    ref_p = p0.clone().requires_grad_(True)
    opt = O.sgd_reference([ref_p], lr=0.008)
    p = p0.clone().to(cuda)
    mom = torch.zeros(n, device=cuda)
    for step in range(3):
        grad = torch.randn(n, generator=g)
        ref_p.grad = grad.clone()
        opt.step()
        ops.sgd_step_(p, grad.to(cuda), mom, lr=0.008, momentum=0.9, weight_decay=3e-5, nesterov=True)
    np.testing.assert_allclose(p.cpu().numpy(), ref_p.detach().numpy(), rtol=1e-5, atol=1e-7)
