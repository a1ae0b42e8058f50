"""GPU parity tests for the tcgen05 implicit-GEMM convolution kernels against torch fp32 on the same
bf16-rounded operands (tolerance: bf16 output rounding, 2e-2 relative per the north star)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from recursion_cellular_image_classification_b200 import ops

pytestmark = pytest.mark.gpu


def _rand_bf16(shape, gen, scale=1.0):
    return (torch.randn(*shape, generator=gen) * scale).to(torch.bfloat16)


def _ref_conv(A_nhwc, W_oihw, pad, scale=None, shift=None):
    x = A_nhwc.float().permute(0, 3, 1, 2)
    if scale is not None:
        # the kernel applies the fold as ONE packed-bf16 fused multiply-add + ReLU in shared memory: scale and shift
        # are bf16 operands, the exact fused result is rounded once to the bf16 the MMA reads
        scale, shift = scale.bfloat16().float(), shift.bfloat16().float()
        x = torch.relu(x * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
        x = x.to(torch.bfloat16).float()
    return F.conv2d(x, W_oihw.float(), padding=pad).permute(0, 2, 3, 1)


def _tap_major(W_oihw):
    return W_oihw.permute(2, 3, 0, 1).contiguous()   # [ty,tx,Cout,Cin]


CASES = [
    # B, H, W, Cin, ldA, Cout, k, prologue
    (2, 16, 16, 64, 64, 128, 1, False),
    (2, 16, 16, 64, 64, 128, 1, True),
    (3, 8, 8, 96, 256, 128, 1, True),       # Cin not a multiple of 64, A is a slice of a wider concat buffer
    (2, 16, 16, 128, 128, 32, 3, True),      # dense-layer 3x3
    (1, 32, 32, 128, 128, 32, 3, False),
    (5, 4, 4, 256, 256, 128, 1, True),       # several images per 128-pixel tile
    (2, 12, 20, 160, 192, 128, 1, True),     # non power-of-two spatial size
    (2, 12, 20, 128, 128, 32, 3, True),
    (1, 16, 16, 512, 512, 256, 1, False),    # transition-like, N = 256
    (1, 16, 16, 32, 32, 64, 4, False),       # space-to-depth stem: 4x4 taps, 32 channels per tap
    (2, 128, 128, 128, 128, 32, 3, True),    # block-1 geometry: x-merged 14-wide tiles (128 = 9*14 + 2)
    (1, 24, 72, 128, 128, 32, 3, True),      # x-merged tiles clipped in x (72 = 5*14 + 2) with three tile rows
    (2, 20, 60, 128, 128, 32, 3, False),     # x-merged, H not a multiple of the 8-row tile
    (1, 40, 72, 32, 32, 64, 4, False),       # stem geometry with a full-halo box clipped on every side
    # the benchmark's regime: many tiles per persistent CTA (2048 tiles / 148 CTAs ~ 14), so the pipeline stages, the
    # two accumulator stages and both staging buffers wrap several times
    (16, 128, 128, 224, 256, 128, 1, True),  # last dense layer of block 1
    (16, 128, 128, 128, 128, 32, 3, True),   # x-merged 3x3, 2960 tiles / 148 CTAs = 20 per CTA
    # ResNet-50 bottleneck shapes (csrc/resnet.cu): wide 3x3 convs whose weight panel is streamed (row-halo mode),
    # several N tiles, odd image sizes, and the 2x2-tap form of a stride-2 3x3 over a space-to-depth input
    (2, 23, 23, 256, 256, 256, 3, True),
    (3, 12, 12, 512, 512, 512, 3, True),
    (2, 46, 46, 128, 128, 128, 3, True),
    (2, 91, 91, 64, 64, 64, 3, True),
    (2, 46, 46, 512, 512, 128, 2, False),
    (4, 6, 6, 2048, 2048, 512, 2, False),
    (2, 12, 12, 2048, 2048, 512, 1, False),
    (2, 23, 23, 256, 256, 1024, 1, True),
]


@pytest.mark.parametrize("B,H,W,Cin,ldA,Cout,k,prologue", CASES)
def test_conv_fwd_matches_torch(cuda, B, H, W, Cin, ldA, Cout, k, prologue):
    gen = torch.Generator().manual_seed(B * 1000 + H * 10 + Cin + k)
    A = _rand_bf16((B, H, W, ldA), gen)
    Wt = _rand_bf16((Cout, Cin, k, k), gen, scale=(Cin * k * k) ** -0.5)
    scale = shift = None
    if prologue:
        scale = torch.rand(Cin, generator=gen) + 0.5
        shift = torch.randn(Cin, generator=gen) * 0.3
    pad = {1: 0, 2: 1, 3: 1, 4: 2}[k]
    # the 4x4 stem reads rows y+ty-2: torch's symmetric padding 2 gives one extra row/col at the end
    ref = _ref_conv(A[..., :Cin], Wt, pad, scale, shift)[:, :H, :W]
    ldC, c_off = Cout + 64, 32
    out = torch.full((B, H, W, ldC), 7.0, dtype=torch.bfloat16, device=cuda)
    out, cs, cq = ops.conv_fwd(A.to(cuda), _tap_major(Wt).to(cuda), Cin=Cin,
                               scale=None if scale is None else scale.to(cuda),
                               shift=None if shift is None else shift.to(cuda), out=out, c_off=c_off, pad=(pad, pad),
                               stats=True)
    torch.cuda.synchronize()
    got = out[..., c_off:c_off + Cout].float().cpu()
    err = (got - ref).abs().max().item()
    ref_mag = ref.abs().max().item()
    assert err <= 2e-2 * ref_mag, "max err %g vs magnitude %g" % (err, ref_mag)
    # channels outside [c_off, c_off+Cout) untouched (concat-by-offset)
    assert torch.all(out[..., :c_off].float() == 7.0) and torch.all(out[..., c_off + Cout:].float() == 7.0)
    # BatchNorm statistics of the bf16-rounded output
    g64 = got.double().reshape(-1, Cout)
    np.testing.assert_allclose(cs[c_off:c_off + Cout].cpu().numpy(), g64.sum(0).numpy(), rtol=2e-3,
                               atol=2e-3 * g64.abs().sum(0).max().item())
    np.testing.assert_allclose(cq[c_off:c_off + Cout].cpu().numpy(), (g64 ** 2).sum(0).numpy(), rtol=2e-3)


WG_CASES = [
    # B, H, W, Cin, ldA, Cout, k, prologue
    (2, 16, 16, 64, 64, 128, 1, False),
    (2, 16, 16, 96, 256, 128, 1, True),
    (2, 16, 16, 128, 128, 32, 3, True),
    (4, 8, 8, 640, 1024, 128, 1, True),      # 5 accumulator groups -> two chunk groups
    (2, 12, 20, 128, 128, 32, 3, False),
    (1, 16, 16, 256, 256, 256, 1, False),    # transition-like
    (1, 16, 16, 32, 32, 64, 4, False),       # stem-like: ONE full-halo input box, a filter row per accumulator group
    (1, 40, 24, 32, 32, 64, 4, False),       # ... with tiles hanging over the image edge
    (2, 64, 64, 128, 128, 32, 3, True),      # full-halo dOut box, a filter row's taps as one N = 96 MMA
    (1, 20, 28, 128, 128, 32, 3, True),      # ... image not a multiple of the 8x16 tile
    (2, 32, 32, 1024, 1024, 128, 1, True),   # 8 channel chunks: dOut tile shared by the chunks of a CTA
    (16, 128, 128, 224, 256, 128, 1, True),  # benchmark regime: ~14 pixel tiles per CTA, pipeline wraps
    (16, 128, 128, 128, 128, 32, 3, True),
]


@pytest.mark.parametrize("B,H,W,Cin,ldA,Cout,k,prologue", WG_CASES)
def test_conv_wgrad_matches_torch(cuda, B, H, W, Cin, ldA, Cout, k, prologue):
    gen = torch.Generator().manual_seed(B * 77 + H + Cin + k)
    A = _rand_bf16((B, H, W, ldA), gen)
    dOut = _rand_bf16((B, H, W, Cout), gen)
    scale = shift = None
    if prologue:
        scale = torch.rand(Cin, generator=gen) + 0.5
        shift = torch.randn(Cin, generator=gen) * 0.3
    pad = {1: 0, 3: 1, 4: 2}[k]
    x = A[..., :Cin].float().permute(0, 3, 1, 2)
    if prologue:
        sb, hb = scale.bfloat16().float(), shift.bfloat16().float()     # bf16 fold operands (see _ref_conv)
        x = torch.relu(x * sb.view(1, -1, 1, 1) + hb.view(1, -1, 1, 1)).to(torch.bfloat16).float()
    x = x.double().requires_grad_(False)
    w = torch.zeros(Cout, Cin, k, k, dtype=torch.double, requires_grad=True)
    y = F.conv2d(x, w, padding=pad)[:, :, :H, :W]
    y.backward(dOut.double().permute(0, 3, 1, 2))
    ref = w.grad.float()
    got = ops.conv_wgrad(A.to(cuda), dOut.to(cuda), Cin, Cout, taps=(k, k), pad=(pad, pad),
                         scale=None if scale is None else scale.to(cuda),
                         shift=None if shift is None else shift.to(cuda)).cpu()
    err = (got - ref).abs().max().item()
    assert err <= 2e-3 * ref.abs().max().item() + 1e-3, "max err %g vs %g" % (err, ref.abs().max().item())


DG_CASES = [
    # B, H, W, Cd (dOut channels), Cx (result channels), ldX, k, out_mode
    (2, 16, 16, 32, 128, 128, 3, 0),       # dense-layer 3x3 dgrad -> dy2
    (2, 12, 20, 32, 128, 128, 3, 0),       # non power-of-two image: tiles hang over the edge
    (2, 16, 16, 128, 96, 256, 1, 2),       # dense-layer 1x1 dgrad accumulating into the concat gradient
    (3, 8, 8, 128, 224, 256, 1, 1),
    (2, 64, 64, 32, 128, 128, 3, 0),
    (1, 128, 128, 128, 160, 256, 1, 2),
    # benchmark regime: ~14 tiles per CTA, so every in-place activation/staging buffer (four for 1x1, two for 3x3) is
    # reused several times by each epilogue group
    (16, 128, 128, 128, 224, 256, 1, 2),
    (16, 128, 128, 32, 128, 128, 3, 0),
    (16, 128, 128, 128, 224, 256, 1, 0),
]


@pytest.mark.parametrize("B,H,W,Cd,Cx,ldX,k,out_mode", DG_CASES)
def test_conv_dgrad_bn_matches_torch(cuda, B, H, W, Cd, Cx, ldX, k, out_mode):
    """rxb_conv_dgrad_bn == conv_transpose-free restatement: acc = conv(dOut, Wt) ; dy = acc*[x*s+h>0] ;
    sums of dy and dy*x ; out per out_mode."""
    gen = torch.Generator().manual_seed(B * 31 + H + Cd + Cx + k)
    dOut = _rand_bf16((B, H, W, Cd), gen)
    Wt = _rand_bf16((Cx, Cd, k, k), gen, scale=(Cd * k * k) ** -0.5)   # operand as the kernel contracts it
    X = _rand_bf16((B, H, W, ldX), gen)
    s = torch.rand(Cx, generator=gen) + 0.5
    h = torch.randn(Cx, generator=gen) * 0.3
    G0 = _rand_bf16((B, H, W, ldX), gen)
    pad = {1: 0, 3: 1}[k]
    acc = _ref_conv(dOut, Wt, pad)                                   # [B,H,W,Cx] fp32
    xv = X[..., :Cx].float()
    dy = acc * ((xv * s + h) > 0)
    if out_mode == 0:
        ref = dy
    elif out_mode == 1:
        ref = s * dy
    else:
        ref = G0[..., :Cx].float() + s * dy
    out = G0.clone().to(cuda)
    out, s1 = ops.conv_dgrad_bn(dOut.to(cuda), _tap_major(Wt).to(cuda), X.to(cuda), s.to(cuda), h.to(cuda), Cx,
                                out_mode=out_mode, out=out, pad=(pad, pad))
    torch.cuda.synchronize()
    got = out[..., :Cx].float().cpu()
    err = (got - ref).abs().max().item()
    assert err <= 2e-2 * ref.abs().max().item(), "max err %g vs %g" % (err, ref.abs().max().item())
    assert torch.equal(out[..., Cx:].cpu(), G0[..., Cx:]), "channels beyond Cout must stay untouched"
    d64 = dy.double().reshape(-1, Cx)
    # the kernel sums the bf16-rounded staged tile on the tensor pipe (fp32 accumulate)
    np.testing.assert_allclose(s1.cpu().numpy(), d64.sum(0).numpy(), rtol=4e-3,
                               atol=4e-3 * d64.abs().sum(0).max().item())


WGF_CASES = [
    # B, H, W, Cd (dOut channels = the forward conv's outputs), Cx (its inputs), ldX, out_mode, degenerate channels
    (2, 16, 16, 128, 96, 256, 2, False),
    (3, 8, 8, 128, 224, 256, 1, False),        # two images per 128-pixel tile, the last tile half empty
    (2, 16, 16, 64, 128, 128, 0, False),       # one dY stage
    (1, 128, 128, 128, 160, 256, 2, False),
    (2, 12, 20, 128, 160, 192, 2, True),       # tiles hang over the image edge; direct reductions before the transform
    (2, 32, 32, 128, 992, 1024, 2, False),     # eight N tiles (block 3/4 widths)
    (16, 128, 128, 128, 224, 256, 2, False),   # benchmark regime: ~14 tiles per CTA, every buffer and stage wraps
]


@pytest.mark.parametrize("B,H,W,Cd,Cx,ldX,out_mode,degenerate", WGF_CASES)
def test_conv_dgrad_bn_fused_wgrad(cuda, B, H, W, Cd, Cx, ldX, out_mode, degenerate):
    """rxb_conv_dgrad_bn_wgrad: the 1x1 data gradient with the weight gradient of the same convolution accumulated by
    the same kernel.  A' = bf16(relu(x*bf16(s) + bf16(h))) is what the forward prologue fed the convolution:
    dW[k][c] = sum_p dOut[p,k] * A'[p,c], and the ReLU mask of dy is A' > 0."""
    gen = torch.Generator().manual_seed(B * 131 + H + Cd + Cx)
    dOut = _rand_bf16((B, H, W, Cd), gen)
    Wt = _rand_bf16((Cx, Cd, 1, 1), gen, scale=Cd ** -0.5)
    X = _rand_bf16((B, H, W, ldX), gen)
    gamma = torch.rand(Cx, generator=gen) + 0.5
    beta = torch.randn(Cx, generator=gen) * 0.3
    flagged = []
    if degenerate:
        gamma[3], gamma[17], gamma[64], gamma[100] = 0.0, 1e-4, -2e-4, 0.01
        beta[100], beta[3] = 0.9, 0.25
        flagged = [3, 17, 64, 100]
    rstd = torch.rand(Cx, generator=gen) + 0.5
    mean = torch.randn(Cx, generator=gen) * 0.1
    s = gamma * rstd
    h = beta - mean * s
    G0 = _rand_bf16((B, H, W, ldX), gen)
    acc = _ref_conv(dOut, Wt, 0)
    xv = X[..., :Cx].float()
    a_prime = torch.relu(xv * s.bfloat16().float() + h.bfloat16().float()).to(torch.bfloat16).float()
    dy = acc * (a_prime > 0)
    ref = {0: dy, 1: s * dy, 2: G0[..., :Cx].float() + s * dy}[out_mode]
    ref_dW = dOut.double().reshape(-1, Cd).t() @ a_prime.double().reshape(-1, Cx)          # [Cd, Cx]
    out = G0.clone().to(cuda)
    kw = dict(bn_gamma=gamma.to(cuda), bn_beta=beta.to(cuda)) if degenerate else {}
    res = ops.conv_dgrad_bn(dOut.to(cuda), _tap_major(Wt).to(cuda), X.to(cuda), s.to(cuda), h.to(cuda), Cx,
                            out_mode=out_mode, out=out, wgrad=True, **kw)
    torch.cuda.synchronize()
    out, s1, dW = res[0], res[1], res[-1]
    got = out[..., :Cx].float().cpu()
    err = (got - ref).abs().max().item()
    assert err <= 2e-2 * ref.abs().max().item(), "max err %g vs %g" % (err, ref.abs().max().item())
    assert torch.equal(out[..., Cx:].cpu(), G0[..., Cx:]), "channels beyond Cout must stay untouched"
    d64 = dy.double().reshape(-1, Cx)
    np.testing.assert_allclose(s1.cpu().numpy(), d64.sum(0).numpy(), rtol=4e-3,
                               atol=4e-3 * d64.abs().sum(0).max().item())
    werr = (dW.cpu().double() - ref_dW).abs().max().item()
    assert werr <= 2e-3 * ref_dW.abs().max().item() + 1e-3, "dW max err %g vs %g" % (werr, ref_dW.abs().max().item())
    if degenerate:
        s2 = res[2].cpu().numpy()
        # the direct reductions use the mask of the raw activation with the fp32 fold (they run before the transform)
        dy_raw = (acc * ((xv * s + h) > 0)).double().reshape(-1, Cx)
        want2 = (dy_raw * xv.double().reshape(-1, Cx)).sum(0).numpy()
        np.testing.assert_allclose(s2[flagged], want2[flagged], rtol=1e-4,
                                   atol=1e-5 * (dy_raw * xv.double().reshape(-1, Cx)).abs().sum(0).max().item())
        assert (s2[[c for c in range(Cx) if c not in flagged]] == 0).all()


def test_conv_dgrad_bn_fused_wgrad_repeatable_output(cuda):
    """50 launches at ~14 tiles per CTA: the data-gradient output is bit-identical every time (the tile hand-offs
    between the epilogue groups, the MMA warp and the TMA engine hold), dW agrees to accumulation-order noise."""
    B, H, W, Cd, Cx, ldX = 16, 128, 128, 128, 224, 256
    gen = torch.Generator().manual_seed(5)
    dOut = _rand_bf16((B, H, W, Cd), gen).to(cuda)
    Wt = _tap_major(_rand_bf16((Cx, Cd, 1, 1), gen, scale=Cd ** -0.5)).to(cuda)
    X = _rand_bf16((B, H, W, ldX), gen).to(cuda)
    s = (torch.rand(Cx, generator=gen) + 0.5).to(cuda)
    h = (torch.randn(Cx, generator=gen) * 0.3).to(cuda)
    first = None
    for i in range(50):
        out, s1, dW = ops.conv_dgrad_bn(dOut, Wt, X, s, h, Cx, out_mode=1, wgrad=True)
        torch.cuda.synchronize()
        if first is None:
            first = (out.clone(), dW.clone())
        else:
            assert torch.equal(out, first[0]), "launch %d: output differs" % i
            assert (dW - first[1]).abs().max().item() <= 1e-4 * first[1].abs().max().item()


WGF3_CASES = [
    # B, H, W, degenerate channels
    (2, 16, 16, False),
    (2, 12, 20, True),         # tiles hang over the image edge; direct reductions before the transform
    (1, 20, 28, False),
    (2, 64, 64, False),
    (3, 32, 32, False),        # block-3 geometry
    (16, 128, 128, False),     # benchmark regime: ~110 tiles per CTA at B = 128, 14 here
]


@pytest.mark.parametrize("B,H,W,degenerate", WGF3_CASES)
def test_conv_dgrad3x3_bn_fused_wgrad(cuda, B, H, W, degenerate):
    """rxb_conv_dgrad_bn_wgrad on the dense layers' 3x3 (dZ 32 channels -> 128): data gradient + ReLU/BN2 backward
    and the OIHW weight gradient of the same convolution from one kernel."""
    Cd, Cx = 32, 128
    gen = torch.Generator().manual_seed(B * 17 + H + W)
    dOut = _rand_bf16((B, H, W, Cd), gen)
    Wt = _rand_bf16((Cx, Cd, 3, 3), gen, scale=(Cd * 9) ** -0.5)
    X = _rand_bf16((B, H, W, Cx), gen)
    gamma = torch.rand(Cx, generator=gen) + 0.5
    beta = torch.randn(Cx, generator=gen) * 0.3
    flagged = []
    if degenerate:
        gamma[3], gamma[17], gamma[64], gamma[100] = 0.0, 1e-4, -2e-4, 0.01
        beta[100], beta[3] = 0.9, 0.25
        flagged = [3, 17, 64, 100]
    rstd = torch.rand(Cx, generator=gen) + 0.5
    mean = torch.randn(Cx, generator=gen) * 0.1
    s = gamma * rstd
    h = beta - mean * s
    acc = _ref_conv(dOut, Wt, 1)
    xv = X.float()
    a_prime = torch.relu(xv * s.bfloat16().float() + h.bfloat16().float()).to(torch.bfloat16).float()
    dy = acc * (a_prime > 0)
    w = torch.zeros(Cd, Cx, 3, 3, dtype=torch.double, requires_grad=True)
    F.conv2d(a_prime.double().permute(0, 3, 1, 2), w, padding=1).backward(dOut.double().permute(0, 3, 1, 2))
    ref_dW = w.grad
    kw = dict(bn_gamma=gamma.to(cuda), bn_beta=beta.to(cuda)) if degenerate else {}
    res = ops.conv_dgrad_bn(dOut.to(cuda), _tap_major(Wt).to(cuda), X.to(cuda), s.to(cuda), h.to(cuda), Cx,
                            out_mode=0, pad=(1, 1), wgrad=True, **kw)
    torch.cuda.synchronize()
    out, s1, dW = res[0], res[1], res[-1]
    err = (out.float().cpu() - dy).abs().max().item()
    assert err <= 2e-2 * dy.abs().max().item(), "max err %g vs %g" % (err, dy.abs().max().item())
    d64 = dy.double().reshape(-1, Cx)
    np.testing.assert_allclose(s1.cpu().numpy(), d64.sum(0).numpy(), rtol=4e-3,
                               atol=4e-3 * d64.abs().sum(0).max().item())
    werr = (dW.cpu().double() - ref_dW).abs().max().item()
    assert werr <= 2e-3 * ref_dW.abs().max().item() + 1e-3, "dW max err %g vs %g" % (werr, ref_dW.abs().max().item())
    if degenerate:
        s2 = res[2].cpu().numpy()
        dy_raw = (acc * ((xv * s + h) > 0)).double().reshape(-1, Cx)
        want2 = (dy_raw * xv.double().reshape(-1, Cx)).sum(0).numpy()
        np.testing.assert_allclose(s2[flagged], want2[flagged], rtol=1e-4,
                                   atol=1e-5 * (dy_raw * xv.double().reshape(-1, Cx)).abs().sum(0).max().item())


def test_conv_dgrad3x3_bn_fused_wgrad_repeatable_output(cuda):
    """50 launches at ~14 tiles per CTA: bit-identical data gradient every time, dW to accumulation-order noise."""
    B, H, W, Cd, Cx = 16, 128, 128, 32, 128
    gen = torch.Generator().manual_seed(6)
    dOut = _rand_bf16((B, H, W, Cd), gen).to(cuda)
    Wt = _tap_major(_rand_bf16((Cx, Cd, 3, 3), gen, scale=(Cd * 9) ** -0.5)).to(cuda)
    X = _rand_bf16((B, H, W, Cx), gen).to(cuda)
    s = (torch.rand(Cx, generator=gen) + 0.5).to(cuda)
    h = (torch.randn(Cx, generator=gen) * 0.3).to(cuda)
    first = None
    for i in range(50):
        out, s1, dW = ops.conv_dgrad_bn(dOut, Wt, X, s, h, Cx, out_mode=0, pad=(1, 1), wgrad=True)
        torch.cuda.synchronize()
        if first is None:
            first = (out.clone(), dW.clone())
        else:
            assert torch.equal(out, first[0]), "launch %d: output differs" % i
            assert (dW - first[1]).abs().max().item() <= 1e-4 * first[1].abs().max().item()


@pytest.mark.parametrize("B,H,W", [(2, 16, 16), (2, 12, 20), (3, 32, 32), (16, 128, 128)])
def test_conv_dgrad3x3_fixup_fold(cuda, B, H, W):
    """rxb_conv_dgrad3x3_bn_wgrad_fixup: dOut = G - corrA - xhat*corrB derived on load from the concat buffers (two
    strided full-halo boxes per stage, packed-bf16 transform in shared memory, zero padding kept), then the fused 3x3
    data + weight gradient.  The reference restates the packed arithmetic: t = bf16(fma(x, bf16(kb), bf16(kc))),
    dOut = bf16(g + t)."""
    ld, c0, Cx = 256, 96, 128
    gen = torch.Generator().manual_seed(B * 19 + H + W)
    G = _rand_bf16((B, H, W, ld), gen)
    Xc = _rand_bf16((B, H, W, ld), gen)
    mean = torch.randn(ld, generator=gen) * 0.2
    rstd = torch.rand(ld, generator=gen) + 0.5
    corrA = torch.randn(ld, generator=gen) * 0.05
    corrB = torch.randn(ld, generator=gen) * 0.05
    kb = (-rstd * corrB)[c0:c0 + 32].bfloat16().double()
    kc = (mean * rstd * corrB - corrA)[c0:c0 + 32].bfloat16().double()
    t = (Xc[..., c0:c0 + 32].double() * kb + kc).to(torch.bfloat16)
    dz = (G[..., c0:c0 + 32].double() + t.double()).to(torch.bfloat16)
    exact = G[..., c0:c0 + 32].float() - corrA[c0:c0 + 32] - (Xc[..., c0:c0 + 32].float() - mean[c0:c0 + 32]) * rstd[c0:c0 + 32] * corrB[c0:c0 + 32]
    assert (dz.float() - exact).abs().max().item() <= 2e-2 * exact.abs().max().item()      # the packed form is the fix-up
    Wt = _rand_bf16((Cx, 32, 3, 3), gen, scale=(32 * 9) ** -0.5)
    X = _rand_bf16((B, H, W, Cx), gen)
    s = torch.rand(Cx, generator=gen) + 0.5
    h = torch.randn(Cx, generator=gen) * 0.3
    acc = _ref_conv(dz, Wt, 1)
    a_prime = torch.relu(X.float() * s.bfloat16().float() + h.bfloat16().float()).to(torch.bfloat16).float()
    dy = acc * (a_prime > 0)
    w = torch.zeros(32, Cx, 3, 3, dtype=torch.double, requires_grad=True)
    F.conv2d(a_prime.double().permute(0, 3, 1, 2), w, padding=1).backward(dz.double().permute(0, 3, 1, 2))
    out, s1, dW = ops.conv_dgrad3x3_bn_wgrad_fixup(G.to(cuda), Xc.to(cuda), c0, mean.to(cuda), rstd.to(cuda), corrA.to(cuda),
                                                   corrB.to(cuda), _tap_major(Wt).to(cuda), X.to(cuda), s.to(cuda), h.to(cuda))
    torch.cuda.synchronize()
    err = (out.float().cpu() - dy).abs().max().item()
    assert err <= 2e-2 * dy.abs().max().item(), "max err %g vs %g" % (err, dy.abs().max().item())
    d64 = dy.double().reshape(-1, Cx)
    np.testing.assert_allclose(s1.cpu().numpy(), d64.sum(0).numpy(), rtol=4e-3, atol=4e-3 * d64.abs().sum(0).max().item())
    werr = (dW.cpu().double() - w.grad).abs().max().item()
    assert werr <= 2e-3 * w.grad.abs().max().item() + 1e-3, "dW max err %g vs %g" % (werr, w.grad.abs().max().item())


@pytest.mark.parametrize("k,Cin,Cout", [(1, 96, 128), (3, 128, 32)])
def test_bn_backward_sums_from_wdw(cuda, k, Cin, Cout):
    """The BatchNorm-backward reduction sum(dy*x) recovered from W.dW equals the direct reduction: forward
    A' = relu(s*x+h) -> conv(W); wgrad gives dW; dgrad gives dy and sum(dy); sum(dy*x) = (W.dW - h*sum dy)/s."""
    B, H, W_ = 2, 16, 16
    gen = torch.Generator().manual_seed(100 + k)
    X = _rand_bf16((B, H, W_, Cin), gen)
    s = torch.rand(Cin, generator=gen) + 0.5
    h = torch.randn(Cin, generator=gen) * 0.3
    Wc = _rand_bf16((Cout, Cin, k, k), gen, scale=(Cin * k * k) ** -0.5)       # forward weights (bf16-exact)
    dOut = _rand_bf16((B, H, W_, Cout), gen)
    pad = k // 2
    dW = ops.conv_wgrad(X.to(cuda), dOut.to(cuda), Cin, Cout, taps=(k, k), pad=(pad, pad), scale=s.to(cuda),
                        shift=h.to(cuda))
    # dgrad operand: Wt[tap_flipped][cin][cout]
    Wd = Wc.flip(2, 3).permute(1, 0, 2, 3).contiguous()                          # [Cin, Cout, k, k]
    out, s1 = ops.conv_dgrad_bn(dOut.to(cuda), _tap_major(Wd).to(cuda), X.to(cuda), s.to(cuda), h.to(cuda), Cin,
                                out_mode=0, pad=(pad, pad))
    s2 = ops.bn_sum_dyx_from_wdw(Wc.float().to(cuda), dW, s.to(cuda), h.to(cuda), s1)
    torch.cuda.synchronize()
    # reference in fp64 through autograd
    x = X.double().permute(0, 3, 1, 2).requires_grad_(True)
    z = x * s.double().view(1, -1, 1, 1) + h.double().view(1, -1, 1, 1)
    a = torch.relu(z).detach().to(torch.bfloat16).double() + (torch.relu(z) - torch.relu(z).detach())  # bf16 value, relu grad
    y = F.conv2d(a, Wc.double(), padding=pad)
    z.retain_grad()
    y.backward(dOut.double().permute(0, 3, 1, 2))
    dy = z.grad                                                                   # [B,Cin,H,W]
    ref_s1 = dy.sum((0, 2, 3))
    ref_s2 = (dy * x.detach()).sum((0, 2, 3))
    scale1 = dy.abs().sum((0, 2, 3)).max().item()
    np.testing.assert_allclose(s1.cpu().numpy(), ref_s1.numpy(), rtol=4e-3, atol=4e-3 * scale1)
    scale2 = (dy * x.detach()).abs().sum((0, 2, 3)).max().item()
    np.testing.assert_allclose(s2.cpu().numpy(), ref_s2.numpy(), rtol=1e-2, atol=1e-2 * scale2)


@pytest.mark.parametrize("kind,nx", [("1x1", None), ("1x1", "3"), ("3x3", None), ("3x3", "3")])
def test_dgrad_repeat_launch_stress(cuda, kind, nx):
    """200 identical launches of the fused data-gradient kernel at ~14 tiles per CTA: bit-identical output every time
    (tools/stress_dgrad.py), for the shipped buffer count and for the odd count (RXB_DBG_NX=3, per-group barriers)."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    if nx is not None:
        env["RXB_DBG_NX"] = nx
    p = subprocess.run([sys.executable, os.path.join(root, "tools", "stress_dgrad.py"), "--iters", "200", "--kind", kind],
                       capture_output=True, text=True, env=env, timeout=300)
    print(p.stdout.strip())
    assert p.returncode == 0, p.stdout + p.stderr
    res = json.loads(p.stdout.strip().splitlines()[-1])
    assert res["mismatching_launches"] == 0 and res["tiles_per_cta"] > 10


@pytest.mark.parametrize("k,out_mode", [(1, 2), (3, 0)])
def test_dgrad_direct_reductions_for_degenerate_channels(cuda, k, out_mode):
    """Channels whose BatchNorm weight is tiny, zero or small against its bias get sum(dy) and sum(dy*x) reduced
    directly in the epilogue (fp32, unscaled dy) — exact even where scale = gamma*rstd is 0 and the staged value
    scale*dy carries no information; every other channel keeps the tensor-pipe sum and leaves sum_dyx untouched."""
    B, H, W = 3, 24, 20
    Cd, Cx = (128, 160) if k == 1 else (32, 128)
    gen = torch.Generator().manual_seed(41 + k)
    dOut = _rand_bf16((B, H, W, Cd), gen)
    Wt = _rand_bf16((Cx, Cd, k, k), gen, scale=(Cd * k * k) ** -0.5)
    X = _rand_bf16((B, H, W, Cx), gen)
    gamma = torch.rand(Cx, generator=gen) + 0.5
    beta = torch.randn(Cx, generator=gen) * 0.3
    gamma[3], gamma[17], gamma[64], gamma[100] = 0.0, 1e-4, -2e-4, 0.01
    beta[100] = 0.9                                     # |gamma| < 0.05 |beta|
    beta[3] = 0.25                                      # gamma = 0, positive shift: the ReLU passes every pixel
    flagged = [3, 17, 64, 100]
    rstd = torch.rand(Cx, generator=gen) + 0.5
    mean = torch.randn(Cx, generator=gen) * 0.1
    s = gamma * rstd
    h = beta - mean * s
    G0 = _rand_bf16((B, H, W, Cx), gen)
    pad = k // 2
    acc = _ref_conv(dOut, Wt, pad)
    xv = X.float()
    dy = acc * ((xv * s + h) > 0)
    ref = dy if out_mode == 0 else G0.float() + s * dy
    out = G0.clone().to(cuda)
    out, s1, s2 = ops.conv_dgrad_bn(dOut.to(cuda), _tap_major(Wt).to(cuda), X.to(cuda), s.to(cuda), h.to(cuda), Cx,
                                    out_mode=out_mode, out=out, pad=(pad, pad), bn_gamma=gamma.to(cuda),
                                    bn_beta=beta.to(cuda))
    torch.cuda.synchronize()
    err = (out.float().cpu() - ref).abs().max().item()
    assert err <= 2e-2 * ref.abs().max().item()
    d64 = dy.double().reshape(-1, Cx)
    x64 = xv.double().reshape(-1, Cx)
    tol1 = 4e-3 * d64.abs().sum(0).max().item()
    np.testing.assert_allclose(s1.cpu().numpy(), d64.sum(0).numpy(), rtol=4e-3, atol=tol1)
    want2 = (d64 * x64).sum(0).numpy()
    got2 = s2.cpu().numpy()
    others = [c for c in range(Cx) if c not in flagged]
    assert (got2[others] == 0).all()
    # direct fp32 reductions: far tighter than the bf16-staged sums
    np.testing.assert_allclose(s1.cpu().numpy()[flagged], d64.sum(0).numpy()[flagged], rtol=1e-4,
                               atol=1e-5 * d64.abs().sum(0).max().item())
    np.testing.assert_allclose(got2[flagged], want2[flagged], rtol=1e-4, atol=1e-5 * (d64 * x64).abs().sum(0).max().item())
    assert abs(d64.sum(0)[3].item()) > 0                  # the gamma = 0 channel really has a non-zero sum(dy)
