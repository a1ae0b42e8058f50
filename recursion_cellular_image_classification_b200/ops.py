"""Tensor-level wrappers over the C ABI (include/rxb.h).  Every function runs on the current CUDA
stream of the tensors' device and raises RxbError on failure; none has a CPU path."""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import ConvDesc, check, load, ptr, require_gpu, stream_ptr

OUT_F32_NCHW, OUT_BF16_NHWC8, OUT_BF16_S2D32 = 0, 1, 2
AUG_VFLIP, AUG_HFLIP, AUG_REF_COMPAT = 1, 2, 16


def aug_code(vflip=False, hflip=False, k=0, ref_compat=False):
    """bit0 vflip, bit1 hflip, bits2-3 k quarter turns CCW, bit4 reference-compatible rotation."""
    return (1 if vflip else 0) | (2 if hflip else 0) | ((int(k) & 3) << 2) | (16 if ref_compat else 0)


def _cuda(t, dtype=None):
    if not t.is_cuda:
        raise _lib.RxbError("expected a CUDA tensor (librxb has no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise _lib.RxbError("expected dtype %s, got %s" % (dtype, t.dtype))
    return t.contiguous()


# ------------------------------------------------------------------ family 1a: statistics
def stats_accumulate(imgs, exp_id, n_exp, acc=None):
    """imgs u8 [n,C,H,W] planar; exp_id int32 [n].  Returns (sum, sumsq, count) int64 [n_exp, C]
    (exact integer accumulators; pass `acc` to keep accumulating over chunks)."""
    require_gpu()
    imgs = _cuda(imgs, torch.uint8)
    exp_id = _cuda(exp_id, torch.int32)
    n, C, H, W = imgs.shape
    if acc is None:
        acc = tuple(torch.zeros(n_exp, C, dtype=torch.int64, device=imgs.device) for _ in range(3))
    s, q, cnt = acc
    check(load().rxb_stats_accumulate(ptr(imgs), ptr(exp_id), n, H, W, C, 0, n_exp, ptr(s), ptr(q), ptr(cnt),
                                      stream_ptr()))
    return acc


def stats_finalize(acc, pre_mean=None, pre_std=None):
    """(mean, std) float64 [n_exp, C] of x/255 — or of (x/255-pre_mean)/pre_std (verification mode)."""
    require_gpu()
    s, q, cnt = acc
    n_exp, C = s.shape
    mean = torch.empty(n_exp, C, dtype=torch.float64, device=s.device)
    std = torch.empty_like(mean)
    pm = _cuda(pre_mean, torch.float64) if pre_mean is not None else None
    ps = _cuda(pre_std, torch.float64) if pre_std is not None else None
    check(load().rxb_stats_finalize(ptr(s), ptr(q), ptr(cnt), n_exp, C, ptr(pm), ptr(ps), ptr(mean), ptr(std),
                                    stream_ptr()))
    return mean, std


# ------------------------------------------------------------------ family 1b: loader
def normalize_constants(mean, std):
    """albumentations-0.3.0 Normalize constants: m = f32(mean)*255, d = 1/(f32(std)*255),
    all float32.  mean/std: float64 [..., 6] as stored in the stats pickle."""
    m = np.asarray(mean, dtype=np.float32) * np.float32(255.0)
    s = np.asarray(std, dtype=np.float32) * np.float32(255.0)
    d = np.reciprocal(s, dtype=np.float32)
    return m.astype(np.float32), d.astype(np.float32)


def load_norm_aug(src, src_idx, exp_id, aug, crop_yx, norm_m, norm_d, out_hw, out_format, out=None):
    """src u8 [n,6,H,W]; src_idx/exp_id int32 [B]; aug u8 [B]; crop_yx int32 [B,2];
    norm_m/norm_d float32 [n_exp,6].  Returns the normalised, augmented batch in `out_format`."""
    require_gpu()
    src = _cuda(src, torch.uint8)
    n, C, H, W = src.shape
    if C != 6:
        raise _lib.RxbError("loader expects 6 channels")
    src_idx = _cuda(src_idx, torch.int32)
    exp_id = _cuda(exp_id, torch.int32)
    aug = _cuda(aug, torch.uint8)
    crop_yx = _cuda(crop_yx, torch.int32)
    norm_m = _cuda(norm_m, torch.float32)
    norm_d = _cuda(norm_d, torch.float32)
    B = src_idx.numel()
    Ho, Wo = out_hw
    if out is None:
        if out_format == OUT_F32_NCHW:
            out = torch.empty(B, 6, Ho, Wo, dtype=torch.float32, device=src.device)
        elif out_format == OUT_BF16_NHWC8:
            out = torch.empty(B, Ho, Wo, 8, dtype=torch.bfloat16, device=src.device)
        else:
            out = torch.empty(B, Ho // 2, Wo // 2, 32, dtype=torch.bfloat16, device=src.device)
    check(load().rxb_load_norm_aug(ptr(src), n, H, W, ptr(src_idx), ptr(exp_id), ptr(aug), ptr(crop_yx),
                                   ptr(norm_m), ptr(norm_d), norm_m.shape[0], ptr(out), B, Ho, Wo, out_format,
                                   stream_ptr()))
    return out


JPEG_STATUS = {1: "not a JPEG / truncated headers", 2: "unsupported JPEG (progressive, lossless, arithmetic, 12-bit or "
               "multi-component)", 3: "missing or malformed quantisation / Huffman table",
               4: "frame size differs from the expected size", 5: "corrupt entropy-coded data"}


def jpeg_frame_size(buf):
    """(H, W) from the SOF segment of a JPEG byte string (host-side header peek: sizes the output tensor)."""
    p = 2
    while p + 9 < len(buf):
        if buf[p] != 0xFF or buf[p + 1] == 0xFF:
            p += 1
            continue
        m = buf[p + 1]
        if m in (0x00, 0x01) or 0xD0 <= m <= 0xD8:
            p += 2
            continue
        if 0xC0 <= m <= 0xCF and m not in (0xC4, 0xC8, 0xCC):
            return (buf[p + 5] << 8) | buf[p + 6], (buf[p + 7] << 8) | buf[p + 8]
        p += 2 + ((buf[p + 2] << 8) | buf[p + 3])
    raise _lib.RxbError("no JPEG frame header found")


def pack_jpeg_buffers(buffers):
    """List of JPEG byte strings -> (blob uint8 [total], offsets int64 [n+1]) host tensors for jpeg_decode_gray."""
    sizes = np.fromiter((len(b) for b in buffers), dtype=np.int64, count=len(buffers))
    offsets = np.zeros(len(buffers) + 1, dtype=np.int64)
    np.cumsum(sizes, out=offsets[1:])
    blob = np.frombuffer(b"".join(buffers), dtype=np.uint8) if len(buffers) else np.zeros(0, np.uint8)
    return torch.from_numpy(blob.copy()), torch.from_numpy(offsets)


def jpeg_decode_gray(blob, offsets, hw, select=None, out=None, check_status=True, parallel=True):
    """Decode single-channel baseline JPEG files on the device (cv2.imdecode(buf, -1) bit-exact).
    blob u8 [total] and offsets int64 [n_files+1] are CUDA tensors (see pack_jpeg_buffers); `select` (int64 CUDA
    tensor of file indices) decodes that subset, in that order.  Returns u8 [n,H,W] (and the int32 [n] status tensor
    when check_status is False — checking costs a device->host sync).  parallel=True gives the kernel a coefficient
    workspace so all lanes of a file's warp decode; False uses the single-lane kernel (same result)."""
    require_gpu()
    blob = _cuda(blob, torch.uint8)
    offsets = _cuda(offsets, torch.int64)
    if select is None:
        begin, end = offsets[:-1], offsets[1:]            # views into the same array: 8-byte aligned
    else:
        select = _cuda(select, torch.int64)
        begin, end = offsets[:-1][select].contiguous(), offsets[1:][select].contiguous()
    n = begin.numel()
    H, W = hw
    if out is None:
        out = torch.empty(n, H, W, dtype=torch.uint8, device=blob.device)
    status = torch.zeros(max(n, 1), dtype=torch.int32, device=blob.device)
    ws_bytes = load().rxb_jpeg_decode_workspace_bytes(n, H, W) if parallel else 0
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=blob.device) if ws_bytes else None
    check(load().rxb_jpeg_decode_gray(ptr(blob), ptr(begin), ptr(end), n, H, W, ptr(out), ptr(status), ptr(ws),
                                      ws_bytes, stream_ptr()))
    if not check_status:
        return out, status[:n]
    bad = torch.nonzero(status[:n]).flatten().tolist()
    if bad:
        code = int(status[bad[0]])
        raise _lib.RxbError("JPEG file %d of %d: %s (status %d)" % (bad[0], n, JPEG_STATUS.get(code, "?"), code))
    return out


def rotation_matrix(w, h, angle, scale=1.0):
    """The matrix albumentations 0.3.0 ShiftScaleRotate hands to cv2.warpAffine (dataloader.py:45-46):
    cv2.getRotationMatrix2D((w/2, h/2), angle, scale), float64 [2,3] — host data, restated so the loader does not
    depend on OpenCV (bitwise equal to OpenCV's on the tested angles: the angle is scaled by the folded pi/180)."""
    import math
    cx, cy = w / 2, h / 2
    a = angle * (math.pi / 180.0)
    al, be = math.cos(a) * scale, math.sin(a) * scale
    return np.array([[al, be, (1 - al) * cx - be * cy], [-be, al, be * cx + (1 - al) * cy]], dtype=np.float64)


def load_norm_affine(src, src_idx, exp_id, flips, M, crop_yx, norm_m, norm_d, out_hw, out_format, out=None):
    """The reference's full train transform on the device: flips -> cv2.warpAffine-exact bilinear warp by the forward
    matrices M (float64 [B,2,3]) -> crop -> normalise.  Other arguments as load_norm_aug; flips u8 [B] uses bit0
    vflip, bit1 hflip."""
    require_gpu()
    src = _cuda(src, torch.uint8)
    n, C, H, W = src.shape
    if C != 6:
        raise _lib.RxbError("loader expects 6 channels")
    src_idx = _cuda(src_idx, torch.int32)
    exp_id = _cuda(exp_id, torch.int32)
    flips = _cuda(flips, torch.uint8)
    M = _cuda(M, torch.float64)
    crop_yx = _cuda(crop_yx, torch.int32)
    norm_m = _cuda(norm_m, torch.float32)
    norm_d = _cuda(norm_d, torch.float32)
    B = src_idx.numel()
    if M.numel() != 6 * B or flips.numel() != B or crop_yx.numel() != 2 * B or exp_id.numel() != B:
        raise _lib.RxbError("load_norm_affine: per-image arguments must have B=%d rows" % B)
    Ho, Wo = out_hw
    if out is None:
        if out_format == OUT_F32_NCHW:
            out = torch.empty(B, 6, Ho, Wo, dtype=torch.float32, device=src.device)
        elif out_format == OUT_BF16_NHWC8:
            out = torch.empty(B, Ho, Wo, 8, dtype=torch.bfloat16, device=src.device)
        else:
            out = torch.empty(B, Ho // 2, Wo // 2, 32, dtype=torch.bfloat16, device=src.device)
    check(load().rxb_load_norm_affine(ptr(src), n, H, W, ptr(src_idx), ptr(exp_id), ptr(flips), ptr(M), ptr(crop_yx),
                                      ptr(norm_m), ptr(norm_d), norm_m.shape[0], ptr(out), B, Ho, Wo, out_format,
                                      stream_ptr()))
    return out


# ------------------------------------------------------------------ family 4: TTA / assignment
def tta_softmax_avg_mask(logits, plate=None, group_col=None):
    """logits f32 [V,N,C] -> rescale(mask(mean_v softmax)) f32 [N,C]."""
    require_gpu()
    logits = _cuda(logits, torch.float32)
    V, N, C = logits.shape
    probs = torch.empty(N, C, dtype=torch.float32, device=logits.device)
    pl = _cuda(plate, torch.int32) if plate is not None else None
    gc = _cuda(group_col, torch.int32) if group_col is not None else None
    check(load().rxb_tta_softmax_avg_mask(ptr(logits), V, N, C, ptr(pl), ptr(gc), ptr(probs), stream_ptr()))
    return probs


def mask_rescale_(preds, plate=None, group_col=None):
    require_gpu()
    preds = _cuda(preds, torch.float32)
    N, C = preds.shape
    pl = _cuda(plate, torch.int32) if plate is not None else None
    gc = _cuda(group_col, torch.int32) if group_col is not None else None
    check(load().rxb_mask_rescale(ptr(preds), N, C, ptr(pl), ptr(gc), stream_ptr()))
    return preds


def greedy_assign(preds):
    """preds f32 [N,C] (masked + rescaled) -> int32 [N] class per row, the loop of test.py:48-56."""
    require_gpu()
    preds = _cuda(preds, torch.float32)
    N, C = preds.shape
    result = torch.zeros(N, dtype=torch.int32, device=preds.device)
    ws = torch.empty(load().rxb_greedy_assign_workspace_bytes(N, C), dtype=torch.uint8, device=preds.device)
    check(load().rxb_greedy_assign(ptr(preds), N, C, ptr(result), ptr(ws), stream_ptr()))
    return result


# ------------------------------------------------------------------ family 3: loss, optimizer
def softmax_ce(logits, target, grad_scale=None):
    """Returns (loss_rows f32 [B], dlogits f32 [B,C] or None)."""
    require_gpu()
    logits = _cuda(logits, torch.float32)
    target = _cuda(target, torch.int64)
    B, C = logits.shape
    loss = torch.empty(B, dtype=torch.float32, device=logits.device)
    d = torch.empty_like(logits) if grad_scale is not None else None
    check(load().rxb_softmax_ce(ptr(logits), C, ptr(target), B, C, ptr(loss), ptr(d),
                                float(grad_scale if grad_scale is not None else 0.0), stream_ptr()))
    return loss, d


def sgd_step_(p, grad, mom, lr, momentum=0.9, weight_decay=0.0, nesterov=True, grad_scale=1.0):
    require_gpu()
    for t in (p, grad, mom):
        _cuda(t, torch.float32)
    check(load().rxb_sgd_step(ptr(p), ptr(grad), ptr(mom), p.numel(), lr, momentum, weight_decay,
                              1 if nesterov else 0, grad_scale, stream_ptr()))
    return p


# ------------------------------------------------------------------ family 2: convolutions
def _desc(B, H, W, Cin, ldA, Cout, ldC, c_off, taps, pad, prologue, stats):
    return ConvDesc(B, H, W, Cin, ldA, Cout, ldC, c_off, taps[0], taps[1], pad[0], pad[1],
                    1 if prologue else 0, 1 if stats else 0)


def conv_fwd(A, Wt, Cin=None, scale=None, shift=None, out=None, c_off=0, pad=(0, 0), stats=False):
    """A bf16 [B,H,W,ldA]; Wt bf16 [ty,tx,Cout,Cin] (tap-major).  out bf16 [B,H,W,ldC] written at
    channels c_off..c_off+Cout.  Returns (out, ch_sum, ch_sumsq)."""
    require_gpu()
    A = _cuda(A, torch.bfloat16)
    Wt = _cuda(Wt, torch.bfloat16)
    B, H, W, ldA = A.shape
    ty, tx, Cout, Cin_w = Wt.shape
    Cin = Cin_w if Cin is None else Cin
    if out is None:
        out = torch.zeros(B, H, W, Cout, dtype=torch.bfloat16, device=A.device)
    ldC = out.shape[-1]
    cs = cq = None
    if stats:
        cs = torch.zeros(ldC, dtype=torch.float32, device=A.device)
        cq = torch.zeros(ldC, dtype=torch.float32, device=A.device)
    d = _desc(B, H, W, Cin, ldA, Cout, ldC, c_off, (ty, tx), pad, scale is not None, stats)
    check(load().rxb_conv_fwd(ctypes.byref(d), ptr(A), ptr(Wt), ptr(scale), ptr(shift), ptr(out), ptr(cs),
                              ptr(cq), stream_ptr()))
    return out, cs, cq


def conv_wgrad(A, dOut, Cin, Cout, taps=(1, 1), pad=(0, 0), scale=None, shift=None):
    """dW f32 [Cout,Cin,ty,tx] (torch OIHW) = sum_p dOut[p,n] * A'(p+tap)[k]."""
    require_gpu()
    A = _cuda(A, torch.bfloat16)
    dOut = _cuda(dOut, torch.bfloat16)
    B, H, W, ldA = A.shape
    ldD = dOut.shape[-1]
    dW = torch.zeros(Cout, Cin, taps[0], taps[1], dtype=torch.float32, device=A.device)
    d = _desc(B, H, W, Cin, ldA, Cout, ldD, 0, taps, pad, scale is not None, False)
    check(load().rxb_conv_wgrad(ctypes.byref(d), ptr(A), ptr(scale), ptr(shift), ptr(dOut), ldD, ptr(dW),
                                stream_ptr()))
    return dW


OUT_DY, OUT_G_WRITE, OUT_G_ACCUM = 0, 1, 2


def conv_dgrad_bn(dOut, Wt, X, bn_scale, bn_shift, Cout, out_mode=OUT_DY, out=None, Cin=None, pad=(0, 0),
                  bn_gamma=None, bn_beta=None, wgrad=False):
    """Data gradient fused with the ReLU/BatchNorm backward of the layer that produced the conv input.
    dOut bf16 [B,H,W,ldD]; Wt bf16 [ty,tx,Cout,Cin] = the dgrad operand (tap-flipped, transposed weights);
    X bf16 [B,H,W,ldX] raw activation whose relu(bn(.)) fed the conv (channels 0..Cout).
    Returns (out bf16 [B,H,W,ldC], sum_dy f32 [Cout]); the second BatchNorm-backward reduction comes from
    bn_sum_dyx_from_wdw.  With bn_gamma / bn_beta (the BatchNorm's weight and bias, f32 [Cout]) the channels the
    library flags as degenerate get direct reductions and a third value is returned: sum_dyx f32 [Cout] (zero for
    the channels that were not flagged).  wgrad=True (1x1, Cin <= 128): the same launch also accumulates the forward
    convolution's OIHW weight gradient dW f32 [Cin, Cout] (= sum_p dOut[p,k] * relu(bn(X))[p,c]), returned last; the
    dense layers' 3x3 (Cin 32 -> Cout 128, pad 1, H > 8, W > 4) too: dW f32 [32, 128, 3, 3]."""
    require_gpu()
    dOut = _cuda(dOut, torch.bfloat16)
    Wt = _cuda(Wt, torch.bfloat16)
    X = _cuda(X, torch.bfloat16)
    B, H, W, ldD = dOut.shape
    ty, tx, Cout_w, Cin_w = Wt.shape
    Cin = Cin_w if Cin is None else Cin
    if out is None:
        out = torch.zeros(B, H, W, Cout, dtype=torch.bfloat16, device=dOut.device)
    ldC = out.shape[-1]
    s1 = torch.zeros(Cout, dtype=torch.float32, device=dOut.device)
    d = _desc(B, H, W, Cin, ldD, Cout, ldC, 0, (ty, tx), pad, False, True)
    if wgrad:
        dW = torch.zeros((Cin, Cout) if ty * tx == 1 else (Cin, Cout, ty, tx), dtype=torch.float32, device=dOut.device)
        s2 = torch.zeros(Cout, dtype=torch.float32, device=dOut.device) if bn_gamma is not None else None
        check(load().rxb_conv_dgrad_bn_wgrad(ctypes.byref(d), ptr(dOut), ptr(Wt), ptr(X), X.shape[-1],
                                             ptr(_cuda(bn_scale)), ptr(_cuda(bn_shift)),
                                             ptr(_cuda(bn_gamma)) if bn_gamma is not None else None,
                                             ptr(_cuda(bn_beta)) if bn_beta is not None else None, out_mode,
                                             ptr(out), ptr(s1), ptr(s2), ptr(dW), stream_ptr()))
        return (out, s1, s2, dW) if s2 is not None else (out, s1, dW)
    if bn_gamma is not None:
        s2 = torch.zeros(Cout, dtype=torch.float32, device=dOut.device)
        check(load().rxb_conv_dgrad_bn_ex(ctypes.byref(d), ptr(dOut), ptr(Wt), ptr(X), X.shape[-1],
                                          ptr(_cuda(bn_scale)), ptr(_cuda(bn_shift)), ptr(_cuda(bn_gamma)),
                                          ptr(_cuda(bn_beta)), out_mode, ptr(out), ptr(s1), ptr(s2), stream_ptr()))
        return out, s1, s2
    check(load().rxb_conv_dgrad_bn(ctypes.byref(d), ptr(dOut), ptr(Wt), ptr(X), X.shape[-1], ptr(_cuda(bn_scale)),
                                   ptr(_cuda(bn_shift)), out_mode, ptr(out), ptr(s1), stream_ptr()))
    return out, s1


def conv_dgrad3x3_bn_wgrad_fixup(G, Xc, c0, mean, rstd, corrA, corrB, Wt, X, bn_scale, bn_shift, out=None):
    """The dense layers' 3x3 data + weight gradient with dOut derived on load from the block's concat buffers:
    dOut = G[..., c0:c0+32] - corrA - xhat*corrB (packed bf16, see rxb.h).  G, Xc bf16 [B,H,W,ld]; Wt bf16 [3,3,128,32];
    X bf16 [B,H,W,ldX] (the bottleneck activation).  Returns (out bf16 [B,H,W,128] = dy, sum_dy, dW f32 [32,128,3,3])."""
    require_gpu()
    G, Xc, Wt, X = (_cuda(t, torch.bfloat16) for t in (G, Xc, Wt, X))
    B, H, W, ld = G.shape
    if out is None:
        out = torch.zeros(B, H, W, 128, dtype=torch.bfloat16, device=G.device)
    s1 = torch.zeros(128, dtype=torch.float32, device=G.device)
    dW = torch.zeros(32, 128, 3, 3, dtype=torch.float32, device=G.device)
    d = _desc(B, H, W, 32, 32, 128, out.shape[-1], 0, (3, 3), (1, 1), False, True)
    check(load().rxb_conv_dgrad3x3_bn_wgrad_fixup(ctypes.byref(d), ptr(G), ptr(Xc), ld, c0, ptr(_cuda(mean)), ptr(_cuda(rstd)),
                                                  ptr(_cuda(corrA)), ptr(_cuda(corrB)), ptr(Wt), ptr(X), X.shape[-1],
                                                  ptr(_cuda(bn_scale)), ptr(_cuda(bn_shift)), OUT_DY, ptr(out), ptr(s1),
                                                  ptr(dW), stream_ptr()))
    return out, s1, dW


def bn_sum_dyx_from_wdw(W, dW, bn_scale, bn_shift, sum_dy):
    """sum_p dy*x per input channel of a conv from its fp32 OIHW weights and finished weight gradient."""
    require_gpu()
    W = _cuda(W, torch.float32)
    dW = _cuda(dW, torch.float32)
    Cout, Cin = W.shape[0], W.shape[1]
    taps = W.numel() // (Cout * Cin)
    out = torch.empty(Cin, dtype=torch.float32, device=W.device)
    check(load().rxb_bn_sum_dyx_from_wdw(ptr(W), ptr(dW), Cout, Cin, taps, ptr(_cuda(bn_scale)), ptr(_cuda(bn_shift)),
                                         ptr(_cuda(sum_dy)), ptr(out), stream_ptr()))
    return out
