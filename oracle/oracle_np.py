"""CPU oracle — TEST INFRASTRUCTURE ONLY.

A numpy / OpenCV / torch-fp32 restatement of the reference algorithms on the hot path.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module; the
product package never does (it fails loudly when librxb.so or the GPU is missing).

Parity status
  * compute_mean_std_arrays, mask_rescale, greedy_assign: PINNED — checked against outputs of the
    reference's own functions (compute_stats_experiments.compute_mean_std, cell_classifier.test.test)
    run in the build container; fixtures under tests/golden/ (tests/golden/make_golden.py).
  * warp_affine_u8 (ShiftScaleRotate at any angle): restates cv2.warpAffine's fixed-point bilinear remap; PINNED
    against cv2.warpAffine itself (OpenCV 4.13, the library call albumentations makes) executed in the tests and
    against tests/golden/warp_golden.npz; rotation_matrix is pinned bitwise against cv2.getRotationMatrix2D.
  * jpeg_decode_gray (baseline grayscale JPEG, dataloader.py:141-146): restates libjpeg's marker parsing, Huffman
    decoding and jpeg_idct_islow; PINNED bit-exact against cv2.imdecode (OpenCV 4.13 / libjpeg-turbo 3.1.2, the
    reference's own call) executed in the tests and against tests/golden/jpeg_golden.npz.
  * d4_augment / normalize: restate albumentations==0.3.0 (requirement.txt:1), which is NOT installed
    and has no tests in the reference -> "parity unpinned" for that third-party boundary; the OpenCV
    calls it makes (cv2.flip, cv2.warpAffine) are executed for real here.
  * the stem recipe and two_sites_features: PINNED against the reference's own TwoSitesNN (models.py:8-57) built in
    the build container under a seed (tests/golden/model_golden.npz).
  * densenet121_6ch: torchvision's densenet121 with the reference's 6-channel stem recipe
    (models.py:17-27); the north star's trunk, fp32 on CPU.
"""
import numpy as np


# ------------------------------------------------------------------------------------------------
# S1: compute_stats_experiments.py:8-24 on decoded planes instead of files.
def compute_mean_std_arrays(planes, mean=None, std=None):
    """planes: uint8 [n, C, H, W].  Returns (mean, std) float64 [C].
    Follows compute_mean_std line by line: im = u8/255 (f64); optional pre-normalisation (:16-17);
    count/sum/sum-of-squares per channel (:18-20); mean = sum/n, std = sqrt(sumsq/n - mean^2) (:21-23)."""
    n, C, H, W = planes.shape
    count = np.zeros(C)
    sum_x = np.zeros(C)
    sum_x2 = np.zeros(C)
    for i in range(n):
        for ch in range(C):
            im = planes[i, ch] / 255
            if (mean is not None) and (std is not None):
                im = (im - mean[ch]) / std[ch]
            count[ch] += 1
            sum_x[ch] += np.sum(im)
            sum_x2[ch] += np.sum(im ** 2)
    count = count * H * W
    m = sum_x / count
    s = np.sqrt((sum_x2 / count) - m ** 2)
    return m, s


# ------------------------------------------------------------------------------------------------
# L2: dataloader.py:141-146 — cv2.imdecode(buffer, -1) of a single-channel baseline JPEG (png_to_jpeg.py:11-15),
# i.e. libjpeg: jdmarker.c (segments), jdhuff.c (canonical Huffman codes, DC prediction, run/size AC symbols, 0xFF00
# unstuffing, restart intervals), jidctint.c jpeg_idct_islow (CONST_BITS 13, PASS1_BITS 2) and the range-limit table.
ZIGZAG_TO_NATURAL = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34,
                              27, 20, 13, 6, 7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44,
                              51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63])


class JpegError(ValueError):
    pass


def _jpeg_segments(buf):
    """Yield (marker, payload) up to and including SOS, then ('scan', entropy-coded bytes)."""
    if buf[:2] != b"\xff\xd8":
        raise JpegError("no SOI")
    p = 2
    while True:
        if p + 4 > len(buf):
            raise JpegError("ran off the end before SOS")
        if buf[p] != 0xFF or buf[p + 1] == 0xFF:
            p += 1
            continue
        m = buf[p + 1]
        if m in (0x00, 0x01) or 0xD0 <= m <= 0xD8:
            p += 2
            continue
        n = (buf[p + 2] << 8) | buf[p + 3]
        yield m, buf[p + 4:p + 2 + n]
        p += 2 + n
        if m == 0xDA:
            yield "scan", buf[p:]
            return


def _huff_codes(counts, symbols):
    """(length, code) -> symbol for the canonical code of a DHT table."""
    table, code, k = {}, 0, 0
    for length in range(1, 17):
        for _ in range(counts[length - 1]):
            table[(length, code)] = symbols[k]
            code += 1
            k += 1
        code <<= 1
    return table


def _idct_islow(blocks):
    """jpeg_idct_islow on dequantised coefficients, int64 [n,8,8] -> uint8 [n,8,8]."""
    F = dict(f0298=2446, f0390=3196, f0541=4433, f0765=6270, f0899=7373, f1175=9633, f1501=12299, f1847=15137,
             f1961=16069, f2053=16819, f2562=20995, f3072=25172)

    def pass_1d(v, shift):          # v: [..., 8] along the transformed axis
        z2, z3 = v[..., 2], v[..., 6]
        z1 = (z2 + z3) * F["f0541"]
        tmp2 = z1 + z3 * (-F["f1847"])
        tmp3 = z1 + z2 * F["f0765"]
        z2, z3 = v[..., 0], v[..., 4]
        tmp0, tmp1 = (z2 + z3) << 13, (z2 - z3) << 13
        tmp10, tmp13, tmp11, tmp12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
        tmp0, tmp1, tmp2, tmp3 = v[..., 7], v[..., 5], v[..., 3], v[..., 1]
        z1, z2, z3, z4 = tmp0 + tmp3, tmp1 + tmp2, tmp0 + tmp2, tmp1 + tmp3
        z5 = (z3 + z4) * F["f1175"]
        tmp0, tmp1, tmp2, tmp3 = tmp0 * F["f0298"], tmp1 * F["f2053"], tmp2 * F["f3072"], tmp3 * F["f1501"]
        z1, z2, z3, z4 = z1 * -F["f0899"], z2 * -F["f2562"], z3 * -F["f1961"] + z5, z4 * -F["f0390"] + z5
        tmp0, tmp1, tmp2, tmp3 = tmp0 + z1 + z3, tmp1 + z2 + z4, tmp2 + z2 + z3, tmp3 + z1 + z4
        out = np.stack([tmp10 + tmp3, tmp11 + tmp2, tmp12 + tmp1, tmp13 + tmp0,
                        tmp13 - tmp0, tmp12 - tmp1, tmp11 - tmp2, tmp10 - tmp3], axis=-1)
        return (out + (1 << (shift - 1))) >> shift

    ws = pass_1d(np.swapaxes(blocks.astype(np.int64), 1, 2), 13 - 2)        # columns
    out = pass_1d(np.swapaxes(ws, 1, 2), 13 + 2 + 3)                        # rows
    i = out & 1023                                                         # range_limit[x & RANGE_MASK]
    return np.where(i < 128, i + 128, np.where(i < 512, 255, np.where(i < 896, 0, i - 896))).astype(np.uint8)


def jpeg_decode_gray(buf):
    """cv2.imdecode(np.frombuffer(buf, np.uint8), -1) for a baseline, 8-bit, single-component JPEG -> uint8 [H,W]."""
    buf = bytes(buf)
    quant, huff, frame, scan_hdr, restart, scan = {}, {}, None, None, 0, None
    for m, seg in _jpeg_segments(buf):
        if m == 0xDB:
            q = 0
            while q < len(seg):
                pq, tid = seg[q] >> 4, seg[q] & 15
                if pq:
                    vals = [(seg[q + 1 + 2 * i] << 8) | seg[q + 2 + 2 * i] for i in range(64)]
                else:
                    vals = list(seg[q + 1:q + 65])
                quant[tid] = np.array(vals, dtype=np.int64)
                q += 1 + (128 if pq else 64)
        elif m == 0xC4:
            q = 0
            while q + 17 <= len(seg):
                counts = list(seg[q + 1:q + 17])
                total = sum(counts)
                huff[(seg[q] >> 4, seg[q] & 15)] = _huff_codes(counts, seg[q + 17:q + 17 + total])
                q += 17 + total
        elif m in (0xC0, 0xC1):
            if seg[0] != 8 or seg[5] != 1:
                raise JpegError("unsupported: not 8-bit single-component")
            frame = ((seg[1] << 8) | seg[2], (seg[3] << 8) | seg[4], seg[8] & 3)
        elif isinstance(m, int) and 0xC2 <= m <= 0xCF and m not in (0xC4, 0xC8):
            raise JpegError("unsupported: progressive / lossless / arithmetic")
        elif m == 0xDD:
            restart = (seg[0] << 8) | seg[1]
        elif m == 0xDA:
            if seg[0] != 1 or seg[3] != 0 or seg[4] != 63:
                raise JpegError("unsupported scan")
            scan_hdr = (seg[2] >> 4, seg[2] & 15)
        elif m == "scan":
            scan = seg
    H, W, tq = frame
    dc, ac, qz = huff[(0, scan_hdr[0])], huff[(1, scan_hdr[1])], quant[tq]

    # entropy-coded segment -> bits, unstuffed; markers split restart intervals
    state = {"p": 0, "acc": 0, "n": 0, "marker": 0}

    def fill(need):
        while state["n"] < need:
            c = 0
            if not state["marker"]:
                if state["p"] >= len(scan):
                    state["marker"] = 0xD9
                else:
                    c = scan[state["p"]]
                    state["p"] += 1
                    if c == 0xFF:
                        c2 = 0xFF
                        while c2 == 0xFF and state["p"] < len(scan):
                            c2 = scan[state["p"]]
                            state["p"] += 1
                        if c2 == 0xFF:
                            c2 = 0xD9
                        if c2 != 0:
                            state["marker"], c = c2, 0
            state["acc"] = ((state["acc"] << 8) | c) & 0xFFFFFFFFFFFF
            state["n"] += 8

    def bits(k):
        fill(k)
        v = (state["acc"] >> (state["n"] - k)) & ((1 << k) - 1)
        state["n"] -= k
        return v

    def symbol(table):
        code = 0
        for length in range(1, 17):
            code = (code << 1) | bits(1)
            if (length, code) in table:
                return table[(length, code)]
        raise JpegError("invalid Huffman code")

    def extend(v, s):
        return v - (1 << s) + 1 if v < (1 << (s - 1)) else v

    bw, bh = (W + 7) // 8, (H + 7) // 8
    coefs = np.zeros((bw * bh, 64), dtype=np.int64)
    pred = 0
    for blk in range(bw * bh):
        if restart and blk and blk % restart == 0:
            state["acc"], state["n"] = 0, 0
            if not state["marker"]:
                while not (scan[state["p"]] == 0xFF and 0xD0 <= scan[state["p"] + 1] <= 0xD7):
                    state["p"] += 1
                state["p"] += 2
            state["marker"] = 0
            pred = 0
        s = symbol(dc)
        pred += extend(bits(s), s) if s else 0
        coefs[blk, 0] = pred * qz[0]
        k = 1
        while k < 64:
            rs = symbol(ac)
            r, s = rs >> 4, rs & 15
            if s:
                k += r
                coefs[blk, ZIGZAG_TO_NATURAL[k]] = extend(bits(s), s) * qz[k]
                k += 1
            elif r == 15:
                k += 16
            else:
                break
    px = _idct_islow(coefs.reshape(-1, 8, 8))
    img = px.reshape(bh, bw, 8, 8).transpose(0, 2, 1, 3).reshape(bh * 8, bw * 8)
    return np.ascontiguousarray(img[:H, :W])


# ------------------------------------------------------------------------------------------------
# L3-L5: dataloader.py:42-51, 128-139 through albumentations 0.3.0 (restated).
def d4_augment(img_hwc, vflip=False, hflip=False, k=0, ref_compat=False):
    """VerticalFlip -> HorizontalFlip -> rotation by k*90 degrees (counter-clockwise).
    ref_compat=False: canonical D4 (np.rot90).  ref_compat=True: what ShiftScaleRotate does at that
    angle: cv2.warpAffine about (w/2, h/2), INTER_LINEAR, BORDER_REFLECT_101."""
    img = img_hwc
    if vflip:
        img = img[::-1]          # cv2.flip(img, 0)
    if hflip:
        img = img[:, ::-1]       # cv2.flip(img, 1)
    img = np.ascontiguousarray(img)
    if ref_compat:
        import cv2
        h, w = img.shape[:2]
        M = cv2.getRotationMatrix2D((w / 2, h / 2), 90.0 * k, 1.0)
        chans = [cv2.warpAffine(np.ascontiguousarray(img[:, :, c]), M, (w, h), flags=cv2.INTER_LINEAR,
                                borderMode=cv2.BORDER_REFLECT_101) for c in range(img.shape[2])]
        img = np.stack(chans, axis=2)
    else:
        img = np.rot90(img, k)
    return np.ascontiguousarray(img)


def rotation_matrix(w, h, angle, scale=1.0):
    """cv2.getRotationMatrix2D((w/2, h/2), angle, scale) as albumentations 0.3.0 ShiftScaleRotate calls it
    (dataloader.py:45-46): float64 [2,3].  OpenCV scales the angle by the folded constant pi/180."""
    import math
    cx, cy = w / 2, h / 2
    a = angle * (math.pi / 180.0)
    al, be = math.cos(a) * scale, math.sin(a) * scale
    return np.array([[al, be, (1 - al) * cx - be * cy], [-be, al, be * cx + (1 - al) * cy]], dtype=np.float64)


def invert_affine(M):
    """The inversion cv::warpAffine applies to a forward matrix, in its operation order."""
    M = np.array(M, dtype=np.float64).reshape(6).copy()
    D = M[0] * M[4] - M[1] * M[3]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = M[4] * D, M[0] * D
    M[0] = A11
    M[1] *= -D
    M[3] *= -D
    M[4] = A22
    b1 = -M[0] * M[2] - M[1] * M[5]
    b2 = -M[3] * M[2] - M[4] * M[5]
    M[2], M[5] = b1, b2
    return M


def reflect101(p, n):
    """cv::borderInterpolate(p, n, BORDER_REFLECT_101) on an integer array."""
    p = np.array(p, dtype=np.int64)
    if n == 1:
        return np.zeros_like(p)
    while True:
        out = (p < 0) | (p >= n)
        if not out.any():
            return p
        p = np.where(p < 0, -p, p)
        p = np.where(p >= n, 2 * n - 2 - p, p)


def warp_affine_u8(img, M):
    """cv2.warpAffine(img, M, (w, h), flags=INTER_LINEAR, borderMode=BORDER_REFLECT_101) for uint8 HW or HWC,
    restated from OpenCV's fixed-point path: destination (x, y) maps to the source at 1/1024-pixel resolution
    (AB_BITS = 10), is rounded to 1/32 pixel (INTER_BITS = 5, round_delta = 16), and the four reflected taps are
    blended with the 15-bit integer table, which for bilinear weights is exactly 32*(32-fx|fx)*(32-fy|fy), so
    out = (sum w*p + 512) >> 10."""
    H, W = img.shape[:2]
    Mi = invert_affine(M)
    x = np.arange(W, dtype=np.float64)
    y = np.arange(H, dtype=np.float64)
    adelta = np.rint(Mi[0] * x * 1024).astype(np.int64)
    bdelta = np.rint(Mi[3] * x * 1024).astype(np.int64)
    X0 = np.rint((Mi[1] * y + Mi[2]) * 1024).astype(np.int64) + 16
    Y0 = np.rint((Mi[4] * y + Mi[5]) * 1024).astype(np.int64) + 16
    X = (X0[:, None] + adelta[None, :]) >> 5
    Y = (Y0[:, None] + bdelta[None, :]) >> 5
    sx, sy = np.clip(X >> 5, -32768, 32767), np.clip(Y >> 5, -32768, 32767)
    fx, fy = X & 31, Y & 31
    xa, xb = reflect101(sx, W), reflect101(sx + 1, W)
    ya, yb = reflect101(sy, H), reflect101(sy + 1, H)
    im = img.astype(np.int64).reshape(H, W, -1)
    acc = (((32 - fx) * (32 - fy))[..., None] * im[ya, xa] + (fx * (32 - fy))[..., None] * im[ya, xb] +
           ((32 - fx) * fy)[..., None] * im[yb, xa] + (fx * fy)[..., None] * im[yb, xb])
    return ((acc + 512) >> 10).astype(np.uint8).reshape(img.shape)


def shift_scale_rotate(img_hwc, angle, use_cv2=False):
    """albumentations 0.3.0 ShiftScaleRotate(shift_limit=0, scale_limit=0) at a given angle (dataloader.py:45-46)."""
    h, w = img_hwc.shape[:2]
    M = rotation_matrix(w, h, angle)
    if use_cv2:
        import cv2
        return cv2.warpAffine(np.ascontiguousarray(img_hwc), M, (w, h), flags=cv2.INTER_LINEAR,
                              borderMode=cv2.BORDER_REFLECT_101)
    return warp_affine_u8(img_hwc, M)


def transform_affine(img_chw_u8, mean, std, vflip=False, hflip=False, angle=0.0, crop_yx=(0, 0), out_hw=None,
                     use_cv2=False):
    """The reference's full train transform (dataloader.py:42-48, 128-139) with explicit parameters: flips ->
    ShiftScaleRotate(angle) -> crop -> Normalize.  Returns float32 CHW."""
    img = np.moveaxis(img_chw_u8, 0, 2)
    if vflip:
        img = img[::-1]
    if hflip:
        img = img[:, ::-1]
    img = shift_scale_rotate(np.ascontiguousarray(img), angle, use_cv2)
    if out_hw is not None:
        img = crop(img, crop_yx[0], crop_yx[1], out_hw[0], out_hw[1])
    img = normalize(img, mean, std)
    return np.ascontiguousarray(np.moveaxis(img, 2, 0))


def crop(img_hwc, y0, x0, h, w):
    return img_hwc[y0:y0 + h, x0:x0 + w]


def normalize_constants(mean, std):
    m = np.asarray(mean, dtype=np.float32) * np.float32(255.0)
    s = np.asarray(std, dtype=np.float32) * np.float32(255.0)
    return m, np.reciprocal(s, dtype=np.float32)


def normalize(img_hwc_u8, mean, std):
    """albumentations.Normalize(mean, std, max_pixel_value=255): float32 throughout."""
    m, d = normalize_constants(mean, std)
    img = img_hwc_u8.astype(np.float32)
    img -= m
    img *= d
    return img


def transform(img_chw_u8, mean, std, vflip=False, hflip=False, k=0, crop_yx=(0, 0), out_hw=None,
              ref_compat=False):
    """ImagesDS._transform (dataloader.py:128-139) with explicit augmentation parameters.
    Returns float32 CHW."""
    img = np.moveaxis(img_chw_u8, 0, 2)
    img = d4_augment(img, vflip, hflip, k, ref_compat)
    if out_hw is not None:
        img = crop(img, crop_yx[0], crop_yx[1], out_hw[0], out_hw[1])
    img = normalize(img, mean, std)
    return np.ascontiguousarray(np.moveaxis(img, 2, 0))


def to_nhwc8_bf16(x_chw_f32):
    """float32 [6,H,W] -> the loader's bf16 NHWC8 as float32 values (channels 6,7 zero)."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(np.moveaxis(x_chw_f32, 0, 2)))
    t = torch.cat([t, torch.zeros(*t.shape[:2], 2)], dim=2)
    return t.to(torch.bfloat16).float().numpy()


def to_s2d32(x_hwc8):
    """[H,W,8] -> [H/2,W/2,32] with channel = (y&1)*16 + (x&1)*8 + c."""
    H, W, C = x_hwc8.shape
    t = x_hwc8.reshape(H // 2, 2, W // 2, 2, C)
    return np.ascontiguousarray(np.transpose(t, (0, 2, 1, 3, 4)).reshape(H // 2, W // 2, 4 * C))


# ------------------------------------------------------------------------------------------------
# T3: test.py:27, 34-56
def softmax(x):
    x = x - x.max(axis=1, keepdims=True)
    e = np.exp(x)
    return e / e.sum(axis=1, keepdims=True)


def rescale(preds):
    """test.py:34-39, verbatim semantics."""
    temp = np.sum(preds, axis=1)
    temp[temp == 0] = 1
    temp = np.repeat(temp[:, np.newaxis], preds.shape[1], axis=1)
    return preds / temp


def mask_rescale(preds, plate_groups_col, plates):
    """test.py:42-46: zero the classes whose plate group differs from the row's plate, then rescale."""
    preds = preds.copy()
    mask = np.repeat(plate_groups_col[np.newaxis, :], len(preds), axis=0) != \
        np.repeat(np.asarray(plates)[:, np.newaxis], preds.shape[1], axis=1)
    preds[mask] = 0
    return rescale(preds)


def greedy_assign(preds):
    """test.py:48-56."""
    preds = preds.copy()
    results = np.zeros(preds.shape[0])
    for _ in range(preds.shape[0]):
        max_per_row_idx = np.argmax(preds, axis=1)
        max_row_idx = np.argmax(preds[np.arange(len(preds)), max_per_row_idx])
        max_column_idx = max_per_row_idx[max_row_idx]
        results[max_row_idx] = max_column_idx
        preds[:, max_column_idx] = 0
        preds[max_row_idx, :] = 0
        preds = rescale(preds)
    return results


def pairwise_sum_f32(a):
    """numpy's float32 pairwise summation (loops_utils.h.src *_pairwise_sum), restated so the CUDA kernel's
    association order can be checked against np.sum itself."""
    a = np.asarray(a, dtype=np.float32)
    n = a.shape[0]
    f = np.float32
    if n < 8:
        res = f(-0.0)
        for i in range(n):
            res = f(res + a[i])
        return res
    if n <= 128:
        r = [f(a[j]) for j in range(8)]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] = f(r[j] + a[i + j])
            i += 8
        res = f(f(f(r[0] + r[1]) + f(r[2] + r[3])) + f(f(r[4] + r[5]) + f(r[6] + r[7])))
        while i < n:
            res = f(res + a[i])
            i += 1
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return f(pairwise_sum_f32(a[:n2]) + pairwise_sum_f32(a[n2:]))


# ------------------------------------------------------------------------------------------------
# M1-M3: models.py
def two_sites_features(features, bs):
    """models.py:46-53 given trunk features [bs*G, F] (torch): reshape to [bs, G, F], split G into the
    image / negative-control / positive-control thirds, mean over the sites of each third, concat."""
    import torch
    features = features.reshape([bs, -1, features.shape[1]])
    shape = int(features.shape[1] / 3)
    f_img = features[:, 0:shape, :].mean(1)
    f_neg = features[:, shape:2 * shape, :].mean(1)
    f_pos = features[:, 2 * shape:, :].mean(1)
    return torch.cat([f_img, f_neg, f_pos], dim=1)


def densenet121_6ch(num_classes=1108, seed=0):
    """torchvision densenet121 with the reference's stem surgery (models.py:17-27): a 6-channel 7x7/2
    conv whose filters are the channel-mean of the 3-channel stem, replicated 6 times."""
    import torch
    import torch.nn as nn
    from torchvision import models
    torch.manual_seed(seed)
    net = models.densenet121(weights=None, num_classes=num_classes)
    trained_kernel = net.features.conv0.weight
    new_conv = nn.Conv2d(6, 64, kernel_size=7, stride=2, padding=3, bias=False)
    with torch.no_grad():
        new_conv.weight[:, :] = torch.stack([torch.mean(trained_kernel, 1)] * 6, dim=1)
    net.features.conv0 = new_conv
    return net


def resnet18_6ch(num_classes=1108, seed=0):
    """BASELINE config 1: ResNet-18-style 6-channel network with the same stem recipe."""
    import torch
    import torch.nn as nn
    from torchvision import models
    torch.manual_seed(seed)
    net = models.resnet18(weights=None, num_classes=num_classes)
    trained_kernel = net.conv1.weight
    new_conv = nn.Conv2d(6, 64, kernel_size=7, stride=2, padding=3, bias=False)
    with torch.no_grad():
        new_conv.weight[:, :] = torch.stack([torch.mean(trained_kernel, 1)] * 6, dim=1)
    net.conv1 = new_conv
    return net


def two_sites_resnet50(nb_classes=1108, size_features=1024, dropout=0.3, seed=0):
    """The reference's real model surface (models.py:8-57) restated: torchvision resnet50 trunk with the
    6-channel stem recipe and fc -> Identity, site/control feature averaging (two_sites_features) and the
    BatchNorm1d -> Dropout -> Linear -> ReLU -> BatchNorm1d -> Dropout -> Linear head.  Module construction order
    follows models.py so that a seed gives the reference's initial weights."""
    import torch
    import torch.nn as nn
    from torchvision import models

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.base_nn = models.resnet50(weights=None)
            trained_kernel = self.base_nn.conv1.weight
            new_conv = nn.Conv2d(6, 64, kernel_size=7, stride=2, padding=3, bias=False)
            with torch.no_grad():
                new_conv.weight[:, :] = torch.stack([torch.mean(trained_kernel, 1)] * 6, dim=1)
            self.base_nn.conv1 = new_conv
            n_feat = 3 * self.base_nn.fc.in_features
            self.base_nn.fc = nn.Identity()
            self.mlp = nn.Sequential(nn.BatchNorm1d(n_feat), nn.Dropout(dropout), nn.Linear(n_feat, size_features),
                                     nn.ReLU(), nn.BatchNorm1d(size_features), nn.Dropout(dropout),
                                     nn.Linear(size_features, nb_classes))

        def forward(self, x):
            bs = x.shape[0]
            feats = self.base_nn(x.reshape([-1, x.shape[2], x.shape[3], x.shape[4]]))
            return self.mlp(two_sites_features(feats, bs))

    torch.manual_seed(seed)
    return Net()


def sgd_reference(params, lr, momentum=0.9, nesterov=True, weight_decay=3e-5):
    """main.py:89-93."""
    import torch
    return torch.optim.SGD(params, lr=lr, momentum=momentum, nesterov=nesterov, weight_decay=weight_decay)
