#!/usr/bin/env python
"""One-process GPU check of the widening rows: runs their parity tests, then times the two kernels.

    python tools/widen_check.py [--no-tests] > gpurun_out/widen.json

Timing: CUDA events on the launching stream, 3 warm-ups, inputs larger than the 126 MB L2.  The JSON line goes to
stdout, everything else to stderr."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def timed(fn, reps, warm=3):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    out = {}
    saved = os.dup(1)
    os.dup2(2, 1)
    if "--no-tests" not in sys.argv:
        import pytest
        t0 = time.time()
        out["pytest_rc"] = int(pytest.main(["-q", "-m", "gpu", "-p", "no:cacheprovider",
                                            os.path.join(ROOT, "tests", "test_gpu_widen_loader_affine.py"),
                                            os.path.join(ROOT, "tests", "test_gpu_widen_jpeg.py"),
                                            os.path.join(ROOT, "tests", "test_gpu_shims.py")]))
        out["pytest_s"] = time.time() - t0
    try:
        import cv2
        import numpy as np
        import torch
        from recursion_cellular_image_classification_b200 import ops
        from recursion_cellular_image_classification_b200.synth import synth_planes, synth_planes_torch
        dev = torch.device("cuda:0")
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        B, S = 128, 512
        src = synth_planes_torch(3, B, dev)
        idx = torch.arange(B, dtype=torch.int32, device=dev)
        exp = torch.zeros(B, dtype=torch.int32, device=dev)
        crops = torch.zeros(B, 2, dtype=torch.int32, device=dev)
        rng = np.random.default_rng(0)
        m, d = ops.normalize_constants(np.full((1, 6), 0.1), np.full((1, 6), 0.08))
        m, d = torch.from_numpy(m).to(dev), torch.from_numpy(d).to(dev)
        mats = torch.from_numpy(np.stack([ops.rotation_matrix(S, S, float(a)) for a in rng.uniform(-180, 180, B)])).to(dev)
        flips = torch.from_numpy(rng.integers(0, 4, B).astype(np.uint8)).to(dev)
        codes = torch.from_numpy(rng.integers(0, 16, B).astype(np.uint8)).to(dev)
        dst = torch.empty(B, S // 2, S // 2, 32, dtype=torch.bfloat16, device=dev)
        bytes_alg = 3 * 6 * S * S * B
        ms = timed(lambda: ops.load_norm_affine(src, idx, exp, flips, mats, crops, m, d, (S, S), ops.OUT_BF16_S2D32,
                                                out=dst), 20)
        out["loader_affine"] = {"ms": ms, "images": B, "achieved_gbs": bytes_alg / ms / 1e6, "peak_gbs": peak,
                                "frac": bytes_alg / ms / 1e6 / peak, "algorithmic_bytes": bytes_alg}
        ms = timed(lambda: ops.load_norm_aug(src, idx, exp, codes, crops, m, d, (S, S), ops.OUT_BF16_S2D32, out=dst), 20)
        out["loader_d4"] = {"ms": ms, "images": B, "achieved_gbs": bytes_alg / ms / 1e6, "frac": bytes_alg / ms / 1e6 / peak}

        planes = synth_planes(6, n=2)
        bufs = [cv2.imencode(".jpg", planes[i, c], [cv2.IMWRITE_JPEG_QUALITY, 95])[1].tobytes()
                for i in range(2) for c in range(6)]
        t0 = time.perf_counter()
        for b_ in bufs * 2:
            cv2.imdecode(np.frombuffer(b_, np.uint8), -1)
        cpu_ms = (time.perf_counter() - t0) / (2 * len(bufs)) * 1e3
        files = bufs * 64                                                  # 768 files = 128 six-channel images
        blob, offsets = ops.pack_jpeg_buffers(files)
        blob, offsets = blob.to(dev), offsets.to(dev)
        planes_out = torch.empty(len(files), S, S, dtype=torch.uint8, device=dev)
        ref = torch.from_numpy(np.stack([cv2.imdecode(np.frombuffer(b_, np.uint8), -1) for b_ in bufs])).to(dev)
        for name, par in (("jpeg_decode", True), ("jpeg_decode_single_lane", False)):
            planes_out.zero_()
            ms = timed(lambda: ops.jpeg_decode_gray(blob, offsets, (S, S), out=planes_out, check_status=False,
                                                    parallel=par), 5, warm=2)
            out[name] = {"ms": ms, "files": len(files), "files_per_s": len(files) / ms * 1e3,
                         "images_per_s": len(files) / 6 / ms * 1e3, "compressed_mb": blob.numel() / 1e6,
                         "compressed_gbs": blob.numel() / ms / 1e6, "output_gbs": planes_out.numel() / ms / 1e6,
                         "bit_exact_768": bool((planes_out.view(64, 12, S, S) == ref[None]).all().item())}
        out["jpeg_cpu_cv2_imdecode"] = {"ms_per_file_1_thread": cpu_ms, "files_per_s_1_thread": 1e3 / cpu_ms,
                                        "cores": os.cpu_count()}
        # a smoother corpus (blurred planes: far fewer bits per block, like sparse fluorescence fields)
        sm = [cv2.imencode(".jpg", cv2.GaussianBlur(planes[i, c], (0, 0), 2.0), [cv2.IMWRITE_JPEG_QUALITY, 95])[1].tobytes()
              for i in range(2) for c in range(6)]
        sblob, soff = ops.pack_jpeg_buffers(sm * 64)
        sblob, soff = sblob.to(dev), soff.to(dev)
        sref = torch.from_numpy(np.stack([cv2.imdecode(np.frombuffer(b_, np.uint8), -1) for b_ in sm])).to(dev)
        for name, par in (("jpeg_decode_smooth", True), ("jpeg_decode_smooth_single_lane", False)):
            planes_out.zero_()
            ms = timed(lambda: ops.jpeg_decode_gray(sblob, soff, (S, S), out=planes_out, check_status=False,
                                                    parallel=par), 5, warm=2)
            out[name] = {"ms": ms, "files": 768, "files_per_s": 768 / ms * 1e3, "compressed_mb": sblob.numel() / 1e6,
                         "bit_exact_768": bool((planes_out.view(64, 12, S, S) == sref[None]).all().item())}
    except Exception as e:      # report what ran
        out["bench_error"] = repr(e)
    os.write(saved, (json.dumps(out) + "\n").encode())


if __name__ == "__main__":
    main()
