"""Micro-benchmark of single conv launches (development aid): time, TFLOP/s and effective GB/s."""
import sys, os, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recursion_cellular_image_classification_b200 import ops

dev = torch.device("cuda")

def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

def run(B, H, W, Cin, ldA, Cout, k, prologue, stats):
    A = torch.randn(B, H, W, ldA, device=dev).to(torch.bfloat16)
    Wt = (torch.randn(k, k, Cout, Cin, device=dev) * 0.05).to(torch.bfloat16)
    sc = torch.rand(Cin, device=dev) + 0.5 if prologue else None
    sh = torch.randn(Cin, device=dev) * 0.1 if prologue else None
    out = torch.empty(B, H, W, Cout, device=dev, dtype=torch.bfloat16)
    pad = {1: 0, 3: 1, 4: 2}[k]
    ms = timeit(lambda: ops.conv_fwd(A, Wt, Cin=Cin, scale=sc, shift=sh, out=out, pad=(pad, pad), stats=stats))
    M = B * H * W
    fl = 2.0 * M * Cout * Cin * k * k
    by = M * (Cin + Cout) * 2
    print("fwd  B%d %dx%d Cin%d(ld%d) Cout%d k%d pro%d stats%d : %.3f ms  %.1f TF/s  %.0f GB/s(min traffic)" %
          (B, H, W, Cin, ldA, Cout, k, prologue, stats, ms, fl / ms / 1e9, by / ms / 1e6))

def run_wgrad(B, H, W, Cin, ldA, Cout, k, prologue):
    A = torch.randn(B, H, W, ldA, device=dev).to(torch.bfloat16)
    dO = torch.randn(B, H, W, Cout, device=dev).to(torch.bfloat16)
    sc = torch.rand(Cin, device=dev) + 0.5 if prologue else None
    sh = torch.randn(Cin, device=dev) * 0.1 if prologue else None
    pad = {1: 0, 3: 1, 4: 2}[k]
    ms = timeit(lambda: ops.conv_wgrad(A, dO, Cin, Cout, taps=(k, k), pad=(pad, pad), scale=sc, shift=sh))
    M = B * H * W
    fl = 2.0 * M * Cout * Cin * k * k
    by = M * (Cin + Cout) * 2
    print("wgrd B%d %dx%d Cin%d(ld%d) Cout%d k%d pro%d : %.3f ms  %.1f TF/s  %.0f GB/s(min traffic)" %
          (B, H, W, Cin, ldA, Cout, k, prologue, ms, fl / ms / 1e9, by / ms / 1e6))

def run_dgrad(B, H, W, Cd, Cx, ldX, k, out_mode, wgrad=0):
    dO = torch.randn(B, H, W, Cd, device=dev).to(torch.bfloat16)
    Wt = (torch.randn(k, k, Cx, Cd, device=dev) * 0.05).to(torch.bfloat16)
    X = torch.randn(B, H, W, ldX, device=dev).to(torch.bfloat16)
    sc = torch.rand(Cx, device=dev) + 0.5
    sh = torch.randn(Cx, device=dev) * 0.1
    out = torch.zeros(B, H, W, ldX if out_mode else Cx, device=dev, dtype=torch.bfloat16)
    pad = {1: 0, 3: 1}[k]
    ms = timeit(lambda: ops.conv_dgrad_bn(dO, Wt, X, sc, sh, Cx, out_mode=out_mode, out=out, pad=(pad, pad),
                                          wgrad=bool(wgrad)))
    M = B * H * W
    fl = 2.0 * M * Cx * Cd * k * k * (2 if wgrad else 1)
    by = M * (Cd + Cx * (3 if out_mode == 2 else 2)) * 2
    print("dgrd B%d %dx%d Cd%d Cx%d(ld%d) k%d mode%d wgrad%d : %.3f ms  %.1f TF/s  %.0f GB/s(min traffic)" %
          (B, H, W, Cd, Cx, ldX, k, out_mode, wgrad, ms, fl / ms / 1e9, by / ms / 1e6))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        # one shape, e.g. for RXB_DBG_TIMELINE=1:  bench_conv.py one fwd|wgrad|dgrad <args of run / run_wgrad / run_dgrad>
        kind, a = sys.argv[2], [int(v) for v in sys.argv[3:]]
        {"fwd": run, "wgrad": run_wgrad, "dgrad": run_dgrad}[kind](*a)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "prof":
        # one launch of each representative block-1 shape (for ncu); optional batch as argv[2]
        B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
        timeit_ = timeit
        globals()["timeit"] = lambda fn, reps=1: (fn(), torch.cuda.synchronize(), 0.001)[2]
        run(B, 128, 128, 128, 128, 32, 3, 1, 1)
        run(B, 128, 128, 224, 256, 128, 1, 1, 1)
        run_dgrad(B, 128, 128, 32, 128, 128, 3, 0)
        run_dgrad(B, 128, 128, 128, 224, 256, 1, 2)
        run_wgrad(B, 128, 128, 128, 128, 32, 3, 1)
        run_wgrad(B, 128, 128, 224, 256, 128, 1, 1)
        run_dgrad(B, 128, 128, 32, 128, 128, 3, 0, 1)      # the fused data+weight-gradient launches the step runs
        run_dgrad(B, 128, 128, 128, 224, 256, 1, 2, 1)
        sys.exit(0)
    B = 64
    for pro, st in itertools.product((0, 1), (0, 1)):
        run(B, 128, 128, 128, 128, 32, 3, pro, st)
    for pro, st in itertools.product((0, 1), (0, 1)):
        run(B, 128, 128, 128, 256, 128, 1, pro, st)
    run(B, 128, 128, 224, 256, 128, 1, 1, 1)
    run(B, 64, 64, 128, 128, 32, 3, 1, 1)
    run(B, 64, 64, 256, 512, 128, 1, 1, 1)
    run(B, 32, 32, 512, 1024, 128, 1, 1, 1)
    run(B, 256, 256, 32, 32, 64, 4, 0, 1)
    run(B, 128, 128, 32, 32, 128, 3, 0, 0)     # dgrad 3x3 shape (EPI_STORE flavour)
    run(B, 128, 128, 128, 128, 224, 1, 0, 0)   # dgrad 1x1 shape (EPI_STORE flavour)
    for pro in (0, 1):
        run_wgrad(B, 128, 128, 128, 128, 32, 3, pro)
        run_wgrad(B, 128, 128, 224, 256, 128, 1, pro)
    run_wgrad(B, 256, 256, 32, 32, 64, 4, 0)
    run_dgrad(B, 128, 128, 32, 128, 128, 3, 0)
    run_dgrad(B, 128, 128, 128, 224, 256, 1, 2)
    run_dgrad(B, 64, 64, 128, 480, 512, 1, 2)
    run_dgrad(B, 32, 32, 128, 992, 1024, 1, 2)
