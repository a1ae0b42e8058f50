#!/usr/bin/env python
"""bench.py — the hot path's headline benchmark (BASELINE.json).

Metric: 6x512x512 DenseNet-121 bf16 TRAIN images/s on N B200s (config[1]/[2]); one "step" = one pass of the
hot path over one batch per GPU: fused D4-augment + per-experiment normalise loader (u8 -> bf16 S2D) ->
DenseNet-121 forward -> CrossEntropy -> backward -> (NCCL gradient all-reduce over NVLink for N>1) -> nesterov SGD.

  python bench.py --gpus N --steps K --warmup W          our arm (launched under torchrun for N>1)
  python bench.py --impl reference ...                   the reference path on the host CPUs (oracle port)

One JSON line on stdout from rank 0 (see the contract in the task description): value (inputs resident in HBM),
e2e (host u8 batch in pinned memory -> H2D -> step -> D2H of the loss), roofline (conv implicit-GEMM family),
hbm_kernels (stats / loader GB/s against the measured HBM peak), cpu_baseline, clocks, gpu_launches.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "densenet121_6x512x512_bf16_train_images_per_sec"
UNIT = "images/s"
IMG = 512
NUM_CLASSES = 1108
FLOP_FWD_BWD_PER_IMG = 90.05e9     # SURVEY §A.3 (torchvision op order), 2*MAC
STATS_BYTES_PER_IMG = 6 * IMG * IMG            # 1,572,864
LOADER_BYTES_PER_IMG = 3 * 6 * IMG * IMG       # 1 B read + 2 B bf16 written per element = 4,718,592


OUT = sys.stdout


def claim_stdout():
    """Keep stdout for the JSON line alone: fd 1 is re-pointed at stderr so that C-level prints (NCCL's version
    banner at N>1, library warnings) cannot land in front of it; the line goes to the saved descriptor."""
    global OUT
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    OUT = os.fdopen(saved, "w")


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks and throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag.is_set():
                    break
                self.samples.append([f.strip() for f in line.split(",")])
        except Exception:
            pass

    def finish(self):
        self.stop_flag.set()
        if self.proc is not None:
            self.proc.terminate()
        sm, smax, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                smax = max(smax, float(s[1]))
                for i, nme in enumerate(names):
                    if s[3 + i].lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- CPU arms
def cpu_threads():
    """Use every host core: under torchrun OMP_NUM_THREADS defaults to 1, which made round 1's N>1 reference lines
    a one-core number.  Returns the thread count torch will really use."""
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        pass
    torch.set_num_threads(n)
    return torch.get_num_threads()


def cpu_port_step_rate(B, max_steps, warmup, seconds_budget=None):
    """THE CPU baseline of both arms (one definition): the reference path on the host — oracle normalise + D4 augment
    (numpy/OpenCV restatement of dataloader.py:128-139) + torchvision DenseNet-121 with the reference's 6-channel stem,
    fp32 forward / CrossEntropy / backward + nesterov SGD (main.py:89-93) — on a bounded sample of the workload:
    batch B of 6x512x512, `warmup` untimed steps, then up to `max_steps` timed steps (or until the budget runs out)."""
    from oracle import oracle_np as O
    from recursion_cellular_image_classification_b200.synth import synth_planes
    cores = cpu_threads()
    torch.manual_seed(0)
    net = O.densenet121_6ch(NUM_CLASSES, seed=0)
    net.train()
    opt = O.sgd_reference(net.parameters(), lr=0.0005 * B)
    lossf = torch.nn.CrossEntropyLoss()
    planes = synth_planes(3, n=B)
    mean = np.full(6, 0.1)
    std = np.full(6, 0.08)
    rng = np.random.default_rng(0)
    y = torch.from_numpy(rng.integers(0, NUM_CLASSES, size=B))

    def step():
        xs = [O.transform(planes[i], mean, std, vflip=bool(rng.integers(2)), hflip=bool(rng.integers(2)),
                          k=int(rng.integers(4)), crop_yx=(0, 0), out_hw=(IMG, IMG)) for i in range(B)]
        x = torch.from_numpy(np.stack(xs))
        opt.zero_grad()
        loss = lossf(net(x), y)
        loss.backward()
        opt.step()
        return loss.item()

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    n = 0
    while n < max_steps and (n < 1 or seconds_budget is None or time.perf_counter() - t0 < seconds_budget):
        step()
        n += 1
    dt = time.perf_counter() - t0
    sample = ("batch %d x 6x512x512, %d timed steps after %d warm-up: numpy/OpenCV normalise+D4 loader + torchvision "
              "densenet121 (6-ch stem) fp32 fwd+CE+bwd+nesterov SGD, %d threads" % (B, n, warmup, cores))
    return {"value": B * n / dt, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}, n, dt


def run_reference(args):
    """`--impl reference`: the CPU baseline above as its own JSON line (rank 0 only; other ranks exit without work)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = args.ref_batch
    steps, warmup = min(args.steps, args.ref_max_steps), min(args.warmup, 1)
    cpu, steps, dt = cpu_port_step_rate(B, steps, warmup)
    val = cpu["value"]
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "DenseNet-121 6-channel 512x512 training step: normalise+D4 loader, fwd, CE, bwd, "
                                   "nesterov SGD (the same step as the default arm; fp32 CPU port, bounded sample)",
                       "batch_per_gpu": B, "global_batch": B, "num_classes": NUM_CLASSES, "parallelism": "cpu",
                       "lr": 0.0005 * B, "host_threads": cpu["cores"], "os_cpu_count": os.cpu_count()},
            "cpu_baseline": cpu,
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=OUT, flush=True)


# ---------------------------------------------------------------------------------------------- our arm
def cpu_baseline(seconds_budget=25.0):
    """The same CPU port as `--impl reference` (batch 4), bounded to ~25 s."""
    return cpu_port_step_rate(4, 8, 1, seconds_budget)[0]


def library_gpu_baseline(dev, B=64, steps=6):
    """INFORMATIONAL, NOT THE REFERENCE ARM: the stock library path on the same box — torchvision densenet121 with the
    6-channel stem, bf16 autocast, channels_last, eager cuDNN (cudnn.benchmark like main.py:68), fwd + CE + bwd + SGD
    on a resident batch (BASELINE.md 1 names this as the bar on the box).  Kept small (batch 64) so it cannot run the
    box out of memory."""
    try:
        import torchvision
        torch.backends.cudnn.benchmark = True
        torch.manual_seed(0)
        net = torchvision.models.densenet121(weights=None, num_classes=NUM_CLASSES)
        net.features.conv0 = torch.nn.Conv2d(6, 64, kernel_size=7, stride=2, padding=3, bias=False)
        net = net.to(dev).to(memory_format=torch.channels_last).train()
        opt = torch.optim.SGD(net.parameters(), lr=0.0005 * B, momentum=0.9, nesterov=True, weight_decay=3e-5)
        x = torch.randn(B, 6, IMG, IMG, device=dev).to(memory_format=torch.channels_last)
        y = torch.randint(0, NUM_CLASSES, (B,), device=dev)
        lossf = torch.nn.CrossEntropyLoss()

        def step():
            opt.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                loss = lossf(net(x).float(), y)
            loss.backward()
            opt.step()

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            step()
        b_.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b_) / steps
        peak_gb = torch.cuda.max_memory_allocated(dev) / 1e9
        del net, opt, x, y
        torch.cuda.empty_cache()
        return {"value": B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "batch": B, "steps": steps,
                "peak_mem_gb": peak_gb,
                "what": "NOT the reference arm: torchvision densenet121 (6-ch stem), bf16 autocast, channels_last, "
                        "eager cuDNN/cuBLAS, model step only (no loader), same box"}
    except Exception as e:
        torch.cuda.empty_cache()
        return {"error": repr(e)}


def step_traffic_model(B):
    """ALGORITHMIC HBM bytes of one training step under the CURRENT data flow (bf16 NHWC activations,
    one concat buffer per block, lazy BatchNorm backward), per kernel family — every operand read once, every result
    written once, the running concat gradient read AND written.  Keys match kernel_breakdown_fine."""
    px = lambda h: B * h * h
    blocks, C, H = (6, 12, 24, 16), 64, IMG // 4
    t = {k: 0.0 for k in ("conv_fwd_1x1", "conv_fwd_3x3", "conv_wgrad_3x3", "conv_dgrad_3x3", "bn_bwd_apply",
                          "conv_wgrad_1x1", "conv_dgrad_1x1", "grad_fixup", "conv_other", "wgrad_other",
                          "elementwise_other", "loader")}
    Ms = px(IMG // 2)
    # stem: conv (S2D input 32 ch -> 64), BN+ReLU+maxpool (+indices), pool backward, BN backward apply, weight gradient
    t["conv_other"] += 2.0 * Ms * (32 + 64)
    t["elementwise_other"] += 2.0 * Ms * 64 + px(H) * (2.0 * 64 + 64)
    t["elementwise_other"] += px(H) * (2.0 * 64 + 64) + 2.0 * Ms * 64 * 2
    t["bn_bwd_apply"] += 2.0 * Ms * 64 * 3
    t["wgrad_other"] += 2.0 * Ms * (64 + 32)
    for b, n_layers in enumerate(blocks):
        M, C0 = px(H), C
        for i in range(n_layers):
            cin = C + 32 * i
            t["conv_fwd_1x1"] += 2.0 * M * (cin + 128)
            t["conv_fwd_3x3"] += 2.0 * M * (128 + 32)
            # (the 3x3 weight gradient rides in the 3x3 data-gradient kernel wherever its 8x16 tiling applies: H > 8; that
            # kernel then also derives its dOut from the G and X slices of the concat buffers on load - no grad_fixup
            # pass writing and re-reading a dense dZ)
            env_on = lambda k: os.environ.get(k, "0") not in ("", "0")
            fused3 = H > 8 and not env_on("RXB_DBG_NO_WGFUSE3")
            folded = fused3 and not env_on("RXB_DBG_NO_FIXFOLD") and not env_on("RXB_DBG_NO_BNTAIL")
            if not fused3:
                t["conv_wgrad_3x3"] += 2.0 * M * (128 + 32)
            if not folded:
                t["grad_fixup"] += 2.0 * M * 32 * 3
            t["conv_dgrad_3x3"] += 2.0 * M * ((64 if folded else 32) + 128 + 128)
            t["bn_bwd_apply"] += 2.0 * M * 128 * 3
            # the 1x1 weight gradient is accumulated by the 1x1 data-gradient kernel from the tiles it already holds
            # (conv_gemm.cu EPI 3) and moves no bytes of its own; RXB_DBG_NO_WGFUSE=1 restores the separate launch
            if os.environ.get("RXB_DBG_NO_WGFUSE", "0") not in ("", "0"):
                t["conv_wgrad_1x1"] += 2.0 * M * (cin + 128)
            t["conv_dgrad_1x1"] += 2.0 * M * (128 + 3 * cin)
        C += 32 * n_layers
        t["grad_fixup"] += 2.0 * M * C0 * 3                         # exact gradient of the block input
        if b < 3:
            Mq = M / 4
            t["elementwise_other"] += 2.0 * (M * C + Mq * C)         # BN+ReLU+avgpool
            t["conv_other"] += 2.0 * Mq * (C + C // 2)               # transition conv (after the pool)
            t["wgrad_other"] += 2.0 * Mq * (C + C // 2)
            t["conv_other"] += 2.0 * Mq * (C // 2 + C)               # its data gradient
            t["elementwise_other"] += 2.0 * (Mq * C + 2 * M * C)     # pool/ReLU/BN backward into G
            C //= 2
            H //= 2
        else:
            t["elementwise_other"] += 2.0 * M * C + 2.0 * 2 * M * C  # norm5+ReLU+GAP forward; backward into G
    t["loader"] = float(LOADER_BYTES_PER_IMG) * B
    return t


def ncu_traffic(B):
    """DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of ONE launch of each representative kernel, read from
    the committed summary of the `ncu --set full` capture (profiles/ncu_traffic.json, written by tools/ncu_traffic.py
    from the raw CSV of the same commands) — null when no capture exists for this batch size."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return {}
    d = json.load(open(path))
    return d.get("batch_%d" % B, {}).get("kernels", {})


def conv_kernel_rooflines(B, dev, peaks):
    """Six representative dense-block-1 launches (the geometry that dominates the step), each timed alone with
    CUDA events on inputs far larger than L2.  Algorithmic bytes = every operand read once + result written once
    (the running gradient of the accumulating dgrad is read AND written); FLOPs = 2*M*N*K."""
    from recursion_cellular_image_classification_b200 import ops

    def timeit(fn, reps=6):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b_.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b_) / reps

    H = W = IMG // 4
    M = B * H * W
    out = {}
    traffic_tab = ncu_traffic(B)

    def entry(name, ms, bytes_, flops, what):
        gbs = bytes_ / (ms * 1e-3) / 1e9
        tf = flops / (ms * 1e-3) / 1e12
        tr = traffic_tab.get(name)
        out[name] = {"what": what, "bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": gbs / peaks["hbm_gbs"], "peak_source": peaks["source"] + " copy", "ms": ms,
                     "algorithmic_bytes": bytes_, "tflops": tf, "tensor_frac": tf / peaks["bf16_tflops_sustained"],
                     "traffic": tr}

    def smem_port(name, wavefronts_per_tile, tiles, breakdown):
        """The 3x3 kernels are bound by the shared-memory port (128 B per clock and SM), not by HBM or the tensor
        pipe: modelled 128-byte wavefronts per tile (tensor-core operand reads at
        (A + B bytes)/128 per MMA, the in-place transforms, the epilogue's reads and writes, TMA writes and store reads;
        the measured counterpart is l1tex__data_pipe_*_wavefronts_mem_shared in profiles/r02_ncu_summary_fused.txt)
        against the cycles a tile takes at the nominal SM clock."""
        try:
            ms = out[name]["ms"]
            cycles_per_tile = ms * 1e-3 * 1.965e9 / (tiles / 148.0)
            out[name]["smem_port"] = {"bound": "shared-memory port (128 B/clk/SM)",
                                      "model_wavefronts_per_tile": wavefronts_per_tile, "cycles_per_tile": cycles_per_tile,
                                      "frac": wavefronts_per_tile / cycles_per_tile, "model": breakdown,
                                      "sm_clock_mhz_assumed": 1965,
                                      "note": "cycles at the NOMINAL clock: under sw_power_cap the SMs run slower (ncu: "
                                              "525k cycles for a 307 us launch = 1.71 GHz), so the real fraction of the "
                                              "port is ~15 % higher; ncu counts 2520 / 5035 wavefronts per tile"}
        except Exception as e:          # informational: never takes the bench line down
            out[name]["smem_port"] = {"error": repr(e)}

    g = torch.Generator(device=dev)
    g.manual_seed(5)
    rnd = lambda *sh: torch.randn(*sh, device=dev, generator=g).to(torch.bfloat16)
    X = rnd(B, H, W, 256)                       # concat buffer of dense block 1 (Ctot = 256)
    Y = rnd(B, H, W, 128)                       # bottleneck activation
    sc224, sh224 = torch.rand(224, device=dev) + 0.5, torch.randn(224, device=dev) * 0.1
    sc128, sh128 = torch.rand(128, device=dev) + 0.5, torch.randn(128, device=dev) * 0.1
    W1 = (torch.randn(1, 1, 128, 224, device=dev) * 0.05).to(torch.bfloat16)
    W3 = (torch.randn(3, 3, 32, 128, device=dev) * 0.03).to(torch.bfloat16)
    # forward 1x1 (Cin 224 -> 128, BN+ReLU prologue, statistics) and 3x3 (128 -> 32 written into the concat buffer)
    ms = timeit(lambda: ops.conv_fwd(X, W1, Cin=224, scale=sc224, shift=sh224, out=Y, stats=True))
    entry("fwd_1x1", ms, M * (224 + 128) * 2, 2.0 * M * 128 * 224, "dense layer 1x1, Cin 224")
    ms = timeit(lambda: ops.conv_fwd(Y, W3, Cin=128, scale=sc128, shift=sh128, out=X, c_off=224, pad=(1, 1), stats=True))
    entry("fwd_3x3", ms, M * (128 + 32) * 2, 2.0 * M * 32 * 128 * 9, "dense layer 3x3, 128 -> 32")
    # x-merged 16x8 tiles with 14 valid columns: ceil(W/14) x H/8 tiles per image
    smem_port("fwd_3x3", 1344 + 736 + 424 + 100, B * ((W + 13) // 14) * (H // 8),
              {"mma_operands": "24 MMAs x (4 KB A + 3 KB B)/128 = 1344", "transform": "2 k-blocks x 160 rows x 128 B, read + write = 640 (+96 constants)",
               "tma": "46 KB in + 7 KB out = 424", "epilogue": 100})
    # backward: 3x3 dgrad (dZ 32 -> dy2 128), 1x1 dgrad accumulating into the concat gradient, both wgrads
    dZ = rnd(B, H, W, 32)
    W3d = (torch.randn(3, 3, 128, 32, device=dev) * 0.03).to(torch.bfloat16)
    dy2 = torch.empty(B, H, W, 128, dtype=torch.bfloat16, device=dev)
    ms = timeit(lambda: ops.conv_dgrad_bn(dZ, W3d, Y, sc128, sh128, 128, out_mode=ops.OUT_DY, out=dy2, pad=(1, 1)))
    entry("dgrad_3x3_bn", ms, M * (32 + 128 + 128) * 2, 2.0 * M * 128 * 32 * 9, "3x3 data gradient + ReLU/BN2 backward")
    ms = timeit(lambda: ops.conv_dgrad_bn(dZ, W3d, Y, sc128, sh128, 128, out_mode=ops.OUT_DY, out=dy2, pad=(1, 1), wgrad=True))
    entry("dgrad_3x3_bn_wgrad", ms, M * (32 + 128 + 128) * 2, 2.0 * M * 128 * 32 * 9 * 2,
          "the same launch also accumulating the 3x3 weight gradient (what the step runs; bound by shared-memory "
          "bandwidth of the tensor-core operand reads, DESIGN.md 5.3)")
    smem_port("dgrad_3x3_bn_wgrad", 1152 + 1344 + 288 + 512 + 512 + 608, B * (W // 8) * (H // 16),
              {"dgrad_mma_operands": "18 MMAs x 8 KB/128 = 1152", "wgrad_mma_operands": "24 MMAs x 7 KB/128 = 1344",
               "statistics_mma": "8 x 4.5 KB/128 = 288", "transform": "32 KB read + 32 KB write = 512",
               "epilogue": "32 KB read + 32 KB write = 512", "tma": "44 KB in + 32 KB store read = 608"})
    W1d = (torch.randn(1, 1, 224, 128, device=dev) * 0.05).to(torch.bfloat16)
    G = torch.zeros(B, H, W, 256, dtype=torch.bfloat16, device=dev)
    ms = timeit(lambda: ops.conv_dgrad_bn(dy2, W1d, X, sc224, sh224, 224, out_mode=ops.OUT_G_ACCUM, out=G))
    entry("dgrad_1x1_bn_accum", ms, M * (128 + 3 * 224) * 2, 2.0 * M * 224 * 128,
          "1x1 data gradient + ReLU/BN1 backward, L2 reduce-add into the concat gradient (Cin 224)")
    ms = timeit(lambda: ops.conv_dgrad_bn(dy2, W1d, X, sc224, sh224, 224, out_mode=ops.OUT_G_ACCUM, out=G, wgrad=True))
    entry("dgrad_1x1_bn_accum_wgrad", ms, M * (128 + 3 * 224) * 2, 2.0 * M * 224 * 128 * 2,
          "the same launch also accumulating the 1x1 weight gradient from the tiles it holds (what the step runs)")
    ms = timeit(lambda: ops.conv_wgrad(Y, dZ, 128, 32, taps=(3, 3), pad=(1, 1), scale=sc128, shift=sh128))
    entry("wgrad_3x3", ms, M * (128 + 32) * 2, 2.0 * M * 32 * 128 * 9, "3x3 weight gradient")
    ms = timeit(lambda: ops.conv_wgrad(X, dy2, 224, 128, scale=sc224, shift=sh224))
    entry("wgrad_1x1", ms, M * (224 + 128) * 2, 2.0 * M * 128 * 224, "1x1 weight gradient, Cin 224")
    return out


def widen_kernel_numbers(dev, peaks, time_kernel, big, big_idx, big_exp, big_crop, norm_m, norm_d):
    """The widening rows, timed alone like hbm_kernels: the arbitrary-angle loader against the HBM
    roofline on the loader's algorithmic bytes, and the device JPEG decoder (latency-bound entropy decoding: files/s,
    with cv2.imdecode on one host core beside it) on 768 q95 512x512 files of two synthetic corpora."""
    import cv2
    from recursion_cellular_image_classification_b200 import ops
    from recursion_cellular_image_classification_b200.synth import synth_planes
    n = big.shape[0]
    rng = np.random.default_rng(0)
    mats = torch.from_numpy(np.stack([ops.rotation_matrix(IMG, IMG, float(a)) for a in rng.uniform(-180, 180, n)])).to(dev)
    flips = torch.from_numpy(rng.integers(0, 4, n).astype(np.uint8)).to(dev)
    dst = torch.empty(n, IMG // 2, IMG // 2, 32, dtype=torch.bfloat16, device=dev)
    t = time_kernel(lambda: ops.load_norm_affine(big, big_idx, big_exp, flips, mats, big_crop, norm_m, norm_d,
                                                 (IMG, IMG), ops.OUT_BF16_S2D32, out=dst))
    ach = LOADER_BYTES_PER_IMG * n / (t * 1e-3) / 1e9
    out = {"loader_affine_kernel": {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                    "frac": ach / peaks["hbm_gbs"], "images": n, "ms": t,
                                    "what": "flips + cv2.warpAffine-exact rotation + normalise, u8 planar -> bf16 S2D32"}}
    del dst
    planes = synth_planes(6, n=2)
    corpora = {"dense_noise": [planes[i, c] for i in range(2) for c in range(6)],
               "smooth": [cv2.GaussianBlur(planes[i, c], (0, 0), 2.0) for i in range(2) for c in range(6)]}
    planes_out = torch.empty(768, IMG, IMG, dtype=torch.uint8, device=dev)
    for name, imgs in corpora.items():
        bufs = [cv2.imencode(".jpg", im, [cv2.IMWRITE_JPEG_QUALITY, 95])[1].tobytes() for im in imgs]
        t0 = time.perf_counter()
        ref = [cv2.imdecode(np.frombuffer(b_, np.uint8), -1) for b_ in bufs]
        cpu_ms = (time.perf_counter() - t0) / len(bufs) * 1e3
        blob, offsets = ops.pack_jpeg_buffers(bufs * 64)                       # 768 files = 128 six-channel images
        blob, offsets = blob.to(dev), offsets.to(dev)
        entry = {"files": 768, "compressed_mb": blob.numel() / 1e6,
                 "cpu_cv2_imdecode_files_per_s_1_core": 1e3 / cpu_ms}
        for key, par in (("all_lanes", True), ("single_lane", False)):
            t = time_kernel(lambda: ops.jpeg_decode_gray(blob, offsets, (IMG, IMG), out=planes_out, check_status=False,
                                                         parallel=par), reps=5)
            entry[key] = {"ms": t, "files_per_s": 768 / (t * 1e-3), "images_per_s": 128 / (t * 1e-3)}
        entry["bit_exact_vs_cv2"] = bool((planes_out.view(64, 12, IMG, IMG) ==
                                          torch.from_numpy(np.stack(ref)).to(dev)[None]).all().item())
        out["jpeg_decode_" + name] = entry
    return out


def run_ours(args):
    import torch.distributed as dist
    from recursion_cellular_image_classification_b200 import _lib, ops
    from recursion_cellular_image_classification_b200.cell_classifier.models import DenseNet121
    from recursion_cellular_image_classification_b200.synth import synth_planes_torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import datetime
        # a short collective timeout: a rank mismatch aborts in two minutes instead of hanging the box
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
    _lib.require_gpu()
    lib = _lib.load()
    peaks = measured_peaks()

    B = args.batch
    gB = B * world
    fake_world = int(os.environ.get("RXB_BENCH_FAKE_WORLD", "1"))   # development: N-rank hyper-parameters on one GPU
    lr = 0.0005 * gB * fake_world                      # main.py:71
    n_exp = 4
    torch.manual_seed(1234 + rank)
    src = synth_planes_torch(rank, B, dev)             # u8 [B,6,512,512] resident in HBM
    host_src = torch.empty(src.shape, dtype=torch.uint8, pin_memory=True)
    host_src.copy_(src)
    src_idx = torch.arange(B, dtype=torch.int32, device=dev)
    exp_id = (torch.arange(B, device=dev) % n_exp).to(torch.int32)
    crop = torch.zeros(B, 2, dtype=torch.int32, device=dev)
    labels = torch.randint(0, NUM_CLASSES, (B,), device=dev)
    # per-experiment statistics through our own stats kernel (family 1a)
    acc = ops.stats_accumulate(src, exp_id, n_exp)
    mean, std = ops.stats_finalize(acc)
    m, d = ops.normalize_constants(mean.cpu().numpy(), std.cpu().numpy())
    norm_m, norm_d = torch.from_numpy(m).to(dev), torch.from_numpy(d).to(dev)

    net = DenseNet121(nb_classes=NUM_CLASSES, device=dev, seed=0)
    net.train()
    # the executor's fixed-address step buffers: the loader writes xs, the labels sit beside it, and every backward
    # phase is replayed from a CUDA graph (the product default, cell_classifier/train.py); --no-graph = stream launches
    xs, labels_static, loss_dev = net.static_buffers(B, IMG, IMG)
    labels_static.copy_(labels)
    labels = labels_static
    use_graph = {"on": not args.no_graph}
    host_loss = torch.empty(1, dtype=torch.float32, pin_memory=True)
    gen = torch.Generator(device=dev)
    gen.manual_seed(rank)
    n_phases = lib.rxb_dn121_num_phases()
    ranges = [net.phase_grad_range(B, IMG, IMG, ph) for ph in range(n_phases)]

    # end-to-end arm: every step's u8 batch comes from pinned host memory.  Like a prefetching data loader, the copy of
    # step i+1 runs on a copy stream into the second device buffer while step i computes; each timed step still
    # performs exactly one H2D copy inside the timed region (the first one is not hidden).
    copy_stream = torch.cuda.Stream(device=dev)
    src2 = [src, torch.empty_like(src)]
    copy_done = [torch.cuda.Event(), torch.cuda.Event()]
    loader_done = [torch.cuda.Event(), torch.cuda.Event()]
    pf = {"i": 0, "primed": False}

    def issue_copy(buf):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(loader_done[buf])                     # the loader that last read this buffer is done
            src2[buf].copy_(host_src, non_blocking=True)                 # H2D of one step's u8 batch
            copy_done[buf].record(copy_stream)

    def step(from_host, last=True):
        cur = torch.cuda.current_stream()
        s_in = src
        if from_host:
            buf = pf["i"] % 2
            if not pf["primed"]:
                issue_copy(buf)
                pf["primed"] = True
            cur.wait_event(copy_done[buf])
            s_in = src2[buf]
        aug = torch.randint(0, 16, (B,), device=dev, generator=gen, dtype=torch.uint8)   # D4 code per image
        ops.load_norm_aug(s_in, src_idx, exp_id, aug, crop, norm_m, norm_d, (IMG, IMG), ops.OUT_BF16_S2D32, out=xs)
        if from_host:
            loader_done[buf].record(cur)
            if last:
                pf["primed"] = False                                     # the next host step starts with its own copy
            else:
                issue_copy(1 - buf)                                      # prefetch the next step's batch
            pf["i"] += 1
        works = []
        for ph in range(n_phases):
            net.train_step(xs, labels, global_batch=gB * fake_world, phase=ph, loss_out=loss_dev, graph=use_graph["on"])
            if world > 1:
                b, e = ranges[ph]
                if os.environ.get("RXB_BENCH_SYNC_AR") == "1":                       # development: no overlap
                    dist.all_reduce(net.flat.grad[b:e])
                else:
                    works.append(dist.all_reduce(net.flat.grad[b:e], async_op=True))   # overlaps the next phase
        for w in works:
            w.wait()
        net.sgd_step(B, IMG, IMG, lr=lr, momentum=0.9, weight_decay=3e-5, nesterov=True)
        if from_host:
            host_loss.copy_(loss_dev, non_blocking=True)                 # D2H of the step's result

    def timed(from_host, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(steps):
            step(from_host, last=(k == steps - 1))
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    n_warm = args.warmup if args.quick else max(args.warmup, 3)
    for _ in range(n_warm):
        step(False)
    torch.cuda.synchronize()
    loss_first = loss_dev.item()
    # host time to ENQUEUE one step into an empty stream (no synchronisation inside): shows whether the step is
    # launch-bound on the CPU side
    t_h = time.perf_counter()
    step(False)
    host_enqueue_ms = (time.perf_counter() - t_h) * 1e3
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    lib.rxb_launch_count_reset()
    g0 = net.graph_launches
    ms = timed(False, args.steps)
    launches = lib.rxb_launch_count() + (net.graph_launches - g0)      # enqueued + replayed from the phase graphs
    clocks = sampler.finish() if rank == 0 else None
    value = gB * args.steps / (ms * 1e-3)
    if args.quick:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                              "warmup": n_warm, "ms_per_step": ms / args.steps, "quick": True,
                              "host_enqueue_ms": host_enqueue_ms,
                              "gpu_launches": int(launches)}), file=OUT, flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    # the same steps with plain stream launches (no graph replay), for the delta
    graph_result = None
    if use_graph["on"]:
        use_graph["on"] = False
        for _ in range(2):
            step(False)
        ms_stream = timed(False, args.steps)
        use_graph["on"] = True
        graph_result = {"value": gB * args.steps / (ms_stream * 1e-3), "ms_per_step": ms_stream / args.steps,
                        "note": "the same steps enqueued kernel by kernel (--no-graph); the headline replays the five "
                                "backward phases from CUDA graphs"}

    for _ in range(2):
        step(True)
    ms_e2e = timed(True, args.steps)
    e2e_value = gB * args.steps / (ms_e2e * 1e-3)
    loss_last = host_loss.item()

    # ---- kernel-family breakdown (event-bracketed launches, separate untimed pass) and HBM kernels
    breakdown, roofline, hbm_kernels, roofline_tensor, conv_kernels, widen_kernels = None, None, None, None, None, None
    fine, step_roofline = None, None
    # every rank runs the profiled steps (they contain the gradient all-reduces); only rank 0 reports
    ncat = 17
    msb = (ctypes.c_float * ncat)()
    cnt = (ctypes.c_int64 * ncat)()
    prof_steps = 2
    torch.cuda.synchronize()
    graph_was_on, use_graph["on"] = use_graph["on"], False      # event brackets need real launches, not graph replays
    lib.rxb_profile_enable(1)
    for _ in range(prof_steps):
        step(False)
    _lib.check(lib.rxb_profile_collect(msb, cnt, ncat))
    lib.rxb_profile_enable(0)
    use_graph["on"] = graph_was_on
    if world > 1:
        dist.barrier()
    if rank == 0:
        fine_names = ["stats", "loader", "conv_fwd_1x1", "conv_dgrad_1x1", "conv_wgrad_1x1", "elementwise_other", "head",
                      "optimizer", "tta", "conv_fwd_3x3", "conv_dgrad_3x3", "conv_wgrad_3x3", "conv_other", "wgrad_other",
                      "bn_bwd_apply", "grad_fixup", "bn_bwd_finalize"]
        model = step_traffic_model(B)
        fine = {}
        for i, nme in enumerate(fine_names):
            e = {"ms_per_step": msb[i] / prof_steps, "launches_per_step": cnt[i] / prof_steps}
            if nme in model and e["ms_per_step"] > 0:
                e["algorithmic_gb_per_step"] = model[nme] / 1e9
                e["hbm_frac"] = model[nme] / (e["ms_per_step"] * 1e-3) / 1e9 / peaks["hbm_gbs"]
            fine[nme] = e
        fam = {"stats": ["stats"], "loader": ["loader"], "conv_fwd": ["conv_fwd_1x1", "conv_fwd_3x3", "conv_other"],
               "conv_dgrad": ["conv_dgrad_1x1", "conv_dgrad_3x3"],
               "conv_wgrad": ["conv_wgrad_1x1", "conv_wgrad_3x3", "wgrad_other"],
               "elementwise": ["elementwise_other", "bn_bwd_apply", "grad_fixup", "bn_bwd_finalize"], "head": ["head"],
               "optimizer": ["optimizer"], "tta": ["tta"]}
        breakdown = {k: {"ms_per_step": sum(fine[n]["ms_per_step"] for n in v),
                         "launches_per_step": sum(fine[n]["launches_per_step"] for n in v)} for k, v in fam.items()}
        prof_total_ms = max(sum(v["ms_per_step"] for v in fine.values()), 1e-9)
        conv_ms = sum(breakdown[k]["ms_per_step"] for k in ("conv_fwd", "conv_dgrad", "conv_wgrad"))
        conv_launches = sum(breakdown[k]["launches_per_step"] for k in ("conv_fwd", "conv_dgrad", "conv_wgrad"))
        achieved_tf = FLOP_FWD_BWD_PER_IMG * B / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
        roofline_tensor = {"kernel": "conv_gemm_kernel + conv_wgrad_kernel (tcgen05 implicit GEMM family), whole step",
                           "bound": "tensor", "achieved": achieved_tf, "peak": peaks["bf16_tflops_sustained"],
                           "unit": "TFLOP/s", "frac": achieved_tf / peaks["bf16_tflops_sustained"],
                           "peak_source": peaks["source"] + " sustained cuBLAS bf16",
                           "algorithmic_flops_per_step": FLOP_FWD_BWD_PER_IMG * B, "launches_per_step": conv_launches,
                           "avg_launch_ms": conv_ms / max(conv_launches, 1),
                           "note": "the stride-1 DenseNet convs at bf16 have 43-230 FLOP/B arithmetic intensity, at or "
                                   "below the B200 ridge (~217 FLOP/B): they are HBM-bound, see DESIGN.md section 5"}
        conv_kernels = conv_kernel_rooflines(B, dev, peaks)
        # Whole-step position: against the tensor roofline SURVEY 8d names for the convolutions, and against the HBM
        # traffic the current data flow moves (the bound the step can reach without moving fewer bytes).
        step_ms = ms / args.steps
        traffic_bytes = sum(model.values())
        step_roofline = {"tensor_frac_step": FLOP_FWD_BWD_PER_IMG * B / (step_ms * 1e-3) / 1e12 / peaks["bf16_tflops_sustained"],
                         "tflops_step": FLOP_FWD_BWD_PER_IMG * B / (step_ms * 1e-3) / 1e12,
                         "algorithmic_hbm_gb_per_step": traffic_bytes / 1e9,
                         "hbm_traffic_bound_ms": traffic_bytes / (peaks["hbm_gbs"] * 1e9) * 1e3,
                         "hbm_frac_step": traffic_bytes / (step_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                         "ms_per_step": step_ms}
        # `roofline`: the kernel SYMBOL with the largest share of the step (launch list in profiles/), its bytes and
        # time summed over ALL its launches of the step (event-bracketed), with the tensor number beside the HBM one.
        symbols = {
            "conv_wgrad_kernel": (["conv_wgrad_1x1", "conv_wgrad_3x3", "wgrad_other"], "wgrad_1x1",
                                  "conv_wgrad_kernel (weight gradients NOT fused into a data-gradient kernel: transitions, stem)"),
            "conv_gemm_kernel<64,false,3>": (["conv_dgrad_1x1"], "dgrad_1x1_bn_accum_wgrad",
                                             "conv_gemm_kernel<64,false,3> (dense-layer 1x1 data gradient + ReLU/BN backward, "
                                             "L2 reduce-add into the concat gradient, + the 1x1 weight gradient and the "
                                             "BatchNorm-backward reductions from the same tiles)"),
            "conv_gemm_kernel<64,true,0>": (["conv_fwd_1x1"], "fwd_1x1", "conv_gemm_kernel<64,true,0> (dense-layer 1x1 forward)"),
            "conv_gemm_kernel<64,true,1>": (["conv_fwd_3x3"], "fwd_3x3", "conv_gemm_kernel<64,true,1> (dense-layer 3x3 forward)"),
            "conv_gemm_kernel<32,false,4>": (["conv_dgrad_3x3"], "dgrad_3x3_bn_wgrad",
                                             "conv_gemm_kernel<32,false,4> (3x3 data gradient + ReLU/BN backward + 3x3 weight gradient)"),
        }
        share = {k: sum(fine[n]["ms_per_step"] for n in v[0]) / prof_total_ms for k, v in symbols.items()}
        top = max(share, key=share.get)
        cats, rep, label = symbols[top]
        t_ms = sum(fine[n]["ms_per_step"] for n in cats)
        n_l = sum(fine[n]["launches_per_step"] for n in cats)
        alg = sum(model[n] for n in cats)
        # algorithmic FLOPs per image of each symbol's launches (dense 1x1 12.17, dense 3x3 12.98 GFLOP
        # forward; a weight gradient exists for every conv: 30.84)
        flops_share = {"conv_wgrad_kernel": 30.84e9 - 12.17e9 - 12.98e9, "conv_gemm_kernel<64,false,3>": 2 * 12.17e9,
                       "conv_gemm_kernel<64,true,0>": 12.17e9, "conv_gemm_kernel<64,true,1>": 12.98e9,
                       "conv_gemm_kernel<32,false,4>": 2 * 12.98e9}[top] * B
        ach = alg / (t_ms * 1e-3) / 1e9
        roofline = {"kernel": label, "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": ach / peaks["hbm_gbs"], "peak_source": peaks["source"] + " copy",
                    "launches_per_step": n_l, "avg_launch_ms": t_ms / max(n_l, 1), "ms_per_step": t_ms,
                    "algorithmic_bytes_per_step": alg, "algorithmic_bytes_per_launch": alg / max(n_l, 1),
                    "share_of_step": share[top], "share_of_step_by_symbol": share,
                    "traffic": conv_kernels[rep]["traffic"], "traffic_launch": rep,
                    "representative_launch": conv_kernels[rep],
                    "tensor": {"achieved": flops_share / (t_ms * 1e-3) / 1e12, "peak": peaks["bf16_tflops_sustained"],
                               "unit": "TFLOP/s", "frac": flops_share / (t_ms * 1e-3) / 1e12 / peaks["bf16_tflops_sustained"],
                               "note": "SURVEY 8d names the tensor roofline for convolutions; at bf16 these kernels sit "
                                       "below the ridge, so the HBM figure is the binding one and both are given"},
                    "step": step_roofline}
        # HBM-bound families, timed alone on >L2 inputs
        def time_kernel(fn, reps=10):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b_.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b_) / reps
        nbig = max(B, 128)
        big = synth_planes_torch(7, nbig, dev) if nbig != B else src
        big_exp = (torch.arange(nbig, device=dev) % n_exp).to(torch.int32)
        acc2 = tuple(torch.zeros(n_exp, 6, dtype=torch.int64, device=dev) for _ in range(3))
        t_stats = time_kernel(lambda: ops.stats_accumulate(big, big_exp, n_exp, acc2))
        big_idx = torch.arange(nbig, dtype=torch.int32, device=dev)
        big_crop = torch.zeros(nbig, 2, dtype=torch.int32, device=dev)
        big_aug = torch.randint(0, 16, (nbig,), device=dev, dtype=torch.uint8)
        big_out = torch.empty(nbig, IMG // 2, IMG // 2, 32, dtype=torch.bfloat16, device=dev)
        t_load = time_kernel(lambda: ops.load_norm_aug(big, big_idx, big_exp, big_aug, big_crop, norm_m, norm_d,
                                                       (IMG, IMG), ops.OUT_BF16_S2D32, out=big_out))
        def hb(bytes_per_img, t_ms):
            ach = bytes_per_img * nbig / (t_ms * 1e-3) / 1e9
            return {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": ach / peaks["hbm_gbs"], "images": nbig, "ms": t_ms, "peak_source": peaks["source"] + " copy"}
        hbm_kernels = {"stats_planar_kernel": hb(STATS_BYTES_PER_IMG, t_stats),
                       "loader_kernel": hb(LOADER_BYTES_PER_IMG, t_load)}
        del big_out
        widen_kernels = None
        if world == 1:
            try:
                widen_kernels = widen_kernel_numbers(dev, peaks, time_kernel, big, big_idx, big_exp, big_crop, norm_m,
                                                     norm_d)
            except Exception as e:       # the widening rows never take the headline line down
                widen_kernels = {"error": repr(e)}
            # the same step fed with JPEG FILES from pinned host memory (ImagesDS(decode='gpu')): H2D of the compressed
            # bytes, device decode, loader, forward/backward, SGD, loss read back — on the two synthetic corpora
            try:
                import cv2
                from recursion_cellular_image_classification_b200.synth import synth_planes
                base = synth_planes(6, n=2)
                fixed_codes = torch.randint(0, 16, (B,), device=dev, dtype=torch.uint8)
                planes_dev = torch.empty(B * 6, IMG, IMG, dtype=torch.uint8, device=dev)
                e2e_jpeg = {}
                for cname, imgs in (("dense_noise", [base[i, c] for i in range(2) for c in range(6)]),
                                    ("smooth", [cv2.GaussianBlur(base[i, c], (0, 0), 2.0) for i in range(2) for c in range(6)])):
                    files = [cv2.imencode(".jpg", im, [cv2.IMWRITE_JPEG_QUALITY, 95])[1].tobytes() for im in imgs]
                    blob, offs = ops.pack_jpeg_buffers((files * ((B * 6 + 11) // 12))[:B * 6])
                    h_blob, h_offs = blob.pin_memory(), offs.pin_memory()
                    d_blob, d_offs = torch.empty_like(blob, device=dev), torch.empty_like(offs, device=dev)

                    def jpeg_step():
                        # (stream launches: the input buffer here is not the executor's static one)
                        d_blob.copy_(h_blob, non_blocking=True)
                        d_offs.copy_(h_offs, non_blocking=True)
                        ops.jpeg_decode_gray(d_blob, d_offs, (IMG, IMG), out=planes_dev, check_status=False)
                        ops.load_norm_aug(planes_dev.view(B, 6, IMG, IMG), src_idx, exp_id, fixed_codes, crop, norm_m,
                                          norm_d, (IMG, IMG), ops.OUT_BF16_S2D32, out=xs)
                        for ph in range(n_phases):
                            net.train_step(xs, labels, global_batch=gB * fake_world, phase=ph, loss_out=loss_dev)
                        net.sgd_step(B, IMG, IMG, lr=lr, momentum=0.9, weight_decay=3e-5, nesterov=True)
                        host_loss.copy_(loss_dev, non_blocking=True)

                    t = time_kernel(jpeg_step, reps=max(args.steps, 3))
                    e2e_jpeg[cname] = {"value": B / (t * 1e-3), "unit": UNIT, "ms_per_step": t,
                                       "h2d_bytes_per_step": int(blob.numel() + 8 * offs.numel()), "d2h_bytes_per_step": 4,
                                       "files_per_step": B * 6}
                    del d_blob, d_offs
                widen_kernels["e2e_from_jpeg_files"] = e2e_jpeg
                del planes_dev
            except Exception as e:
                if isinstance(widen_kernels, dict):
                    widen_kernels["e2e_from_jpeg_files"] = {"error": repr(e)}

    if rank == 0:
        cpu = cpu_baseline() if world == 1 and not args.no_cpu_baseline else None
        lib_gpu = library_gpu_baseline(dev) if world == 1 and not args.no_library_baseline else None
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": n_warm, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": "DenseNet-121 6-channel 512x512 bf16 training step: fused normalise+D4 loader, "
                                       "fwd, CE, bwd, %snesterov SGD" % ("NCCL grad all-reduce, " if world > 1 else ""),
                           "batch_per_gpu": B, "global_batch": gB, "num_classes": NUM_CLASSES,
                           "parallelism": "dp%d" % world, "l2": "inputs and activations exceed the 126 MB L2",
                           "lr": lr},
                "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                        "h2d_bytes_per_step": int(host_src.numel()) * world, "d2h_bytes_per_step": 4 * world},
                "gpu_launches": int(launches), "host_enqueue_ms_per_step": host_enqueue_ms,
                "clocks": clocks,
                "roofline": roofline, "roofline_tensor_conv_family": roofline_tensor, "conv_kernels": conv_kernels,
                "step_roofline": step_roofline,
                "hbm_kernels": hbm_kernels, "widen_kernels": widen_kernels, "kernel_breakdown": breakdown,
                "kernel_breakdown_fine": fine,
                "cpu_baseline": cpu, "library_gpu_baseline": lib_gpu,
                "loss": {"after_warmup": loss_first, "last": loss_last}}
        line["config"]["launch"] = "cuda graph per backward phase" if use_graph["on"] else "stream launches"
        if graph_result is not None:
            line["stream_launch_comparison"] = graph_result
        print(json.dumps(line), file=OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=128, help="images per GPU (SURVEY 8d sweep: 32/64/128)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-batch", type=int, default=4)
    ap.add_argument("--ref-max-steps", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-library-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true",
                    help="enqueue every kernel on the stream instead of replaying the phases from CUDA graphs")
    ap.add_argument("--quick", action="store_true",
                    help="profiling aid (ncu): exactly --warmup + --steps steps, no e2e / breakdown / cpu baseline")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
