#!/usr/bin/env bash
# round 2, call p: N=128 weight-gradient MMA chain + row-walking grad_fixup: parity, step; forward 1x1 timelines
set -u
out=gpurun_out/r02p
mkdir -p "$out"
timeout 600 python -m pytest tests/test_gpu_conv.py -x -q -k "fused_wgrad" > "$out/pytest_conv.log" 2>&1; tail -3 "$out/pytest_conv.log"
timeout 600 python -m pytest tests/test_gpu_densenet.py tests/test_gpu_aa_regime.py -x -q > "$out/pytest_net.log" 2>&1; tail -4 "$out/pytest_net.log"
timeout 120 python bench.py --quick --steps 10 --warmup 3 > "$out/bench_quick.json" 2> "$out/bench_quick.err"; echo "step $(cat $out/bench_quick.json | python -c 'import json,sys; d=json.load(sys.stdin); print(d["ms_per_step"], d["value"])')"
{
python tools/bench_conv.py one dgrad 128 128 128 128 224 256 1 2 1
python tools/bench_conv.py one dgrad 128 32 32 128 992 1024 1 2 1
python tools/bench_conv.py one fwd 128 128 128 224 256 128 1 1 1
python tools/bench_conv.py one fwd 128 128 128 64 256 128 1 1 1
python tools/bench_conv.py one fwd 128 32 32 992 1024 128 1 1 1
python tools/bench_conv.py one fwd 128 32 32 512 1024 128 1 1 1
} > "$out/times.log" 2>&1; cat "$out/times.log"
RXB_DBG_TIMELINE=1 python tools/bench_conv.py one fwd 128 128 128 224 256 128 1 1 1 > "$out/timeline_fwd224.log" 2>&1; head -38 "$out/timeline_fwd224.log"
RXB_DBG_TIMELINE=1 python tools/bench_conv.py one fwd 128 128 128 64 256 128 1 1 1 > "$out/timeline_fwd64.log" 2>&1; head -38 "$out/timeline_fwd64.log"
