#!/usr/bin/env bash
# round 2, call w: EPI 3 with dedicated transform warps (workers 0-7) and two four-warp epilogue groups
set -u
out=gpurun_out/r02w
mkdir -p "$out"
timeout 600 python -m pytest tests/test_gpu_conv.py -x -q -k "fused_wgrad or dgrad" > "$out/pytest_conv.log" 2>&1; tail -3 "$out/pytest_conv.log"
timeout 600 python -m pytest tests/test_gpu_densenet.py tests/test_gpu_aa_regime.py -x -q > "$out/pytest_net.log" 2>&1; tail -3 "$out/pytest_net.log"
timeout 120 python bench.py --quick --steps 10 --warmup 3 > "$out/bench_quick.json" 2> "$out/bench_quick.err"; echo "step $(cat $out/bench_quick.json | python -c 'import json,sys; d=json.load(sys.stdin); print(d["ms_per_step"], d["value"])') $(tail -1 $out/bench_quick.err | cut -c1-200)"
{
python tools/bench_conv.py one dgrad 128 128 128 128 224 256 1 2 1
python tools/bench_conv.py one dgrad 128 128 128 128 64 256 1 2 1
python tools/bench_conv.py one dgrad 128 64 64 128 480 512 1 2 1
python tools/bench_conv.py one dgrad 128 32 32 128 992 1024 1 2 1
python tools/bench_conv.py one dgrad 128 16 16 128 992 1024 1 2 1
} > "$out/times.log" 2>&1; cat "$out/times.log"
