#!/usr/bin/env bash
# round 2, call z: fewer warp-uniform constant loads in the fused epilogues, fold constants of the transform roles in registers
set -u
out=gpurun_out/r02z
mkdir -p "$out"
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_densenet.py tests/test_gpu_aa_regime.py tests/test_gpu_resnet.py -x -q > "$out/pytest.log" 2>&1; tail -4 "$out/pytest.log"
timeout 120 python bench.py --quick --steps 10 --warmup 3 > "$out/bench_quick.json" 2> "$out/bench_quick.err"; echo "step $(cat $out/bench_quick.json | python -c 'import json,sys; d=json.load(sys.stdin); print(d["ms_per_step"], d["value"])') $(tail -1 $out/bench_quick.err | cut -c1-200)"
timeout 120 python bench.py --quick --steps 10 --warmup 3 > "$out/bench_quick2.json" 2> "$out/bench_quick2.err"; echo "step $(cat $out/bench_quick2.json | python -c 'import json,sys; d=json.load(sys.stdin); print(d["ms_per_step"], d["value"])')"
{
python tools/bench_conv.py one dgrad 128 128 128 128 224 256 1 2 1
python tools/bench_conv.py one dgrad 128 128 128 32 128 128 3 0 1
python tools/bench_conv.py one fwd 128 128 128 128 128 32 3 1 1
python tools/bench_conv.py one fwd 128 128 128 128 256 128 1 1 1
} > "$out/times.log" 2>&1; cat "$out/times.log"
