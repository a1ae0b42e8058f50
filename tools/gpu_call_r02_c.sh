#!/usr/bin/env bash
# Round 2, third GPU call: parity suite on the shipped kernels (three in-place dgrad buffers, four TMEM stages for the
# narrow epilogue, bf16 fold), the full bench line (CUDA-graph replay is the default), ncu captures for profiles/.
set -u
out=gpurun_out/r02c
mkdir -p "$out"
( time timeout 900 python -m pytest tests -m gpu -q -s ) > "$out/pytest_gpu.log" 2>&1; echo "pytest rc=$?"; tail -4 "$out/pytest_gpu.log"
grep -E "512x512|deviation|noise-only|degenerate gammas|config 4" "$out/pytest_gpu.log"
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > "$out/smoke.log" 2>&1; echo "smoke rc=$?"; tail -1 "$out/smoke.log"
timeout 300 python bench.py > "$out/bench.json" 2> "$out/bench.err"; echo "bench rc=$?"; tail -3 "$out/bench.err"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02c/bench.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "stream", d.get("stream_launch_comparison"), "launches", d["gpu_launches"])
    print({k: round(v["ms_per_step"], 2) for k, v in d["kernel_breakdown_fine"].items()})
    print({k: (round(v["frac"], 3), round(v["ms"], 3)) for k, v in d["conv_kernels"].items()})
    print("roofline", d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"]["share_of_step"], d["step_roofline"])
    print("lib", d.get("library_gpu_baseline"), "cpu", d.get("cpu_baseline"))
except Exception as e:
    print("bench line unreadable:", e)
PY
for v in "nx4:RXB_DBG_NX=4" "b256:"; do
  name=${v%%:*}; envs=${v#*:}; extra=""; [ "$name" = "b256" ] && extra="--batch 256"
  env $envs timeout 120 python bench.py --quick --steps 8 --warmup 3 $extra > "$out/bench_quick_$name.json" 2>/dev/null; echo "$name $(cat $out/bench_quick_$name.json | python -c 'import json,sys; d=json.load(sys.stdin); print(d["ms_per_step"], d["value"])')"
done
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"loader_kernel" -s 1 -c 1 \
  -o "$out/prof_loader" python tools/prof_hbm_kernels.py > "$out/ncu_loader.log" 2>&1; echo "ncu loader rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"conv_gemm_kernel|conv_wgrad_kernel" -c 6 \
  -o "$out/prof_conv_b128" python tools/bench_conv.py prof 128 > "$out/ncu_conv.log" 2>&1; echo "ncu conv rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 1400 --csv --log-file "$out/launches_b128.csv" \
  python bench.py --quick --no-graph --steps 1 --warmup 1 > "$out/ncu_launches.log" 2>&1; echo "ncu launches rc=$?"
ls "$out"
