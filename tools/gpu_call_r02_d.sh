#!/usr/bin/env bash
set -u
out=gpurun_out/r02d
mkdir -p "$out"
timeout 300 python bench.py > "$out/bench.json" 2> "$out/bench.err"; echo "bench rc=$?"; tail -3 "$out/bench.err"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02d/bench.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "stream", d.get("stream_launch_comparison"), "launches", d["gpu_launches"])
    print({k: round(v["ms_per_step"], 2) for k, v in d["kernel_breakdown_fine"].items()})
    print({k: (round(v["frac"], 3), round(v["ms"], 3)) for k, v in d["conv_kernels"].items()})
    print("roofline", d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"]["share_of_step"], d["step_roofline"])
    print("lib", d.get("library_gpu_baseline"), "cpu", d.get("cpu_baseline"), d["clocks"])
except Exception as e:
    print("bench line unreadable:", e)
PY
timeout 120 python bench.py --impl reference --steps 3 --warmup 1 > "$out/bench_reference.json" 2>/dev/null; cat "$out/bench_reference.json" | cut -c1-400
