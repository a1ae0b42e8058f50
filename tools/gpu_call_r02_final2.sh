#!/usr/bin/env bash
# round 2, call final2: what the driver runs at round end on the current tree + the launch list of one step
set -u
out=gpurun_out/r02_final2
mkdir -p "$out"
( time timeout 900 python -m pytest tests -q -m gpu ) > "$out/pytest_gpu.log" 2>&1; echo "pytest rc=$?"; tail -4 "$out/pytest_gpu.log"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > "$out/smoke.log" 2>&1; echo "smoke rc=$?"; tail -1 "$out/smoke.log"
( time timeout 500 python bench.py > "$out/bench.json" 2> "$out/bench.err" ); echo "bench rc=$?"; tail -2 "$out/bench.err"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_final2/bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "stream", d["stream_launch_comparison"]["value"], "launches", d["gpu_launches"], d["clocks"])
print("roofline", d["roofline"]["kernel"][:60], round(d["roofline"]["frac"], 3), round(d["roofline"]["share_of_step"], 3), "step", {k: round(v, 3) for k, v in d["step_roofline"].items()})
for k, v in d["kernel_breakdown_fine"].items():
    print("  %-20s %7.3f ms %5d launches  hbm_frac %s" % (k, v["ms_per_step"], v["launches_per_step"], round(v.get("hbm_frac", 0), 3)))
PY
( time timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > "$out/bench_reference.json" 2>/dev/null ); cut -c1-200 "$out/bench_reference.json"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file "$out/launches.csv" python bench.py --quick --no-graph --steps 1 --warmup 1 > "$out/ncu.log" 2>&1
python tools/launch_summary.py "$out/launches.csv" | tee "$out/launch_summary.txt"
