#!/usr/bin/env bash
# round 2, call t: the full bench line and the launch list of one step with the fused backward kernels
set -u
out=gpurun_out/r02t
mkdir -p "$out"
( time timeout 500 python bench.py > "$out/bench.json" 2> "$out/bench.err" ); echo "bench rc=$?"; tail -2 "$out/bench.err"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02t/bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "stream", d["stream_launch_comparison"]["value"], "launches", d["gpu_launches"], d["clocks"])
print("roofline", d["roofline"]["kernel"][:60], round(d["roofline"]["frac"], 3), round(d["roofline"]["share_of_step"], 3), "step", {k: round(v, 3) for k, v in d["step_roofline"].items()})
for k, v in d["kernel_breakdown_fine"].items():
    print("  %-20s %7.3f ms %5d launches  hbm_frac %s" % (k, v["ms_per_step"], v["launches_per_step"], round(v.get("hbm_frac", 0), 3)))
PY
timeout 300 python bench.py --quick --no-graph --steps 1 --warmup 1 > "$out/plain.log" 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file "$out/launches.csv" python bench.py --quick --no-graph --steps 1 --warmup 1 > "$out/ncu.log" 2>&1
wc -l "$out/launches.csv"
python tools/launch_summary.py "$out/launches.csv" | tee "$out/launch_summary.txt"
