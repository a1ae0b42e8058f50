#!/usr/bin/env bash
# N-GPU validation of the shipped defaults (three in-place dgrad buffers, graph replay per phase) and the scaling line
set -u
N=${1:-8}
out=gpurun_out/r02_n$N
mkdir -p "$out"
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port "$1" "${@:2}"; }
for i in 1 2 3; do
  timeout 200 bash -c "$(declare -f run); N=$N; run $((29600 + i)) bench.py --gpus $N --quick --steps 20 --warmup 3" > "$out/quick_$i.json" 2> "$out/quick_$i.err"; echo "quick $i rc=$? $(cat $out/quick_$i.json | cut -c1-200)"
done
timeout 400 bash -c "$(declare -f run); N=$N; run 29610 bench.py --gpus $N --steps 10 --warmup 3" > "$out/bench.json" 2> "$out/bench.err"; echo "bench rc=$?"
python - "$out/bench.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("n", d["n_gpus"], "value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"], "stream", d.get("stream_launch_comparison"))
except Exception as e:
    print("bench line unreadable:", e)
PY
timeout 200 bash -c "$(declare -f run); N=$N; run 29611 bench.py --impl reference --gpus $N --steps 3 --warmup 1" > "$out/bench_reference.json" 2>/dev/null; cut -c1-300 "$out/bench_reference.json"
grep -l "unspecified launch failure\|watchdog\|illegal" "$out"/*.err 2>/dev/null
