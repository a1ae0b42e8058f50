#!/usr/bin/env bash
# round 2, call y (gpurun --gpus N): the fused backward kernels on N ranks - train() replica equality, 3 quick bench runs
# (the round-1 fault showed only at N=4), the bench line at N
set -u
N=${1:-4}
out=gpurun_out/r02y_n$N
mkdir -p "$out"
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port "$1" "${@:2}"; }
timeout 180 bash -c "$(declare -f run); N=$N; run 29511 tests/multi_gpu/run_train_ranks.py" > "$out/train_ranks_64.json" 2> "$out/train_ranks_64.err"; echo "train 64 rc=$?"; tail -c 500 "$out/train_ranks_64.json"
for i in 1 2 3; do
  timeout 200 bash -c "$(declare -f run); N=$N; run $((29520 + i)) bench.py --gpus $N --quick --steps 10 --warmup 3" > "$out/bench_quick_$i.json" 2> "$out/bench_quick_$i.err"; echo "quick $i rc=$? $(cut -c1-160 $out/bench_quick_$i.json) $(tail -1 $out/bench_quick_$i.err | cut -c1-160)"
done
timeout 400 bash -c "$(declare -f run); N=$N; run 29516 bench.py --gpus $N --steps 10 --warmup 3" > "$out/bench.json" 2> "$out/bench.err"; echo "bench rc=$?"
python - "$out/bench.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("n", d["n_gpus"], "value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"], d["clocks"])
except Exception as e:
    print("bench line unreadable:", e)
PY
