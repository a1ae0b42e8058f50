#!/usr/bin/env bash
# Round 2, multi-GPU call (gpurun --gpus N): the reference-facing entry points themselves on N ranks —
# train() (replica equality + images/s through the DataLoader), test() (wells sharded, logits all-gathered),
# the config-5 statistics/normalisation sweep sharded by experiment, and the bench line at N.
set -u
N=${1:-2}
out=gpurun_out/r02_ranks_n$N
mkdir -p "$out"
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port "$1" "${@:2}"; }
timeout 180 bash -c "$(declare -f run); N=$N; run 29511 tests/multi_gpu/run_train_ranks.py" > "$out/train_ranks_64.json" 2> "$out/train_ranks_64.err"; echo "train 64 rc=$?"; tail -c 600 "$out/train_ranks_64.json"
timeout 300 bash -c "$(declare -f run); N=$N; run 29512 tests/multi_gpu/run_train_ranks.py --size 512 --samples $((256 * N)) --bs $((64 * N)) --decode gpu --workers 4" > "$out/train_ranks_512_gpu_decode.json" 2> "$out/train_ranks_512_gpu_decode.err"; echo "train 512 gpu-decode rc=$?"; tail -c 400 "$out/train_ranks_512_gpu_decode.json"
timeout 300 bash -c "$(declare -f run); N=$N; run 29513 tests/multi_gpu/run_train_ranks.py --size 512 --samples $((256 * N)) --bs $((64 * N)) --decode host --workers 8" > "$out/train_ranks_512_host_decode.json" 2> "$out/train_ranks_512_host_decode.err"; echo "train 512 host-decode rc=$?"; tail -c 400 "$out/train_ranks_512_host_decode.json"
timeout 240 bash -c "$(declare -f run); N=$N; run 29514 tests/multi_gpu/run_test_ranks.py" > "$out/test_ranks.json" 2> "$out/test_ranks.err"; echo "test ranks rc=$?"; cat "$out/test_ranks.json"
timeout 240 bash -c "$(declare -f run); N=$N; run 29517 tests/multi_gpu/run_stats_ranks.py" > "$out/stats_ranks.json" 2> "$out/stats_ranks.err"; echo "stats ranks rc=$?"; cat "$out/stats_ranks.json"
timeout 240 bash -c "$(declare -f run); N=$N; run 29515 tools/corpus_sweep.py --experiments $((6 * N))" > "$out/corpus_sweep.json" 2> "$out/corpus_sweep.err"; echo "sweep rc=$?"; cat "$out/corpus_sweep.json"
timeout 300 bash -c "$(declare -f run); N=$N; run 29516 bench.py --gpus $N --steps 10 --warmup 3" > "$out/bench.json" 2> "$out/bench.err"; echo "bench rc=$?"
python - "$out/bench.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("n", d["n_gpus"], "value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"])
except Exception as e:
    print("bench line unreadable:", e)
PY
