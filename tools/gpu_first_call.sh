#!/usr/bin/env bash
# One gpurun call that re-establishes the state of the tree on a fresh B200 and collects what round 1 left unmeasured
# (see DESIGN.md 8a).  Usage:  gpurun --timeout 420 -- 'bash tools/gpu_first_call.sh'
# Writes everything under gpurun_out/first_call/ ; each step has its own timeout so one overrun does not eat the rest.
set -u
out=gpurun_out/first_call
mkdir -p "$out"
timeout 90 python -m pytest tests -m gpu -x -q > "$out/pytest_gpu.log" 2>&1; echo "pytest rc=$?"; tail -2 "$out/pytest_gpu.log"
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > "$out/smoke.log" 2>&1; echo "smoke rc=$?"; tail -1 "$out/smoke.log"
# headline bench + the CUDA-graph experiment (stream launches vs one graph per step)
timeout 150 python bench.py --graph > "$out/bench_graph.json" 2> "$out/bench_graph.err"; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/first_call/bench_graph.json").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "graph", d.get("cuda_graph_experiment"))
except Exception as e:
    print("bench line unreadable:", e)
PY
# BASELINE config 5 at a tenth of the corpus (5 experiments): statistics / normalisation GB/s on a streamed corpus
timeout 60 python tools/corpus_sweep.py --experiments 5 > "$out/corpus_sweep.json" 2> "$out/corpus_sweep.err"; echo "sweep rc=$?"; cat "$out/corpus_sweep.json"
# ncu: the all-lanes JPEG kernel (round 1 captured only the affine loader)
timeout 60 ncu --set full --clock-control none --import-source on -k regex:jpeg_decode_par_kernel -s 2 -c 1 \
  -o "$out/prof_jpeg_par" python tools/widen_check.py --no-tests > "$out/ncu_jpeg.json" 2> "$out/ncu_jpeg.log"; echo "ncu rc=$?"
# train() itself on two ranks (needs `gpurun --gpus 2`; skipped on a one-GPU box)
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    tools/train_two_ranks.py > "$out/train_two_ranks.json" 2> "$out/train_two_ranks.err"; echo "train ranks rc=$?"; cat "$out/train_two_ranks.json"
fi
