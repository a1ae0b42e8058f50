#!/usr/bin/env bash
# Round 2, second GPU call: the full parity suite (first call stopped at a test bug), the conv-kernel changes measured
# (packed-bf16 fold restored, four TMEM stages for the narrow epilogue, degenerate-channel pre-pass), switches.
set -u
out=gpurun_out/r02b
mkdir -p "$out"
( time timeout 900 python -m pytest tests -m gpu -q -s ) > "$out/pytest_gpu.log" 2>&1; echo "pytest rc=$?"; tail -5 "$out/pytest_gpu.log"
RXB_FOLD_FP32=1 timeout 200 python -m pytest tests/test_gpu_aa_regime.py tests/test_gpu_densenet.py -m gpu -q -s > "$out/regime_fold_fp32.log" 2>&1; echo "fold32 rc=$?"; grep -E "512x512|deviation|logits rel|passed|failed" "$out/regime_fold_fp32.log"
for v in "default:" "nx4:RXB_DBG_NX=4" "nograph:RXB_NO_GRAPH=1"; do
  name=${v%%:*}; envs=${v#*:}; extra=""; [ "$name" = "nx3_b256" ] && extra="--batch 256"
  env $envs timeout 90 python bench.py --quick --steps 8 --warmup 3 $extra > "$out/bench_quick_$name.json" 2>/dev/null; echo "$name $(cat $out/bench_quick_$name.json | python -c 'import json,sys; d=json.load(sys.stdin); print(d["ms_per_step"], d["value"])')"
done
RXB_DBG_NX=3 timeout 240 python bench.py --graph --no-cpu-baseline --no-library-baseline > "$out/bench_nx3.json" 2> "$out/bench_nx3.err"; echo "bench nx3 rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02b/bench_nx3.json").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "graph", d.get("cuda_graph_experiment"))
    print({k: round(v["ms_per_step"], 2) for k, v in d["kernel_breakdown_fine"].items()})
    print({k: (round(v["frac"], 3), round(v["ms"], 3)) for k, v in d["conv_kernels"].items()})
except Exception as e:
    print("bench line unreadable:", e)
PY
for spec in "fwd 128 128 128 128 128 32 3 1 1" "dgrad 128 128 128 32 128 128 3 0"; do
  echo "== $spec" >> "$out/timelines.log"
  RXB_DBG_NX=3 RXB_DBG_TIMELINE=1 timeout 60 python tools/bench_conv.py one $spec >> "$out/timelines.log" 2>&1
done
ls "$out"
