#!/usr/bin/env bash
# Round 2, first GPU call: parity suite with the new regime tests, headline bench (+ CUDA graph experiment), smoke,
# config-5 corpus sweep, two switches measured (fp32 fold, odd dgrad buffer count), per-kernel timelines, ncu captures.
set -u
out=gpurun_out/r02a
mkdir -p "$out"
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > "$out/gpu.txt" 2>&1
( time timeout 600 python -m pytest tests -m gpu -x -q -s ) > "$out/pytest_gpu.log" 2>&1; echo "pytest rc=$?"; tail -3 "$out/pytest_gpu.log"
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > "$out/smoke.log" 2>&1; echo "smoke rc=$?"; tail -1 "$out/smoke.log"
timeout 240 python bench.py --graph > "$out/bench.json" 2> "$out/bench.err"; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02a/bench.json").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "graph", d.get("cuda_graph_experiment"), "lib", d.get("library_gpu_baseline"))
    print({k: round(v["ms_per_step"], 2) for k, v in d["kernel_breakdown_fine"].items()})
except Exception as e:
    print("bench line unreadable:", e)
PY
timeout 120 python tools/corpus_sweep.py > "$out/corpus_sweep.json" 2> "$out/corpus_sweep.err"; echo "sweep rc=$?"; cat "$out/corpus_sweep.json"
# switches
RXB_FOLD_FP32=1 timeout 120 python -m pytest tests/test_gpu_aa_regime.py -m gpu -q -s -k "512 or trajectory" > "$out/regime_fold_fp32.log" 2>&1; echo "fold32 regime rc=$?"; grep -E "512x512|deviation" "$out/regime_fold_fp32.log"
RXB_FOLD_FP32=1 timeout 90 python bench.py --quick --steps 6 --warmup 3 > "$out/bench_quick_fold_fp32.json" 2>/dev/null; cat "$out/bench_quick_fold_fp32.json"
RXB_DBG_NX=3 timeout 90 python bench.py --quick --steps 6 --warmup 3 > "$out/bench_quick_nx3.json" 2>/dev/null; cat "$out/bench_quick_nx3.json"
timeout 90 python bench.py --quick --steps 6 --warmup 3 > "$out/bench_quick_default.json" 2>/dev/null; cat "$out/bench_quick_default.json"
timeout 90 python bench.py --quick --steps 6 --warmup 3 --batch 256 > "$out/bench_quick_b256.json" 2>/dev/null; cat "$out/bench_quick_b256.json"
# timelines of CTA 0 (cycles), one process per shape: B H W Cin ldA Cout k prologue [stats]
for spec in "fwd 128 128 128 128 128 32 3 1 1" "fwd 128 128 128 224 256 128 1 1 1" "fwd 128 32 32 640 1024 128 1 1 1" \
            "wgrad 128 128 128 224 256 128 1 1" "wgrad 128 128 128 128 128 32 3 1" "wgrad 128 32 32 640 1024 128 1 1" \
            "wgrad 128 16 16 768 1024 128 1 1" "dgrad 128 128 128 32 128 128 3 0" "dgrad 128 32 32 128 640 1024 1 2"; do
  echo "== $spec" >> "$out/timelines.log"
  RXB_DBG_TIMELINE=1 timeout 60 python tools/bench_conv.py one $spec >> "$out/timelines.log" 2>&1
done
echo "timelines done"
# ncu: launch list of one warm step, then full captures of the kernels that carry roofline claims
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 1400 --csv --log-file "$out/launches_b128.csv" \
  python bench.py --quick --steps 1 --warmup 1 > "$out/ncu_launches.log" 2>&1; echo "ncu launches rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"stats_planar_kernel|loader_kernel" -s 2 -c 2 \
  -o "$out/prof_hbm" python tools/prof_hbm_kernels.py > "$out/ncu_hbm.log" 2>&1; echo "ncu hbm rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:bn_bwd_apply_kernel -s 52 -c 1 \
  -o "$out/prof_apply" python bench.py --quick --steps 1 --warmup 1 > "$out/ncu_apply.log" 2>&1; echo "ncu apply rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:jpeg_decode_par_kernel -s 2 -c 1 \
  -o "$out/prof_jpeg" python tools/widen_check.py --no-tests > "$out/ncu_jpeg.log" 2>&1; echo "ncu jpeg rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"conv_gemm_kernel|conv_wgrad_kernel" -c 6 \
  -o "$out/prof_conv_b128" python tools/bench_conv.py prof 128 > "$out/ncu_conv.log" 2>&1; echo "ncu conv rc=$?"
ls -la "$out"
