timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_all.log 2>&1; tail -5 gpurun_out/t_all.log
for v in "" "RXB_DBG_NO_RESIDENT=2"; do echo "== $v"; env $v timeout 200 python tools/bench_conv.py 2>&1; done > gpurun_out/bc4.log 2>&1
cat gpurun_out/bc4.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench4.json 2> gpurun_out/bench4.err; cat gpurun_out/bench4.json | cut -c1-300; tail -3 gpurun_out/bench4.err
