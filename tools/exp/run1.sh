timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_all.log 2>&1; tail -5 gpurun_out/t_all.log
timeout 200 python tools/bench_conv.py 2>&1 | grep -E "k3|k4" > gpurun_out/bc6.log 2>&1
cat gpurun_out/bc6.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench6.json 2> gpurun_out/bench6.err; cat gpurun_out/bench6.json | cut -c1-300; tail -3 gpurun_out/bench6.err
