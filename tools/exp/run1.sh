timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_densenet.py -x -q -m gpu > gpurun_out/t_all.log 2>&1; tail -12 gpurun_out/t_all.log
timeout 200 python tools/bench_conv.py 2>&1 | grep -E "^fwd.*k3" > gpurun_out/bc21.log 2>&1; cat gpurun_out/bc21.log
