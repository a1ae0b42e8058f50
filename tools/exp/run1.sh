timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_all.log 2>&1; tail -3 gpurun_out/t_all.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench22.json 2> gpurun_out/bench22.err; cut -c1-200 gpurun_out/bench22.json; tail -3 gpurun_out/bench22.err
