timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_all.log 2>&1; tail -3 gpurun_out/t_all.log
for v in 0 1; do RXB_DBG_NO_XMERGE=$v timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench16_$v.json 2> gpurun_out/bench16.err; cut -c1-200 gpurun_out/bench16_$v.json; tail -3 gpurun_out/bench16.err; done
