timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_all.log 2>&1; tail -4 gpurun_out/t_all.log
timeout 200 python tools/bench_conv.py 2>&1 | grep -E "wgrd" > gpurun_out/bc18.log 2>&1; cat gpurun_out/bc18.log
for b in 128; do timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --batch $b > gpurun_out/bench10_$b.json 2> gpurun_out/bench10.err; cat gpurun_out/bench10_$b.json | cut -c1-200; tail -3 gpurun_out/bench10.err; done
