timeout 900 python -m pytest tests -x -q -m gpu -s > gpurun_out/t_all.log 2>&1; tail -5 gpurun_out/t_all.log; grep "loss ours" gpurun_out/t_all.log
timeout 200 python tools/bench_conv.py > gpurun_out/bc5.log 2>&1
cat gpurun_out/bc5.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench5.json 2> gpurun_out/bench5.err; cat gpurun_out/bench5.json | cut -c1-300; tail -3 gpurun_out/bench5.err
