timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_final.log 2>&1; tail -2 gpurun_out/t_final.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err; cut -c1-200 gpurun_out/bench_final.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
