timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -3 gpurun_out/bench_final.err; cut -c1-200 gpurun_out/bench_final.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; cut -c1-200 gpurun_out/bench_final_ref.json
