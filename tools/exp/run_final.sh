timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_final.log 2>&1; tail -2 gpurun_out/t_final.log
timeout 300 python tools/bench_conv.py prof 128 > gpurun_out/prof_plain.log 2>&1 || { tail -5 gpurun_out/prof_plain.log; exit 1; }
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:conv_ -o gpurun_out/prof_conv_b128 -f python tools/bench_conv.py prof 128 > gpurun_out/ncu_conv_b128.log 2>&1
timeout 300 python bench.py --quick --steps 2 --warmup 2 > gpurun_out/plain.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches9.csv python bench.py --quick --steps 1 --warmup 1 > gpurun_out/ncu.log 2>&1
wc -l gpurun_out/launches9.csv
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err; cut -c1-200 gpurun_out/bench_final.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; cut -c1-160 gpurun_out/bench_final_ref.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
