timeout 300 python tools/bench_conv.py prof 128 > gpurun_out/prof_plain.log 2>&1 || { tail -5 gpurun_out/prof_plain.log; exit 1; }
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:conv_ -o gpurun_out/prof_conv_b128 -f python tools/bench_conv.py prof 128 > gpurun_out/ncu_conv_b128.log 2>&1
tail -2 gpurun_out/ncu_conv_b128.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench11.json 2> gpurun_out/bench11.err; tail -3 gpurun_out/bench11.err; cut -c1-300 gpurun_out/bench11.json
