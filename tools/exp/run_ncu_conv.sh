timeout 300 python tools/bench_conv.py prof > gpurun_out/prof_plain.log 2>&1 || { tail -5 gpurun_out/prof_plain.log; exit 1; }
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:conv_ -o gpurun_out/prof_conv7 -f python tools/bench_conv.py prof > gpurun_out/ncu_conv7.log 2>&1
tail -3 gpurun_out/ncu_conv7.log; ls -la gpurun_out/prof_conv7.ncu-rep
