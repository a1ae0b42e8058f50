timeout 300 python bench.py --quick --steps 1 --warmup 1 > gpurun_out/plain.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/launches3.csv python bench.py --quick --steps 1 --warmup 1 > gpurun_out/ncu.log 2>&1
wc -l gpurun_out/launches3.csv
