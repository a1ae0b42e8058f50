timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_densenet.py -x -q -m gpu > gpurun_out/t_all.log 2>&1; tail -4 gpurun_out/t_all.log
for nx in 3 4 2; do echo "== NX $nx"; RXB_DBG_NX=$nx timeout 200 python tools/bench_conv.py 2>&1 | grep dgrd; done > gpurun_out/bc12.log 2>&1
cat gpurun_out/bc12.log
RXB_DBG_TIMELINE=1 timeout 200 python tools/exp/tl.py > gpurun_out/tl5.log 2>&1; grep -E "timeline|it0[6789]" gpurun_out/tl5.log | head -14
