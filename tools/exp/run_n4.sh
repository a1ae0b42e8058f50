RXB_DBG_NX=3 timeout 300 python -m pytest tests/test_gpu_conv.py tests/test_gpu_densenet.py -x -q -m gpu 2>&1 | tail -2
for i in 1 2 3; do
  RXB_DBG_NX=3 timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2959$i bench.py --gpus 4 --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/nx3b_n4_$i.json 2> gpurun_out/nx3b_n4_$i.err
  echo "NX=3 split-barrier run $i rc=$?"; grep -m2 "librxb error\|launch failure" gpurun_out/nx3b_n4_$i.err | grep -v raise | cut -c1-200; tail -1 gpurun_out/nx3b_n4_$i.json | cut -c1-150
done
