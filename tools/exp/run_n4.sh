timeout 300 python tools/bench_conv.py 2>&1 | grep dgrd
for i in 1 2 3; do
  timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2958$i bench.py --gpus 4 --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/nx4_n4_$i.json 2> gpurun_out/nx4_n4_$i.err
  echo "NX=4 run $i rc=$?"; grep -m2 "librxb error\|launch failure" gpurun_out/nx4_n4_$i.err | grep -v raise | cut -c1-200; tail -1 gpurun_out/nx4_n4_$i.json | cut -c1-150
done
