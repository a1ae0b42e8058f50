for i in 1 2; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2951$i bench.py --gpus 4 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n4_$i.json 2> gpurun_out/bench_n4_$i.err
  echo "run $i rc=$?"; grep -m2 "launch failure\|rxb:" gpurun_out/bench_n4_$i.err gpurun_out/bench_n4_$i.json | cut -c1-160; cut -c1-200 gpurun_out/bench_n4_$i.json | tail -1
done
