for i in 1 2 3; do
  RXB_BENCH_FAKE_WORLD=4 RXB_DBG_SYNC=1 timeout 200 python bench.py --quick --steps 60 --warmup 3 > gpurun_out/fake4_$i.json 2> gpurun_out/fake4_$i.err
  echo "run $i rc=$?"; grep -m2 "librxb error\|rxb:" gpurun_out/fake4_$i.err | grep -v raise | cut -c1-260; tail -1 gpurun_out/fake4_$i.json | cut -c1-200
done
