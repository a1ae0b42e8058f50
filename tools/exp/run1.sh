timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench11.json 2> gpurun_out/bench11.err; tail -3 gpurun_out/bench11.err; cut -c1-200 gpurun_out/bench11.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench11_ref.json 2> gpurun_out/bench11_ref.err; tail -3 gpurun_out/bench11_ref.err; cut -c1-300 gpurun_out/bench11_ref.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
