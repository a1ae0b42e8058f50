#!/usr/bin/env bash
# round 2, call ad: occupancy of three element-wise kernels (launch bounds 3 blocks/SM) - ncu durations, A = previous build
set -u
out=gpurun_out/r02ad
mkdir -p "$out"
A=$PWD/tools/exp/ab/librxb_a.so
for arm in a b; do
  if [ $arm = a ]; then export RXB_LIB=$A; else unset RXB_LIB; fi
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"stem_bn" --csv --log-file "$out/$arm.csv" python bench.py --quick --no-graph --steps 2 --warmup 1 > "$out/$arm.log" 2>&1
  python - "$out/$arm.csv" $arm <<'PY'
import csv, sys, collections, re
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
agg = collections.defaultdict(list)
for r in csv.DictReader(lines):
    k = re.sub(r'\(.*', '', r['Kernel Name']).replace('void ', '').replace('rxb::', '')
    agg[k + r['Grid Size']].append(float(r['Metric Value']) / 1e3)
for k, v in sorted(agg.items()):
    print(sys.argv[2], k[:60], 'n', len(v), 'median us %.1f' % sorted(v)[len(v) // 2])
PY
done
