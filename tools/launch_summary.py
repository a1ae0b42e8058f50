"""Development aid: summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python tools/launch_summary.py launches.csv [--last-step]  -> per-kernel totals, and for the conv kernels a
per-dense-block table (block index inferred from grid/order is not available: grouped by duration pattern)."""
import csv, collections, re, sys
lines = open(sys.argv[1]).read().splitlines()
i = [n for n, l in enumerate(lines) if l.startswith('"ID"')][0]
rows = list(csv.DictReader(lines[i:]))
# keep the last full step: from the last loader_kernel launch backwards one step (loader is first kernel of a step)
idx = [n for n, r in enumerate(rows) if 'loader_kernel' in r['Kernel Name']]
if len(idx) >= 2:
    rows = rows[idx[-2]:idx[-1]]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    k = re.sub(r'\(.*', '', r['Kernel Name']).replace('void ', '').replace('rxb::', '')
    agg[k][0] += 1
    agg[k][1] += float(r['Metric Value']) / 1e3
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-50s %5d %10.1f us %5.1f%%" % (k[:50], v[0], v[1], 100 * v[1] / tot))
print("total %.1f us over %d launches" % (tot, len(rows)))
