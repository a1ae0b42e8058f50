"""Development aid: top SASS instructions by warp-stall samples from `ncu --page source --csv` output.
usage: ncu -i rep --page source --csv --launch-skip K --launch-count 1 > src.csv ; python tools/ncu_top.py src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) >= len(hdr) - 2 and r[col["# Samples"]].isdigit()]
tot = sum(int(r[col["# Samples"]] or 0) for r in data)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
order = sorted(range(len(data)), key=lambda i: -int(data[i][col["# Samples"]] or 0))
print("total samples", tot, "instructions", len(data))
for i in order[:n]:
    r = data[i]
    st = sorted(((int(r[col[s]] or 0), s) for s in stall_cols), reverse=True)[:3]
    print("%5d %6.2f%% exec=%-9s %-60s %s" % (i, 100.0 * int(r[col["# Samples"]]) / max(tot, 1), r[col["Instructions Executed"]],
                                          r[col["Source"]].strip()[:60], " ".join("%s=%d" % (s[6:], v) for v, s in st if v)))
