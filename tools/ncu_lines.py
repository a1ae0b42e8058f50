"""Development aid: attribute ncu warp-stall samples to CUDA source lines.
   usage: python tools/ncu_lines.py <ncu-rep> <launch-index> <kernel-mangled-substring> [topN]
   Joins `ncu --page source --csv` (SASS order) with `nvdisasm -g` line info of the in-tree librxb.so."""
import csv, io, os, re, subprocess, sys, tempfile, glob

rep, launch, ksub = sys.argv[1], int(sys.argv[2]), sys.argv[3]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 30
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "recursion_cellular_image_classification_b200", "librxb.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
lines = None
for cub in glob.glob(os.path.join(tmp, "*.cubin")):
    out = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
    cur, fn, seq = None, None, {}
    for l in out.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", l)
        if m:
            fn = m.group(1); seq[fn] = []; continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
        if fn and re.match(r"\s*/\*[0-9a-f]+\*/", l):
            seq[fn].append(cur)
    for fn, s in seq.items():
        if ksub in fn:
            lines = s; kname = fn
if lines is None:
    sys.exit("kernel not found")
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(launch), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]; col = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) >= len(hdr) - 2 and r[col["# Samples"]].isdigit()]
data = data[:len(lines)]
print(rows[0][1][:100], "| sass instrs", len(lines), "csv rows", len(data))
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {}
for i, r in enumerate(data):
    key = lines[i]
    a = agg.setdefault(key, {"n": 0, "exec": 0, "st": {}})
    a["n"] += int(r[col["# Samples"]]); a["exec"] += int(r[col["Instructions Executed"]] or 0)
    for s in stall_cols:
        v = int(r[col[s]] or 0)
        if v: a["st"][s[6:]] = a["st"].get(s[6:], 0) + v
tot = sum(a["n"] for a in agg.values())
src = {}
for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["n"])[:topn]:
    f, ln = key if key else ("?", 0)
    if f not in src:
        p = os.path.join(ROOT, "recursion_cellular_image_classification_b200", "csrc", f)
        src[f] = open(p).read().splitlines() if os.path.exists(p) else []
    text = src[f][ln - 1].strip()[:70] if 0 < ln <= len(src[f]) else ""
    st = " ".join("%s=%d" % kv for kv in sorted(a["st"].items(), key=lambda kv: -kv[1])[:3])
    print("%5.1f%% %-14s:%-4d exec=%-9d %-70s %s" % (100.0 * a["n"] / max(tot, 1), f, ln, a["exec"], text, st))
