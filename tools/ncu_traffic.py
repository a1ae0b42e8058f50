#!/usr/bin/env python
"""Summarise `ncu --set full` reports (read here, no GPU needed): per captured launch the duration, DRAM bytes read and
written, DRAM / L1TEX / L2 / tensor-pipe utilisation and registers — written as a text table and merged into
profiles/ncu_traffic.json (the `traffic` numbers bench.py reports).

    python tools/ncu_traffic.py --out profiles/r02_ncu_summary.txt [--batch 128 --conv gpurun_out/x/prof_conv_b128.ncu-rep] rep1 rep2 ..."""
import argparse
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
           ("FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
           ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex_pct"),
           ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
           ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"),
           ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
           ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
           ("smsp__inst_executed.sum", "warp_insts")]
UNIT = {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")}
        for m, key in METRICS:
            if m in hdr:
                i = hdr.index(m)
                try:
                    v = float(r[i].replace(",", "")) if r[i] not in ("", "n/a") else None
                except ValueError:      # 'no data'
                    v = None
                if v is not None and units[i] in UNIT:
                    v *= UNIT[units[i]]
                d[key] = v
        res.append(d)
    return res


# order of the launches of `tools/bench_conv.py prof B`
CONV_ORDER = ["fwd_3x3", "fwd_1x1", "dgrad_3x3_bn", "dgrad_1x1_bn_accum", "wgrad_3x3", "wgrad_1x1",
              "dgrad_3x3_bn_wgrad", "dgrad_1x1_bn_accum_wgrad"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("reps", nargs="*")
    ap.add_argument("--conv", help="report of `tools/bench_conv.py prof B` (six launches)")
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--out", required=True)
    args = ap.parse_args()
    lines = ["%-44s %10s %11s %11s %7s %7s %7s %7s %5s" % ("kernel", "time us", "dram rd MB", "dram wr MB", "dram%", "l1tex%",
                                                             "l2%", "tensor%", "regs")]
    fmt = lambda d: "%-44s %10.1f %11.1f %11.1f %7.1f %7.1f %7.1f %7.1f %5d" % (
        d["kernel"][:44], d["duration"] * 1e6, d["dram_read"] / 1e6, d["dram_write"] / 1e6, d.get("dram_pct") or 0,
        d.get("l1tex_pct") or 0, d.get("l2_pct") or 0, d.get("tensor_pct") or 0, int(d.get("regs") or 0))
    tab_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    tab = json.load(open(tab_path)) if os.path.exists(tab_path) else {}
    for rep in args.reps:
        lines.append("# " + os.path.basename(rep))
        for d in rows_of(rep):
            lines.append(fmt(d))
            tab.setdefault("other", {})[d["kernel"].split("<")[0]] = {
                "dram_bytes": d["dram_read"] + d["dram_write"], "duration_us": d["duration"] * 1e6, "source": os.path.basename(rep)}
    if args.conv:
        lines.append("# %s (tools/bench_conv.py prof %d: %s)" % (os.path.basename(args.conv), args.batch, ", ".join(CONV_ORDER)))
        rows = rows_of(args.conv)
        kern = {}
        for name, d in zip(CONV_ORDER, rows):
            lines.append(fmt(d) + "   " + name)
            kern[name] = int(d["dram_read"] + d["dram_write"])
        tab["batch_%d" % args.batch] = {"source": os.path.basename(args.conv) + " (ncu --set full, one launch each, tools/bench_conv.py prof %d)" % args.batch,
                                        "kernels": kern}
    open(args.out, "w").write("\n".join(lines) + "\n")
    json.dump(tab, open(tab_path, "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
