#!/usr/bin/env python
"""Summarise Nsight Compute exports into the small, reviewable tables kept under profiles/.

  tools/ncu_summary.py launches <launches.csv>          # per-kernel totals / shares of a --metrics launch list
  tools/ncu_summary.py full <raw.csv>                   # one line per profiled launch of an `ncu --page raw --csv` dump
"""
import collections
import csv
import sys

FULL_COLS = [
    ("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time_us"), ("dram__bytes_read.sum", "dram_rd_MB"),
    ("dram__bytes_write.sum", "dram_wr_MB"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%peak"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_%"), ("launch__registers_per_thread", "regs"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_%"), ("lts__t_sector_hit_rate.pct", "l2_hit_%"),
    ("smsp__inst_executed.sum", "warp_inst"), ("launch__grid_size", "grid"),
]


def launches(path):
    rows = list(csv.reader(open(path, errors="ignore")))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        name = r[kn].split("(")[0].replace("void ", "")
        agg[name][0] += 1
        agg[name][1] += float(r[mv].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.1f %% |" % (k, v[0], v[1] / 1e3, 100 * v[1] / tot))
    print("\ntotal %.1f us over %d launches" % (tot / 1e3, sum(v[0] for v in agg.values())))


def full(path):
    rows = list(csv.reader(open(path, errors="ignore")))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, units = rows[hi], rows[hi + 1]
    idx = [(hdr.index(c) if c in hdr else None, lab) for c, lab in FULL_COLS]
    print("| " + " | ".join(lab for _, lab in idx) + " |\n|" + "---|" * len(idx))
    for r in rows[hi + 2:]:
        out = []
        for i, lab in idx:
            if i is None or i >= len(r):
                out.append("")
                continue
            v = r[i]
            if lab == "kernel":
                v = "`" + v.split("(")[0].replace("void ", "") + "`"
            else:
                try:
                    f = float(v.replace(",", ""))
                    v = ("%.0f" % f) if abs(f) >= 1000 else ("%.2f" % f)
                except ValueError:
                    pass
            out.append(v)
        print("| " + " | ".join(out) + " |")
    print("\nunits: " + ", ".join("%s [%s]" % (lab, units[i]) for i, lab in idx if i is not None and units[i]))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
