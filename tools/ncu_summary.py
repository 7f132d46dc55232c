#!/usr/bin/env python
"""Summarise Nsight Compute exports into the small, reviewable tables kept under profiles/.

  tools/ncu_summary.py launches <launches.csv> [traffic.json nx ny nz]
                                                        # per-kernel totals / shares (and DRAM bytes per launch) of a
                                                        # --metrics launch list; optional per-class traffic JSON for bench.py
  tools/ncu_summary.py full <raw.csv>                   # one line per profiled launch of an `ncu --page raw --csv` dump
"""
import collections
import csv
import os
import sys

FULL_COLS = [
    ("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time_us"), ("dram__bytes_read.sum", "dram_rd_MB"),
    ("dram__bytes_write.sum", "dram_wr_MB"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%peak"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_%"), ("launch__registers_per_thread", "regs"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_%"), ("lts__t_sector_hit_rate.pct", "l2_hit_%"),
    ("smsp__inst_executed.sum", "warp_inst"), ("launch__grid_size", "grid"),
]


CLASS_OF = [   # kernel class of the library's table (sipb_kernel_class_name) <- kernel function prefix
    ("yl_update_fused", "k_yl_multi<"), ("yl_update_pass1", "k_yl<float, 1,"), ("yl_update_pass1", "k_yl_spec<"), ("yl_update_pass2", "k_yl<float, 2,"),
    ("cds_spmv_dot", "k_spmv<float, 1>"), ("cds_spmv_dot", "k_spmv_tile<float, 1,"), ("cg_init", "k_cg_init<"),
    ("cg_init", "k_spmv_tile<float, 2,"), ("cg_update_xr", "k_cg_xr<"),
    ("cg_update_p", "k_cg_p<"), ("rhs_compose", "k_rhs<"), ("stop_reduce", "k_stop<"), ("l1_threshold_pass", "k_l1_pass<"),
]


def launches(path, traffic_json=None, grid=None):
    """Launch list of `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv`."""
    rows = list(csv.reader(open(path, errors="ignore")))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    kn, mn, mv, idc = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    agg = collections.defaultdict(lambda: {"ids": set(), "ns": 0.0, "rd": 0.0, "wr": 0.0})
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        name = r[kn].replace("void ", "")
        name = name[:name.index("(")] if "(" in name else name
        a = agg[name]
        a["ids"].add(r[idc])
        v = float(r[mv].replace(",", ""))
        if r[mn] == "gpu__time_duration.sum":
            a["ns"] += v
        elif r[mn] == "dram__bytes_read.sum":
            a["rd"] += v
        elif r[mn] == "dram__bytes_write.sum":
            a["wr"] += v
    tot = sum(v["ns"] for v in agg.values())
    has_dram = any(v["rd"] or v["wr"] for v in agg.values())
    print("| kernel | launches | total us | share |" + (" DRAM rd MB/launch | DRAM wr MB/launch |" if has_dram else "") +
          "\n|---|---:|---:|---:|" + ("---:|---:|" if has_dram else ""))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
        n = len(v["ids"])
        line = "| `%s` | %d | %.1f | %.1f %% |" % (k, n, v["ns"] / 1e3, 100 * v["ns"] / tot)
        if has_dram:
            line += " %.1f | %.1f |" % (v["rd"] / n / 1e6, v["wr"] / n / 1e6)
        print(line)
    print("\ntotal %.1f us over %d launches" % (tot / 1e3, sum(len(v["ids"]) for v in agg.values())))
    if traffic_json and has_dram:
        import json
        out = {"workload": os.environ.get("SIPB_TRAFFIC_WORKLOAD", "config3"), "grid": grid, "source": path, "note": "dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged "
               "over all launches of the kernel class in the launch list (cold-cache, serialised replays)", "kernels": {}}
        for cls in dict.fromkeys(c for c, _ in CLASS_OF):
            prefixes = [pf for c, pf in CLASS_OF if c == cls]
            sel = [v for k, v in agg.items() if any(k.startswith(pf) for pf in prefixes)]
            n = sum(len(v["ids"]) for v in sel)
            if n:
                out["kernels"][cls] = {"dram_bytes_per_launch": sum(v["rd"] + v["wr"] for v in sel) / n, "launches": n,
                                       "avg_us": sum(v["ns"] for v in sel) / n / 1e3}
        json.dump(out, open(traffic_json, "w"), indent=1)


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
    if sys.argv[1] == "launches":      # launches <csv> [traffic.json nx ny nz]
        extra = sys.argv[3:]
        launches(sys.argv[2], extra[0] if extra else None, [int(v) for v in extra[1:4]] if len(extra) >= 4 else None)
    else:
        full(sys.argv[2])
