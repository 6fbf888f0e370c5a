#!/usr/bin/env python
"""Reduce the ncu CSVs under profiles/ to one table (profiles/<tag>_summary.md + <tag>_summary.json).
   python scripts/summarize_profiles.py r01
Inputs (written by scripts/collect_profiles.sh): <tag>_launches_cfg2_1gpu.csv (gpu__time_duration per launch),
<tag>_ncu_full_<kernel>.csv (--set full, --page raw), <tag>_dram_split_search_all_levels.csv (DRAM bytes of every
split-search launch of one CV), <tag>_bench_cfg2_1gpu.json (the un-profiled bench line)."""
import collections
import csv
import glob
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")


def short(name):
    name = re.sub(r"\(.*", "", name)
    return re.sub(r"<.*", "", name).replace("void ", "").strip()


def launches(tag):
    path = os.path.join(P, tag + "_launches_cfg2_1gpu.csv")
    if not os.path.exists(path):
        return {}
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    h = rows[0]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for r in rows[1:]:
        k = short(r[ki])
        tot[k] += float(r[vi].replace(",", "")) / 1e6
        cnt[k] += 1
    return {k: {"launches": cnt[k], "ms": tot[k]} for k in tot}


def full(tag):
    out = {}
    cols = {"gpu__time_duration.sum": "us", "dram__bytes_read.sum": "dram_read_MB", "dram__bytes_write.sum": "dram_write_MB",
            "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
            "launch__registers_per_thread": "regs", "smsp__inst_executed.sum": "inst", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1tex_pct",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct"}
    scale = {"us": {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}, "MB": {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}}
    for path in sorted(glob.glob(os.path.join(P, tag + "_ncu_full_*.csv"))):
        rows = list(csv.reader(open(path)))
        h, units = rows[0], rows[1]
        kern = os.path.basename(path)[len(tag) + 10:-4]
        recs = []
        for r in rows[2:]:
            d = {"grid": r[h.index("Grid Size")]}
            for c, nm in cols.items():
                if c in h:
                    v = float(r[h.index(c)].replace(",", "")); u = units[h.index(c)]
                    if nm == "us":
                        v *= scale["us"].get(u, 1)
                    elif nm.endswith("_MB"):
                        v *= scale["MB"].get(u, 1)
                    d[nm] = round(v, 3)
            if "us" in d and "dram_read_MB" in d:
                d["dram_GBps"] = round((d["dram_read_MB"] + d["dram_write_MB"]) / d["us"] * 1e3, 1)
            recs.append(d)
        out[kern] = recs
    return out


def dram_levels(tag):
    path = os.path.join(P, tag + "_dram_split_search_all_levels.csv")
    if not os.path.exists(path):
        return None
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    h = rows[0]
    d = collections.defaultdict(dict)
    for r in rows[1:]:
        d[int(r[0])][r[h.index("Metric Name")]] = float(r[h.index("Metric Value")].replace(",", ""))
    n = len(d)
    rd = sum(v.get("dram__bytes_read.sum", 0) for v in d.values()); wr = sum(v.get("dram__bytes_write.sum", 0) for v in d.values())
    ns = sum(v.get("gpu__time_duration.sum", 0) for v in d.values())
    return {"launches": n, "dram_read_bytes": rd, "dram_write_bytes": wr, "traffic_bytes_per_launch": (rd + wr) / max(n, 1), "ms_under_ncu": ns / 1e6}


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    L, F, D = launches(tag), full(tag), dram_levels(tag)
    bench = None
    bp = os.path.join(P, tag + "_bench_cfg2_1gpu.json")
    if os.path.exists(bp):
        bench = json.loads(open(bp).read().strip().splitlines()[-1])
    summary = {"tag": tag, "launch_list": L, "full": F, "split_search_dram": D}
    json.dump(summary, open(os.path.join(P, tag + "_summary.json"), "w"), indent=1)
    md = ["# %s profile summary (BASELINE configs[1], one B200)\n" % tag]
    if bench:
        r = bench["roofline"]
        md += ["Bench line (no profiler): %.1f ms per CV, %.0f partitions/s; e2e %.0f partitions/s; split search %.0f GB/s algorithmic = %.3f of %.0f GB/s.\n"
               % (bench["ms_per_step"], bench["value"], bench["e2e"]["value"], r["achieved"], r["frac"], r["peak"])]
    if L:
        tot = sum(v["ms"] for v in L.values())
        md += ["## Launch list (ncu gpu__time_duration, serialised and cold-cache: compare SHARES with the live event timing)\n",
               "| kernel | launches | ms | share | live ms (bench, events) |", "|---|---|---|---|---|"]
        live = bench["roofline"]["kernel_ms_per_step"] if bench else {}
        m2 = {"split_search_kernel": "split_search", "split_search_merged_kernel": "split_search", "build_lists_multi_kernel": "build_lists", "partition_lists_kernel": "partition_lists", "feature_sort_split_kernel": "feature_sort", "predict_compact_kernel": "predict",
              "build_lists_kernel": "build_lists", "level_finalize_kernel": "level_finalize", "tree_level_kernel": "level_finalize", "split_select_kernel": "split_select",
              "feature_sort_radix_kernel": "feature_sort", "bootstrap_kernel": "bootstrap", "assemble_kernel": "assemble"}
        for k in sorted(L, key=lambda k_: -L[k_]["ms"]):
            lv = live.get(m2.get(k, ""), None)
            md.append("| %s | %d | %.2f | %.1f%% | %s |" % (k, L[k]["launches"], L[k]["ms"], 100 * L[k]["ms"] / tot, ("%.2f" % lv) if lv is not None else ""))
        md.append("")
    if D:
        md += ["## Split search DRAM traffic over one CV (%d launches)\n" % D["launches"],
               "read %.2f GB, written %.3f GB -> traffic per launch %.1f MB; algorithmic bytes per launch (SURVEY 8d, fp64 units) %.1f MB: the compact 4-byte "
               "list entries move a third of the contract's byte count and nothing is re-read.\n" % (D["dram_read_bytes"] / 1e9, D["dram_write_bytes"] / 1e9, D["traffic_bytes_per_launch"] / 1e6,
                                                                                                    bench["roofline"]["algorithmic_bytes_per_launch"] / 1e6 if bench else float("nan"))]
    for kern, recs in F.items():
        md += ["## ncu --set full: %s\n" % kern, "| grid | us | DRAM GB/s | issue active % | warps active % | regs | warp instructions | L1 % | L2 % |", "|---|---|---|---|---|---|---|---|---|"]
        for d in recs:
            md.append("| %s | %.1f | %s | %.1f | %.1f | %d | %.3g | %.1f | %.1f |" % (d["grid"], d.get("us", 0), d.get("dram_GBps", ""), d.get("issue_active_pct", 0), d.get("warps_active_pct", 0),
                                                                            d.get("regs", 0), d.get("inst", 0), d.get("l1tex_pct", 0), d.get("l2_pct", 0)))
        md.append("")
    open(os.path.join(P, tag + "_summary.md"), "w").write("\n".join(md) + "\n")
    print("\n".join(md))


if __name__ == "__main__":
    main()
