"""Turns the raw outputs of profiles/prof_r1.sh (in gpurun_out/) into the tracked summaries:

    gpurun_out/launches_r1.csv  -> profiles/launches_r1.csv (copy) + profiles/launches_r1_summary.csv
    gpurun_out/prof_r1_raw.csv  -> profiles/ncu_r1_summary.json   (ncu -i prof_r1.ncu-rep --page raw --csv)
    gpurun_out/bench_r1*.json   -> profiles/ (last line of each)

    python profiles/summarize.py
"""
import csv
import glob
import json
import os
import re
import shutil
from collections import OrderedDict

import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RND = sys.argv[1] if len(sys.argv) > 1 else "r2"      # python profiles/summarize.py [r1|r2]
OUT = os.path.join(ROOT, "gpurun_out")
HERE = os.path.join(ROOT, "profiles")


def short(name: str) -> str:
    m = re.search(r"(?:vdjg::)?(k_[a-z0-9_]+(?:<\d+>)?)", name)
    if m:
        return m.group(1)
    m = re.search(r"(DeviceRadixSort\w+|DeviceScan\w+)", name)
    return "cub::" + m.group(1) if m else name[:60]


def launches():
    src = os.path.join(OUT, f"launches_{RND}.csv")
    rows = [r for r in csv.reader(line for line in open(src) if line.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    for r in rows:
        n = short(r[k])
        c = agg.setdefault(n, [0, 0.0])
        c[0] += 1
        c[1] += float(r[v].replace(",", "")) / 1e6
    total = sum(c[1] for c in agg.values())
    shutil.copy(src, os.path.join(HERE, f"launches_{RND}.csv"))
    with open(os.path.join(HERE, f"launches_{RND}_summary.csv"), "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum) of `python bench.py --steps 2 --warmup 1 --no-cpu-baseline` (profiles/prof_" + RND + ".sh), c2 workload;\n")
        f.write("# k_pack (staging, one launch per chunk) filtered out.  Per-launch times under ncu are cold-cache and serialised: compare SHARES.\n")
        f.write("kernel,launches,total_ms,share\n")
        for n, (cnt, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{n},{cnt},{ms:.3f},{ms / total:.3f}\n")


METRICS = OrderedDict([
    ("time_ms", ("gpu__time_duration.sum", 1.0)),
    ("dram_read_GB", ("dram__bytes_read.sum", 1.0)),
    ("dram_write_GB", ("dram__bytes_write.sum", 1.0)),
    ("l2_sectors", ("lts__t_sectors.sum", 1.0)),
    ("l2_hit_pct", ("lts__t_sector_hit_rate.pct", 1.0)),
    ("l2_sectors_atom", ("lts__t_sectors_srcunit_tex_op_atom.sum", 1.0)),
    ("l2_sectors_red", ("lts__t_sectors_srcunit_tex_op_red.sum", 1.0)),
    ("l2_throughput_pct", ("lts__throughput.avg.pct_of_peak_sustained_elapsed", 1.0)),
    ("l1tex_throughput_pct", ("l1tex__throughput.avg.pct_of_peak_sustained_active", 1.0)),
    ("sm_throughput_pct", ("sm__throughput.avg.pct_of_peak_sustained_elapsed", 1.0)),
    ("achieved_occupancy_pct", ("sm__warps_active.avg.pct_of_peak_sustained_active", 1.0)),
    ("registers", ("launch__registers_per_thread", 1.0)),
    ("grid", ("launch__grid_size", 1.0)),
    ("warp_instructions", ("smsp__inst_executed.sum", 1.0)),
    ("stall_long_scoreboard_per_issue", ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", 1.0)),
    ("stall_barrier_per_issue", ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", 1.0)),
    ("issue_active_pct", ("smsp__issue_active.avg.pct_of_peak_sustained_active", 1.0)),
    # where the L2 misses come from: the streamed runs / reads (tex reads that miss) vs the table atomics
    ("l2_tex_read_sectors", ("lts__t_sectors_srcunit_tex_op_read.sum", 1.0)),
    ("l2_tex_read_hit_sectors", ("lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum", 1.0)),
    ("l2_tex_read_miss_sectors", ("lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum", 1.0)),
    ("l1_hit_pct", ("l1tex__t_sector_hit_rate.pct", 1.0)),
    ("lanes_per_instruction", ("smsp__thread_inst_executed_per_inst_executed.ratio", 1.0)),
])
UNIT = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3,                     # -> ms
        "byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}         # -> GB


def ncu_full():
    src = os.path.join(OUT, f"prof_{RND}_raw.csv")
    rows = list(csv.reader(line for line in open(src) if line.startswith('"')))
    hdr, units, rows = rows[0], rows[1], rows[2:]
    kn = hdr.index("Kernel Name")
    out = OrderedDict()
    for r in rows:
        d = OrderedDict()
        for key, (metric, _) in METRICS.items():
            cands = [i for i, h in enumerate(hdr) if h == metric or h.endswith("." + metric)]
            if not cands:
                continue
            val = None
            for i in cands:
                try:
                    val = float(r[i].replace(",", ""))
                    break
                except ValueError:
                    continue
            if val is None:
                continue
            if key == "time_ms" or key.endswith("_GB"):
                val *= UNIT.get(units[i], 1.0)
            d[key] = round(val, 4)
        if "dram_read_GB" in d:
            d["dram_traffic_bytes_per_launch"] = int(round((d["dram_read_GB"] + d.get("dram_write_GB", 0)) * 1e9, -5))
        out[short(r[kn])] = d
    json.dump({"source": "ncu --set full --clock-control none --import-source on (profiles/prof_" + RND + ".sh), c2 workload "
                         "(5M pairs 2x50, k=35), one launch each (cold caches, serialised)", "kernels": out},
              open(os.path.join(HERE, f"ncu_{RND}_summary.json"), "w"), indent=1)


def benches():
    for p in glob.glob(os.path.join(OUT, f"bench_{RND}*.json")):
        lines = [ln for ln in open(p).read().splitlines() if ln.startswith("{")]
        if lines:
            open(os.path.join(HERE, os.path.basename(p)), "w").write(lines[-1] + "\n")


if __name__ == "__main__":
    launches()
    ncu_full()
    benches()
    print(open(os.path.join(HERE, f"launches_{RND}_summary.csv")).read())
