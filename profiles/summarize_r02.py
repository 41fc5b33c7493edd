"""Turns the ncu exports a GPU call brought back in gpurun_out/ into the tracked summaries under profiles/ (round 2).
The .ncu-rep files are too large to travel (64 MiB limit), so the GPU-side script exports `ncu --page raw --csv` per capture:
    python profiles/summarize_r02.py r02f
writes profiles/r02/ncu_full_<tag>_summary.json (selected metrics per captured launch), profiles/r02/launches_<tag>_summary.csv and
profiles/ncu_traffic.json (DRAM bytes per launch, keyed by kernel symbol -- what bench.py reports as roofline.traffic)."""
import collections
import csv
import json
import os
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r02f"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
os.makedirs(os.path.join(ROOT, "profiles", "r02"), exist_ok=True)

UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "second": 1e6, "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3,
        "s": 1e6, "ms": 1e3, "us": 1.0, "ns": 1e-3}
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def num(v, unit):
    return float(v.replace(",", "")) * UNIT.get(unit, 1.0)


# ---- launch list ----
lp = os.path.join(G, f"launches_{tag}.csv")
if os.path.exists(lp):
    rows = list(csv.reader(open(lp)))
    hdr = next(r for r in rows if r and r[0] == "ID")
    ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows:
        if len(r) <= 5 or not r[0].isdigit():
            continue
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("thcm::", "")
        us = num(r[ix["Metric Value"]], r[ix["Metric Unit"]])
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(ROOT, "profiles", "r02", f"launches_{tag}_summary.csv"), "w") as f:
        f.write("kernel,launches,total_us,avg_us,share\n")
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{k}\",{n},{us:.1f},{us / n:.2f},{us / tot:.4f}\n")

# ---- full captures ----
out, traffic = [], {}
for cap in ("asm", "krylov25", "krylov50"):
    p = os.path.join(G, f"ncu_raw_{cap}_{tag}.csv")
    if not os.path.exists(p):
        continue
    rows = list(csv.reader(open(p)))
    h, units = rows[0], rows[1]
    ix = {c: i for i, c in enumerate(h)}
    for r in rows[2:]:
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("thcm::", "")
        e = {"capture": cap, "kernel": name}
        for w in WANT:
            if w in ix:
                e[w] = f"{r[ix[w]]} {units[ix[w]]}".strip()
        rd, wr = num(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]]), num(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
        dur = num(r[ix["gpu__time_duration.sum"]], units[ix["gpu__time_duration.sum"]])
        e["dram_bytes"] = rd + wr
        e["duration_us"] = dur
        e["dram_gbs"] = (rd + wr) / (dur * 1e-6) / 1e9
        out.append(e)
        sym = name.split("<")[0]
        what = {"asm": "1-degree grid, one assembly", "krylov25": "1-degree grid, compact Krylov vectors, nv ~ 25 basis vectors (the average of a 50-iteration cycle)",
                "krylov50": "same, nv ~ 49"}[cap]
        if sym == "thcm_jac_tma_kernel":   # two launches (row groups) per assembly: sum one of each
            t = traffic.setdefault(sym, {"dram_bytes": 0.0, "duration_us": 0.0, "parts": [], "capture": f"profiles/r02/ncu_full_{tag}_summary.json: {what}"})
            if name not in t["parts"]:
                t["parts"].append(name); t["dram_bytes"] += rd + wr; t["duration_us"] += dur
        elif sym not in traffic:
            traffic[sym] = {"dram_bytes": rd + wr, "duration_us": dur, "capture": f"profiles/r02/ncu_full_{tag}_summary.json: {what}"}
json.dump(out, open(os.path.join(ROOT, "profiles", "r02", f"ncu_full_{tag}_summary.json"), "w"), indent=1)
json.dump(traffic, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
for e in out:
    print(f"{e['capture']:9s} {e['kernel'][:44]:44s} {e['duration_us']:9.1f} us  dram {e['dram_bytes'] / 1e6:9.1f} MB  {e['dram_gbs']:7.0f} GB/s  warps {e.get('sm__warps_active.avg.pct_of_peak_sustained_active', '')[:5]}  regs {e.get('launch__registers_per_thread', '')}")
