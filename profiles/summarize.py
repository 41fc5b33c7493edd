"""Turns the raw ncu captures brought back in gpurun_out/ into the small, tracked summaries under profiles/.
   python profiles/summarize.py r01"""
import collections
import csv
import json
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
# ---- launch list (ncu --metrics gpu__time_duration.sum): per-kernel launches, total and share of the step ----
rows = [r for r in csv.reader(open(f"gpurun_out/launches_{tag}.csv")) if len(r) > 5 and r[0].isdigit()]
hdr = None
for r in csv.reader(open(f"gpurun_out/launches_{tag}.csv")):
    if r and r[0] == "ID":
        hdr = r
        break
ix = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows:
    name = r[ix["Kernel Name"]].split("(")[0]
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    us = v / 1e3 if unit in ("nsecond", "ns") else (v if unit in ("usecond", "us") else v * 1e3)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
with open(f"profiles/launches_{tag}_summary.csv", "w") as f:
    f.write("kernel,launches,total_us,avg_us,share\n")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"\"{k}\",{n},{us:.1f},{us / n:.2f},{us / tot:.4f}\n")
# ---- full capture: the metrics the roofline discussion uses ----
import glob
rr = None
for rep in sorted(glob.glob(f"gpurun_out/prof_{tag}*.ncu-rep")):   # one or several captures of the same round
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    part = list(csv.reader(raw.splitlines()))
    if rr is None:
        rr = part
    else:
        pix = {c: i for i, c in enumerate(part[0])}
        for r in part[2:]:
            rr.append([r[pix[c]] if c in pix else "" for c in rr[0]])
h, units = rr[0], rr[1]
ix = {c: i for i, c in enumerate(h)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
seen = collections.Counter()
out = []
for r in rr[2:]:
    name = r[ix["Kernel Name"]].split("(")[0]
    seen[name] += 1
    if seen[name] > 2:
        continue
    out.append({"kernel": name, **{w: (r[ix[w]] + " " + units[ix[w]]) for w in want if w in ix}})
json.dump(out, open(f"profiles/ncu_full_{tag}_summary.json", "w"), indent=1)
# ---- DRAM traffic per launch of the kernels bench.py reports (roofline.traffic) ----
def _num(v):
    val, unit = v.split()[0], (v.split() + [""])[1]
    return float(val.replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(unit, 1.0)
names = {"spmv_csr_kernel": "spmv_csr", "multi_dot_kernel": "multi_dot", "multi_axpy_dot_kernel": "multi_axpy", "blockdiag_apply_kernel": "blockdiag_apply",
         "dot_kernel": "dot", "thcm_assemble_kernel<0>": "thcm_assemble<RHS>"}
traffic = {}
for o in out:
    k = o["kernel"].replace("void ", "").replace("thcm::", "")
    key = next((v for n_, v in names.items() if k.startswith(n_)), None)
    b = _num(o["dram__bytes_read.sum"]) + _num(o["dram__bytes_write.sum"])
    if k.startswith("thcm_jac_tma_kernel"):   # the Jacobian is two launches (row groups): sum them
        e = traffic.setdefault("thcm_assemble<JAC_GRAPH>", {"dram_bytes": 0.0, "parts": [], "source": f"profiles/ncu_full_{tag}_summary.json"})
        if k not in e["parts"]:
            e["parts"].append(k); e["dram_bytes"] += b
    elif key and key not in traffic:
        traffic[key] = {"dram_bytes": b, "duration": o["gpu__time_duration.sum"], "source": f"profiles/ncu_full_{tag}_summary.json"}
json.dump(traffic, open("profiles/ncu_traffic.json", "w"), indent=1)
print(open(f"profiles/launches_{tag}_summary.csv").read())
for o in out:
    print(o["kernel"], o["gpu__time_duration.sum"], "read", o["dram__bytes_read.sum"], "write", o["dram__bytes_write.sum"])
