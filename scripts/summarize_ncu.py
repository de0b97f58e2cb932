#!/usr/bin/env python
"""Turn the ncu artefacts in gpurun_out/ into the tracked summaries under profiles/.
usage: python scripts/summarize_ncu.py <round tag, e.g. r1>"""
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
os.makedirs(PROF, exist_ok=True)

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "smsp__issue_active.avg.pct_of_peak_sustained_active",
]

# ---- launch list -----------------------------------------------------------------------------
path = os.path.join(OUT, "launches.csv")
if os.path.exists(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").strip()
        name = re.sub(r"<.*", "", name)
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(PROF, f"{tag}_launch_list_summary.md"), "w") as f:
        cmdf = os.path.join(OUT, "launches.cmd")
        cmd = open(cmdf).read().strip() if os.path.exists(cmdf) else "(see scripts/gpu_ncu.sh)"
        f.write(f"# {tag}: ncu launch list (gpu__time_duration.sum, --clock-control none)\n\n"
                f"Command: `{cmd}`\n\nOne whole conversion pass of the bench's default workload (config 2: "
                "whisper-small DiT + BigVGAN-22k, B=32, T=2580, 25 Euler steps), our kernels only (torch's "
                "one-time weight preparation filtered out).  Times under ncu are cold-cache and serialised: "
                "compare SHARES with `kernel_breakdown` / `roofline.share_of_step` of the bench line.\n\n"
                "| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {a[0]} | {a[1]:.3f} | {a[1] / tot:.3f} |\n")
        f.write(f"\nTotal {tot:.2f} ms over {sum(a[0] for a in agg.values())} launches.\n")
    print("wrote launch list summary")

# ---- DRAM traffic of every gemm_tc launch of one config-2 pass --------------------------------
path = os.path.join(OUT, "gemm_traffic.csv")
if os.path.exists(path):
    import json
    lines = [l for l in open(path) if not l.startswith("==")]
    tot = collections.defaultdict(float)
    ids = set()
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"].lower()
        scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-6, "nsecond": 1e-6,
                 "us": 1e-3, "usecond": 1e-3, "ms": 1, "msecond": 1}.get(u, 1)
        tot[row["Metric Name"]] += v * scale
        ids.add(row["ID"])
    n = len(ids)
    out = {"kernel": "gemm_tc_kernel", "launches": n,
           "command": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum "
                      "--clock-control none -k regex:gemm_tc_kernel -c 2142 python bench.py --steps 1 "
                      "--warmup 1 --no-cpu-baseline --no-profile (first conversion pass of config 2)",
           "dram_read_bytes": tot["dram__bytes_read.sum"], "dram_write_bytes": tot["dram__bytes_write.sum"],
           "traffic_bytes_per_launch": (tot["dram__bytes_read.sum"] + tot["dram__bytes_write.sum"]) / max(n, 1),
           "ncu_ms_total": tot["gpu__time_duration.sum"]}
    with open(os.path.join(PROF, f"{tag}_gemm_traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote gemm traffic", out["launches"], "launches")

# ---- full captures ---------------------------------------------------------------------------
for rep in sorted(os.listdir(OUT)):
    if rep.endswith("_raw.csv"):               # exported on the GPU box (scripts/gpu_ncu.sh: export_rep)
        text = open(os.path.join(OUT, rep)).read()
        rep = rep.replace("_raw.csv", ".ncu-rep")
    elif rep.endswith(".ncu-rep") and rep != "prof_k.ncu-rep":
        text = subprocess.run(["ncu", "-i", os.path.join(OUT, rep), "--page", "raw", "--csv"],
                              capture_output=True, text=True).stdout
    else:
        continue
    rows = list(csv.reader(text.splitlines()))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(os.path.join(PROF, f"{tag}_{rep.replace('.ncu-rep', '')}_full.md"), "w") as f:
        f.write(f"# {tag}: `ncu --set full --clock-control none --import-source on` - {rep}\n\n")
        for r in rows[2:]:
            f.write(f"## {r[idx['Kernel Name']][:140]}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in idx:
                    f.write(f"| {k} | {r[idx[k]]} | {units[idx[k]]} |\n")
            rd, wr = r[idx["dram__bytes_read.sum"]], r[idx["dram__bytes_write.sum"]]
            f.write(f"\ntraffic = dram read + write = {rd} + {wr} {units[idx['dram__bytes_read.sum']]}\n\n")
    print("wrote", rep)
