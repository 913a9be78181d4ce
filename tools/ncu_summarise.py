"""tools/ncu_summarise.py -- turn the CSVs of tools/ncu_capture.sh into the summaries committed under profiles/:
    python tools/ncu_summarise.py <round tag, e.g. r02> <commit>
  gpurun_out/launches_iter.csv  -> profiles/<tag>_launch_shares.txt   (per-kernel launch counts, total time, shares)
  gpurun_out/traffic_bj.csv     -> profiles/<tag>_ncu_traffic.json    (DRAM bytes of one block-Jacobi apply)"""
import csv
import json
import os
import re
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, commit = sys.argv[1], sys.argv[2]


def rows(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    return list(csv.DictReader(lines))


def short(name):
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    return re.sub(r"\(.*$", "", name).strip()


def metric(rws, name):
    """{(ID, kernel): value} for one metric; values converted to base units"""
    out = OrderedDict()
    for r in rws:
        if r["Metric Name"] != name:
            continue
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        out[(r["ID"], short(r["Kernel Name"]))] = v * scale
    return out


it = os.path.join(ROOT, "gpurun_out", "launches_iter.csv")
if os.path.exists(it):
    dur = metric(rows(it), "gpu__time_duration.sum")
    agg = OrderedDict()
    for (_, k), us in dur.items():
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1; a[1] += us
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(ROOT, "profiles", tag + "_launch_shares.txt"), "w") as f:
        f.write("# ncu launch list at commit %s: `ncu --metrics gpu__time_duration.sum --clock-control none -k regex:... python tools/profile_apply.py 128 3 4`\n" % commit)
        f.write("# (tools/ncu_capture.sh step 1: Poisson 128^3, t=8, S=8 on one B200; the initial block-Jacobi apply + SpMM and 7 ECG iterations;\n")
        f.write("#  cold-cache, serialised: compare SHARES, not absolutes).  Full per-launch CSV: profiles/%s_launches_iter.csv\n" % tag)
        f.write("%-62s %8s %12s %7s\n" % ("kernel", "launches", "total_us", "share"))
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-62s %8d %12.1f %6.1f%%\n" % (k[:62], n, us, 100 * us / tot))
        f.write("%-62s %8d %12.1f\n" % ("TOTAL", sum(a[0] for a in agg.values()), tot))
        grp = {"block-Jacobi (sweep + assemble)": ("sweep", "assemble"), "SpMM": ("spmm",), "dense ECG passes": ("gram2", "ortho_update", "update_z", "reduce_partials", "fro2", "split_rhs")}
        f.write("\n")
        for g, pats in grp.items():
            us = sum(v[1] for k, v in agg.items() if any(p in k for p in pats))
            f.write("# %-34s %5.1f%% of the kernel time\n" % (g, 100 * us / tot))
    os.replace(it, os.path.join(ROOT, "profiles", tag + "_launches_iter.csv")) if False else None
    print(open(os.path.join(ROOT, "profiles", tag + "_launch_shares.txt")).read())

tr = os.path.join(ROOT, "gpurun_out", "traffic_bj.csv")
if os.path.exists(tr):
    rws = rows(tr)
    rd, wr, du = metric(rws, "dram__bytes_read.sum"), metric(rws, "dram__bytes_write.sum"), metric(rws, "gpu__time_duration.sum")
    # profile_apply.py 128 1 1 runs the apply three times (two warm-ups + one timed): keep the LAST third of the launches
    keys = list(du.keys())
    n = len(keys) // 3
    last = keys[-n:]
    out = {"what": "DRAM traffic of ONE block-Jacobi apply (%d launches: assemble_kernel, assemble_wide_kernel, sweep_kernel; Poisson 128^3, t=8, S=8, one B200)" % n,
           "command": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'sweep|assemble' python tools/profile_apply.py 128 1 1  (tools/ncu_capture.sh step 2)",
           "commit": commit, "launches": n,
           "dram_bytes_read": sum(rd[k] for k in last), "dram_bytes_write": sum(wr[k] for k in last),
           "sum_of_launch_durations_us": sum(du[k] for k in last)}
    out["traffic_bytes_per_apply"] = out["dram_bytes_read"] + out["dram_bytes_write"]
    json.dump(out, open(os.path.join(ROOT, "profiles", tag + "_ncu_traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))
