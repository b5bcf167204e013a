#!/usr/bin/env python3
"""Turn gpurun_out ncu artefacts into small tracked summaries under profiles/.
usage: summarize_profile.py TAG [ncu-rep for the dominant kernel]"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
out = {"tag": tag}
lc = os.path.join(ROOT, "gpurun_out", "launches_%s.csv" % tag)
if os.path.exists(lc):
    rows = list(csv.reader(open(lc)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]; ki = H.index("Kernel Name"); vi = H.index("Metric Value")
    agg = {}
    for r in rows[hdr + 2:]:
        if len(r) <= vi: continue
        k = r[ki].split("(")[0].replace("zb::", "")
        agg.setdefault(k, []).append(float(r[vi].replace(",", "")) / 1e6)
    tot = sum(sum(v) for v in agg.values())
    out["launch_list"] = {"command": "ncu --metrics gpu__time_duration.sum --clock-control none python scripts/one_deflate.py "
                                     "(one level-6 deflate of silesia-small.tar, device resident)",
                          "total_ms": round(tot, 3),
                          "kernels": {k: {"launches": len(v), "total_ms": round(sum(v), 3), "share": round(sum(v) / tot, 4),
                                          "per_launch_ms": [round(x, 3) for x in v]} for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))}}
if len(sys.argv) > 2 and os.path.exists(sys.argv[2]):
    raw = subprocess.run(["ncu", "-i", sys.argv[2], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    H, U, V = rows[0], rows[1], rows[2]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "sm__cycles_active.avg", "sm__cycles_active.max", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
    m = {}
    for i, h in enumerate(H):
        if h in want: m[h] = {"value": V[i], "unit": U[i]}
    out["ncu_full"] = {"report": os.path.basename(sys.argv[2]), "kernel": H and rows[2][H.index("Kernel Name")] if "Kernel Name" in H else None, "metrics": m}
    def tobytes(e):
        v = float(e["value"].replace(",", "")); u = e["unit"].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    if "dram__bytes_read.sum" in m:
        out["ncu_full"]["dram_bytes_per_launch"] = tobytes(m["dram__bytes_read.sum"]) + tobytes(m["dram__bytes_write.sum"])
for f in ("bench_%s.json" % tag, "bench_ref_%s.json" % tag):
    p = os.path.join(ROOT, "gpurun_out", f)
    if os.path.exists(p):
        try: out[f.replace(".json", "")] = json.loads(open(p).read().strip().splitlines()[-1])
        except Exception: pass
dst = os.path.join(ROOT, "profiles", "%s_summary.json" % tag)
json.dump(out, open(dst, "w"), indent=1)
print("wrote", dst)
