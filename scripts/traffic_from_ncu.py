#!/usr/bin/env python3
"""profiles/k_match_traffic.json from an `ncu --set full` capture of the first k_match launch (gpurun_out/prof_k_match_TAG.ncu-rep):
dram bytes per launch + the hash of the kernel source the capture was made from (bench.py reports the number only for that build)."""
import csv, hashlib, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
H, U, V = rows[0], rows[1], rows[2]
def val(name):
    i = H.index(name)
    v = float(V[i].replace(",", "")); u = U[i].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
out = {"kernel": "k_match (first, full launch)", "dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
       "dram_read": val("dram__bytes_read.sum"), "dram_write": val("dram__bytes_write.sum"),
       "duration_under_ncu": V[H.index("gpu__time_duration.sum")] + " " + U[H.index("gpu__time_duration.sum")],
       "inst_executed": val("smsp__inst_executed.sum"), "threads_per_inst": val("smsp__thread_inst_executed_per_inst_executed.ratio"),
       "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
       "kernels_sha256": hashlib.sha256(open(os.path.join(ROOT, "zlib_rs_b200", "csrc", "zb_kernels.cu"), "rb").read()).hexdigest(),
       "capture": "ncu --set full --clock-control none -k regex:k_match -c 1 python scripts/one_deflate.py (" + os.path.basename(rep) + ")"}
json.dump(out, open(os.path.join(ROOT, "profiles", "k_match_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
