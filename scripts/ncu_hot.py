#!/usr/bin/env python3
"""Source-level digest of an `ncu --set full --import-source on` capture: headline counters, stall mix, and every SASS instruction
that carries >= 0.4 % of the executed warp instructions with its average active lanes.  usage: ncu_hot.py REPORT.ncu-rep OUT.txt"""
import csv, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
H, U, V = raw[0], raw[1], raw[2]
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
L = ["# " + rep, ""]
for w in want:
    if w in H: L.append("%-70s %s %s" % (w, V[H.index(w)], U[H.index(w)]))
src = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
HS = src[1]
ia, isr, ie, it = HS.index("Address"), HS.index("Source"), HS.index("Instructions Executed"), HS.index("Thread Instructions Executed")
stalls = [h for h in HS if h.startswith("stall_") and "Not Issued" not in h]
rows, tot, st = [], 0, {h: 0 for h in stalls}
for r in src[2:]:
    try:
        e, t = int(r[ie]), int(r[it])
    except Exception:
        continue
    rows.append((r[ia], r[isr], e, t)); tot += e
    for h in stalls:
        try: st[h] += int(r[HS.index(h)])
        except Exception: pass
S = sum(st.values()) or 1
L += ["", "stall samples: " + ", ".join("%s %.1f%%" % (h[6:], 100.0 * v / S) for h, v in sorted(st.items(), key=lambda kv: -kv[1]) if v * 200 > S), "",
      "instructions with >= 0.4 %% of %d executed warp instructions (share, average active lanes, SASS):" % tot]
base = int(rows[0][0], 16) if rows and rows[0][0].startswith("0x") else 0
for a, s, e, t in rows:
    if e * 250 >= tot:
        off = (int(a, 16) - base) if a.startswith("0x") else 0
        L.append("  +0x%04x %5.2f%% %5.1f  %s" % (off, 100.0 * e / tot, t / max(e, 1), s[:90]))
open(out, "w").write("\n".join(L) + "\n")
print("wrote", out, len(L), "lines")
