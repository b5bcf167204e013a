"""Region digest of an ncu --set full --import-source on capture: headline counters, then runs of SASS instructions with similar execution counts
(share of executed warp instructions, share of stall samples, average active lanes, first opcodes).  usage: ncu_regions.py REPORT [threshold]"""
import csv, subprocess, sys
rep=sys.argv[1]; thr=float(sys.argv[2]) if len(sys.argv)>2 else 0.02
raw = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
H,V=raw[0],raw[2]
for w in ["gpu__time_duration.sum","smsp__inst_executed.sum","smsp__thread_inst_executed_per_inst_executed.ratio","smsp__issue_active.avg.pct_of_peak_sustained_active","sm__warps_active.avg.pct_of_peak_sustained_active","launch__grid_size","launch__block_size","sm__cycles_active.max","sm__cycles_active.avg"]:
    if w in H: print("%-60s %s"%(w,V[H.index(w)]))
src = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
HS = src[1]
ia, isr, ie, it = HS.index("Address"), HS.index("Source"), HS.index("Instructions Executed"), HS.index("Thread Instructions Executed")
stalls=[h for h in HS if h.startswith("stall_") and "Not Issued" not in h]
rows=[]
for r in src[2:]:
    try: rows.append((int(r[ia],16), r[isr], int(r[ie]), int(r[it]), sum(int(r[HS.index(h)] or 0) for h in stalls)))
    except: pass
base=rows[0][0]; tot=sum(r[2] for r in rows); stot=sum(r[4] for r in rows) or 1
out=[]; cur=None
for a,s,e,t,st in rows:
    op=s.split()[0] if not s.startswith('@') else s.split()[1]
    if cur and abs(cur['e']-e) <= 0.05*max(cur['e'],1):
        cur['n']+=1; cur['sum']+=e; cur['t']+=t; cur['st']+=st; cur['end']=a-base; cur['ops'].append(op)
    else:
        if cur: out.append(cur)
        cur={'start':a-base,'end':a-base,'e':e,'n':1,'sum':e,'t':t,'st':st,'ops':[op]}
out.append(cur)
for c in out:
    if c['sum']/tot>thr or c['st']/stot>thr:
        print("+0x%04x..+0x%04x n=%3d instr=%5.2f%% samples=%5.2f%% lanes=%4.1f  %s" % (c['start'],c['end'],c['n'],100*c['sum']/tot,100*c['st']/stot,c['t']/max(c['sum'],1),' '.join(c['ops'][:10])))
