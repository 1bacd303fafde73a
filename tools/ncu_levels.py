import csv, io, subprocess, sys
rep=sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[1]
ix={n:h.index(n) for n in ("Source","# Samples","Instructions Executed","stall_no_inst","stall_wait","stall_long_sb","stall_short_sb","stall_branch_resolving","stall_not_selected","stall_selected","stall_math","stall_lg","stall_mio","stall_dispatch")}
data=[]
for r in rows[2:]:
    if r and r[0]=="Kernel Name": break
    try: data.append({k:(r[i].strip() if k=="Source" else int(r[i] or 0)) for k,i in ix.items()})
    except Exception as ex: pass
tot={k:sum(d[k] for d in data) for k in ix if k!="Source"}
print(tot)
# no_inst by position relative to previous branch: classify instruction i by whether previous instruction is a branch/BSYNC/label target
# print top no_inst instructions with context
top=sorted(range(len(data)), key=lambda i:-data[i]["stall_no_inst"])[:40]
cum=0
for i in top:
    d=data[i]; cum+=d["stall_no_inst"]
    prev=data[i-1]["Source"][:38] if i else ""
    print("%5d noinst %5d exec %8d  %-46s | prev: %s"%(i,d["stall_no_inst"],d["Instructions Executed"],d["Source"][:46],prev))
print("top40 share of no_inst: %.1f%%"%(100*cum/tot["stall_no_inst"]))
# distribution: fraction of no_inst on instructions whose exec differs from the previous one's (block entry) 
be=0
for i in range(1,len(data)):
    if data[i]["Instructions Executed"]!=data[i-1]["Instructions Executed"] or "BRA" in data[i-1]["Source"] or "BSYNC" in data[i-1]["Source"]:
        be+=data[i]["stall_no_inst"]
print("no_inst at block entries: %.1f%%"%(100*be/tot["stall_no_inst"]))
def lvl(x):
    if x>2_000_000: return "pair-loop"
    if x>600_000: return "per-stage"
    if x>150_000: return "0.15-0.6M"
    if x>60_000: return "per-env-step"
    if x>5_000: return "rare"
    return "cold"
import collections
agg=collections.OrderedDict()
keys=["# Samples","stall_no_inst","stall_wait","stall_long_sb","stall_short_sb","stall_selected","stall_not_selected","stall_branch_resolving","stall_math","stall_dispatch"]
for d in data:
    a=agg.setdefault(lvl(d["Instructions Executed"]),collections.Counter())
    for k in keys: a[k]+=d[k]
    a["dyn"]+=d["Instructions Executed"]; a["static"]+=1
print("%-14s %6s %9s "%("level","static","dyn%")+" ".join("%9s"%k.replace("stall_","")[:9] for k in keys))
for k,a in agg.items():
    print("%-14s %6d %8.1f%% "%(k,a["static"],100*a["dyn"]/tot["Instructions Executed"])+" ".join("%9d"%a[x] for x in keys))
