#!/usr/bin/env python
"""Stall samples of one ncu --set full capture, grouped by how often an instruction executes.

usage: ncu_levels.py <rep>
In the persistent step kernel the execution count of an instruction tells which loop level it belongs to
(stage-sum loop > per stage > per attempt > per env-step > rare paths > cold), so this shows where the
no_instruction / long_scoreboard / short_scoreboard / wait samples are, and the top instructions of one stall."""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[1]
cols = ("Source", "# Samples", "Instructions Executed", "stall_no_inst", "stall_wait", "stall_long_sb", "stall_short_sb",
        "stall_branch_resolving", "stall_not_selected", "stall_selected", "stall_math", "stall_lg", "stall_mio", "stall_dispatch")
ix = {n: h.index(n) for n in cols}
data = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break
    try:
        data.append({k: (r[i].strip() if k == "Source" else int(r[i] or 0)) for k, i in ix.items()})
    except Exception:
        pass
tot = {k: sum(d[k] for d in data) for k in cols if k != "Source"}
print("totals:", tot)


def level(x):
    if x > 2_000_000: return "stage-sum loop"
    if x > 600_000: return "per stage"
    if x > 150_000: return "0.15-0.6M"
    if x > 60_000: return "per env-step"
    if x > 5_000: return "rare"
    return "cold"


keys = ["# Samples", "stall_no_inst", "stall_wait", "stall_long_sb", "stall_short_sb", "stall_selected", "stall_not_selected",
        "stall_branch_resolving", "stall_math", "stall_dispatch"]
agg = collections.OrderedDict()
for d in data:
    a = agg.setdefault(level(d["Instructions Executed"]), collections.Counter())
    for k in keys:
        a[k] += d[k]
    a["dyn"] += d["Instructions Executed"]; a["static"] += 1
print("%-14s %6s %9s " % ("level", "static", "dyn%") + " ".join("%9s" % k.replace("stall_", "")[:9] for k in keys))
for k, a in agg.items():
    print("%-14s %6d %8.1f%% " % (k, a["static"], 100 * a["dyn"] / max(1, tot["Instructions Executed"])) + " ".join("%9d" % a[x] for x in keys))
if len(sys.argv) > 2:   # top instructions of one stall column, e.g. stall_long_sb
    col = sys.argv[2]
    print("-- top instructions by", col)
    for i in sorted(range(len(data)), key=lambda i: -data[i][col])[:20]:
        print("%5d %6d exec %8d  %-50s | prev: %s" % (i, data[i][col], data[i]["Instructions Executed"], data[i]["Source"][:50], data[i - 1]["Source"][:40]))
