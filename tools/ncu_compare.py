#!/usr/bin/env python
"""Side-by-side of two `ncu --set full` captures of the step kernel: headline metrics, stall reasons per issued
instruction and stall samples per loop level.  usage: ncu_compare.py <before.ncu-rep> <after.ncu-rep>"""
import collections, csv, io, subprocess, sys

METRICS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__inst_issued.avg.per_cycle_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
           "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "smsp__warps_eligible.avg.per_cycle_active"]
STALLS = ["stall_long_sb", "stall_short_sb", "stall_wait", "stall_no_inst", "stall_not_selected", "stall_selected", "stall_math",
          "stall_lg", "stall_mio", "stall_branch_resolving", "stall_dispatch"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        return {}
    h, units, vals = rows[0], rows[1], rows[2]
    return {k: (v, u) for k, u, v in zip(h, units, vals)}


def source(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h = rows[1]
    ix = {n: h.index(n) for n in ["# Samples", "Instructions Executed"] + STALLS}
    data = []
    for r in rows[2:]:
        if r and r[0] == "Kernel Name":
            break
        try:
            data.append({k: int(r[i] or 0) for k, i in ix.items()})
        except Exception:
            pass
    return data


def level(x):
    return ("stage-sum loop" if x > 2_000_000 else "per stage" if x > 600_000 else "0.15-0.6M" if x > 150_000
            else "per env-step" if x > 60_000 else "rare" if x > 5_000 else "cold")


def fnum(s):
    try:
        return float(s.replace(",", ""))
    except Exception:
        return None


def main():
    a, b = sys.argv[1], sys.argv[2]
    ra, rb = raw(a), raw(b)
    print("%-72s %16s %16s %8s" % ("metric", "before", "after", "ratio"))
    for m in METRICS:
        va, vb = fnum(ra.get(m, ("", ""))[0]), fnum(rb.get(m, ("", ""))[0])
        if va is None or vb is None:
            continue
        print("%-72s %16.4g %16.4g %8.3f" % (m + " [" + ra[m][1] + "]", va, vb, vb / va if va else float("nan")))
    sa, sb = source(a), source(b)
    ia, ib = sum(d["Instructions Executed"] for d in sa), sum(d["Instructions Executed"] for d in sb)
    print("\nstalls per issued warp instruction (samples / selected samples)")
    for st in STALLS:
        xa = sum(d[st] for d in sa) / max(1, sum(d["stall_selected"] for d in sa))
        xb = sum(d[st] for d in sb) / max(1, sum(d["stall_selected"] for d in sb))
        print("  %-26s %8.3f %8.3f %+8.3f" % (st, xa, xb, xb - xa))
    print("\nper loop level: share of dynamic instructions / share of samples")
    for name, data, tot in (("before", sa, ia), ("after", sb, ib)):
        agg = collections.OrderedDict()
        ns = sum(d["# Samples"] for d in data)
        for d in data:
            g = agg.setdefault(level(d["Instructions Executed"]), [0, 0, 0])
            g[0] += 1; g[1] += d["Instructions Executed"]; g[2] += d["# Samples"]
        print("  " + name + ": " + "; ".join("%s %d static %.1f%%/%.1f%%" % (k, v[0], 100 * v[1] / max(1, tot), 100 * v[2] / max(1, ns)) for k, v in agg.items()))
    print("\ndynamic warp instructions: %d -> %d (%.3f)" % (ia, ib, ib / max(1, ia)))


if __name__ == "__main__":
    main()
