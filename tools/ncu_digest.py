#!/usr/bin/env python
"""Digest of an .ncu-rep: headline metrics, stall reasons, and the per-instruction hot spots (needs ncu on PATH)."""
import collections
import csv
import io
import subprocess
import sys


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__cycles_elapsed.max", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "sm__inst_issued.avg.per_cycle_active", "lts__t_bytes.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]


def main(rep, nhot=25):
    rows = page(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    for k in KEYS:
        if k in d:
            print("%-75s %s %s" % (k, d[k][0], d[k][1]))
    print("-- stall reasons per issued instruction")
    st = [(h.split("issue_stalled_")[1].split("_per_issue")[0], float(v)) for h, v in zip(hdr, vals)
          if "smsp__average_warps_issue_stalled_" in h and "_per_issue_active.ratio" in h and v]
    for n, v in sorted(st, key=lambda kv: -kv[1])[:8]:
        print("   %-24s %.3f" % (n, v))
    src = page(rep, "source")
    h = src[1]
    iS, iN, iX, iT = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
    data = []
    for r in src[2:]:
        if r and r[0] == "Kernel Name":
            break
        try:
            data.append((r[iS].strip(), int(r[iN]), int(r[iX]), int(r[iT])))
        except Exception:
            pass
    tot = sum(x[2] for x in data); samp = sum(x[1] for x in data)
    print("-- static SASS instructions %d, dynamic warp instructions %d, samples %d, avg active threads %.1f" % (
        len(data), tot, samp, sum(x[3] for x in data) / max(tot, 1)))
    op = collections.Counter(); ops = collections.Counter()
    for s, n, x, t in data:
        o = s.split()[1] if s.startswith("@") else s.split()[0]
        op[o.split(".")[0]] += x; ops[o.split(".")[0]] += n
    print("-- dynamic share / sample share by opcode")
    for o, c in op.most_common(16):
        print("   %-10s %5.1f %%   samples %5.1f %%" % (o, 100.0 * c / tot, 100.0 * ops[o] / max(samp, 1)))
    print("-- hottest instructions by samples")
    for s, n, x, t in sorted(data, key=lambda r: -r[1])[:nhot]:
        print("   %6d  %5.2f%%  exec %9d  %s" % (n, 100.0 * n / max(samp, 1), x, s[:90]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
