#!/usr/bin/env python
"""Taken branches and instruction-supply stalls of one captured kernel (ncu --set full --import-source on, built with -lineinfo).
usage: ncu_branches.py <capture.ncu-rep> <lib.so> <mangled kernel substring> [THR]
Part 1: every branch the warp executes at least 0.1 times per round and takes in more than 15 % of its executions (taken
fraction estimated from the execution counts of the branch and of the instruction after it), with its source line.
Part 2: every instruction that holds at least THR % (default 0.08) of all samples as stall_no_instruction.
A "round" is one pass of the kernel's main loop: the most frequent execution count among the hot instructions."""
import csv, io, os, re, subprocess, sys, tempfile

rep, lib, kern = sys.argv[1], sys.argv[2], sys.argv[3]
thr = float(sys.argv[4]) if len(sys.argv) > 4 else 0.08
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
dis, start = None, None
for cubin in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):   # one cubin per translation unit
    d = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
    hits = [i for i, l in enumerate(d) if l.startswith(".text.") and kern in l and l.rstrip().endswith(":")]
    if hits:
        dis, start = d, hits[0]
        break
assert dis is not None, "kernel not found in " + lib
seq, group, fresh = [], [], True        # per instruction: the inline chain of (file, line), innermost first
for l in dis[start + 1:]:
    if l.startswith("//--------------------- .text."):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if fresh:
            group, fresh = [], False
        group.append((m.group(1).split("/")[-1].replace("qr_", "").replace(".cuh", ""), int(m.group(2))))
        continue
    if re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l):
        seq.append(list(group)); fresh = True
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out))); h = rows[1]
iS, iN, iX, iNI = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed"), h.index("stall_no_inst")
d = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break
    try:
        d.append((r[iS].strip(), int(r[iN]), int(r[iX]), int(r[iNI] or 0)))
    except Exception:
        pass
assert len(d) == len(seq), (len(d), len(seq))
tot = sum(x[1] for x in d)
import collections
mx = max(x[2] for x in d)
R = float(collections.Counter(x[2] for x in d if x[2] > 0.3 * mx).most_common(1)[0][0])   # the loop's own count: the most frequent one
where = lambda i: " < ".join("%s:%d" % g for g in seq[i][-3:])
print("== taken branches (rounds of the main loop: %.0f)" % R)
n_taken = 0.0
for i, x in enumerate(d[:-1]):
    s = x[0]
    if re.search(r"\bBRA\b", s) and x[2] > 0.1 * R:
        uncond = not s.startswith("@") and "BRA.U" not in s and "BRA.DIV" not in s
        frac = 1.0 if uncond else max(0.0, 1 - d[i + 1][2] / x[2])
        if frac > 0.15:
            n_taken += frac * x[2] / R
            print("%5d executed %4.2f/round, taken ~%4.2f  %-44s %s" % (i, x[2] / R, frac, s[:44], where(i)))
print("taken branches per round ~ %.1f" % n_taken)
print("== instructions holding >= %.2f %% of all samples as stall_no_instruction" % thr)
acc = 0
for i, x in enumerate(d):
    if 100.0 * x[3] / tot >= thr:
        acc += x[3]
        print("%5d x=%9d no_inst=%5.2f%% | prev %-34s | %-40s | %s" % (i, x[2], 100.0 * x[3] / tot, d[i - 1][0][:34], x[0][:40], where(i)))
print("listed: %.2f %% of the samples; all stall_no_instruction: %.2f %%" % (100.0 * acc / tot, 100.0 * sum(x[3] for x in d) / tot))
