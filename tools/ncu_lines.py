#!/usr/bin/env python
"""Map an ncu source-page capture back to CUDA source lines (nvdisasm -g) for one kernel of a .so.
usage: ncu_lines.py <rep> <lib.so> <mangled kernel substring> [n]"""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, lib, kern = sys.argv[1], sys.argv[2], sys.argv[3]
n = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(dis) if l.startswith(".text.") and kern in l and l.rstrip().endswith(":")][0]
cur, seq, stack = None, [], None
for l in dis[start + 1:]:
    if l.startswith("//--------------------- .text."):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        cur = "%s:%s" % (m.group(1).split("/")[-1], m.group(2))
        if m.group(3):
            cur += " <- %s:%s" % (m.group(3).split("/")[-1], m.group(4))
        continue
    m2 = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m2:
        seq.append((cur, m2.group(2)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[1]
iS, iN, iX = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
data = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break
    try:
        data.append((r[iS].strip(), int(r[iN]), int(r[iX])))
    except Exception:
        pass
assert len(seq) == len(data), (len(seq), len(data))
ns = sum(d[1] for d in data)
print("-- hottest instructions with their source line")
for (loc, ins), (s, smp, x) in sorted(zip(seq, data), key=lambda t: -t[1][1])[:n]:
    print("%6d %5.2f%% exec %9d  %-44s %s" % (smp, 100.0 * smp / ns, x, s[:44], loc))
agg = collections.defaultdict(lambda: [0, 0, 0])
for (loc, ins), (s, smp, x) in zip(seq, data):
    key = loc.split(" <- ")[-1] if loc else "?"
    agg[key][0] += 1; agg[key][1] += x; agg[key][2] += smp
print("-- by outermost source line (inlined callers)")
for loc, (c, x, smp) in sorted(agg.items(), key=lambda kv: -kv[1][2])[:n]:
    print("%-34s static %5d  dyn %10d  samples %5.1f%%" % (loc, c, x, 100.0 * smp / ns))
