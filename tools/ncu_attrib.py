#!/usr/bin/env python
"""Attribute an ncu source-page capture of one kernel to (call-site line in the kernel, inlined callee) using
nvdisasm -gi inline chains.  usage: ncu_attrib.py <rep> <lib.so> <mangled kernel substring> [kernel source file]
env: TOP=<rows>, STALLS=stall_no_inst,stall_wait,... (adds each site's share of those stall samples; column names of
`ncu --page source --csv`), SORT=<column index: 1 dynamic instructions, 2 samples, 3.. the STALLS columns>"""
import bisect, collections, csv, io, os, re, subprocess, sys, tempfile

rep, lib, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top_file = sys.argv[4] if len(sys.argv) > 4 else "qr_kernels.cuh"
src_dir = os.path.join(os.path.dirname(os.path.abspath(lib)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
for cubin in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):   # one cubin per translation unit
    dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
    hits = [i for i, l in enumerate(dis) if l.startswith(".text.") and kern in l and l.rstrip().endswith(":")]
    if hits:
        break
start = hits[0]

# function start lines per source file (crude: lines that look like a device function header)
funcs = {}
for f in os.listdir(src_dir):
    if f.endswith(".cuh"):
        starts = []
        for n, l in enumerate(open(os.path.join(src_dir, f)), 1):
            m = re.search(r"(?:QR_DEV|__device__|__global__)[^;]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", l)
            if m and not l.lstrip().startswith("//"):
                starts.append((n, m.group(1)))
        funcs[f] = starts
def func_of(loc):
    f, n = loc
    st = funcs.get(f)
    if not st: return f
    i = bisect.bisect_right([a for a, _ in st], n) - 1
    return st[i][1] if i >= 0 else f

seq, group, fresh = [], [], True
for l in dis[start + 1:]:
    if l.startswith("//--------------------- .text."): break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if fresh: group = []; fresh = False
        group.append((m.group(1).split("/")[-1], int(m.group(2))))
        continue
    if re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l):
        seq.append(list(group)); fresh = True
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out))); h = rows[1]
iN, iX = h.index("# Samples"), h.index("Instructions Executed")
STALLS = [c for c in os.environ.get("STALLS", "").split(",") if c]          # e.g. STALLS=stall_no_inst,stall_wait,stall_math
iS = [h.index(c) for c in STALLS]
data = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name": break
    try: data.append((int(r[iN]), int(r[iX])) + tuple(int(r[i] or 0) for i in iS))
    except Exception: pass
assert len(seq) == len(data), (len(seq), len(data))
ns, nx = sum(d[0] for d in data), sum(d[1] for d in data)
agg = collections.defaultdict(lambda: [0, 0, 0] + [0] * len(STALLS))
for g, (smp, x, *st) in zip(seq, data):
    if not g: key = ("?", "?")
    else:
        outer = g[-1]
        callee = func_of(g[-2]) if len(g) >= 2 else "-"
        key = ("%s:%d" % outer if outer[0] == top_file else "%s:%d" % outer, callee)
    a = agg[key]; a[0] += 1; a[1] += x; a[2] += smp
    for i, v in enumerate(st): a[3 + i] += v
tot = [sum(a[3 + i] for a in agg.values()) or 1 for i in range(len(STALLS))]
print("%-26s %-22s %6s %8s %8s" % ("call site", "inlined callee", "static", "dyn %", "samples %") + "".join(" %9s" % c.replace("stall_", "")[:9] for c in STALLS))
for (site, callee), (c, x, smp, *st) in sorted(agg.items(), key=lambda kv: -kv[1][int(os.environ.get('SORT', '2'))])[: int(os.environ.get("TOP", "45"))]:
    print("%-26s %-22s %6d %7.2f%% %7.2f%%" % (site, callee, c, 100.0 * x / nx, 100.0 * smp / ns) + "".join(" %8.2f%%" % (100.0 * v / t) for v, t in zip(st, tot)))
