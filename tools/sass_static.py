#!/usr/bin/env python
"""Static SASS census of one kernel by (call-site line in the kernel body, inlined callee) from nvdisasm -gi inline chains.
usage: sass_static.py <cubin|lib.so> <mangled kernel substring> [top source file]   (needs -lineinfo)
Columns: instructions, of which packed FP (FFMA2/FMUL2/FADD2), scalar FP, shared ld/st, local ld/st, global ld/st, other."""
import collections, os, re, subprocess, sys, tempfile

obj, kern = sys.argv[1], sys.argv[2]
top_file = sys.argv[3] if len(sys.argv) > 3 else "qr_kernels.cuh"
if not obj.endswith(".cubin"):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cands = [os.path.join(tmp, f) for f in sorted(os.listdir(tmp)) if f.endswith(".cubin")]
else:
    cands = [obj]
for obj in cands:   # the library holds one cubin per translation unit: take the one that has the kernel
    dis = subprocess.run(["nvdisasm", "-gi", "-c", obj], capture_output=True, text=True).stdout.split("\n")
    hits = [i for i, l in enumerate(dis) if l.startswith(".text.") and kern in l and l.rstrip().endswith(":")]
    if hits:
        break
start = hits[0]
src_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gym_rotor_b200", "csrc")
funcs = {}
for f in os.listdir(src_dir):
    if f.endswith(".cuh"):
        st = []
        for n, l in enumerate(open(os.path.join(src_dir, f)), 1):
            m = re.search(r"(?:QR_DEV|__device__|__global__)[^;]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", l)
            if m and not l.lstrip().startswith("//"):
                st.append((n, m.group(1)))
        funcs[f] = st
def func_of(loc):
    f, n = loc
    st = funcs.get(f)
    if not st: return f
    best = f
    for a, name in st:
        if a <= n: best = name
    return best
def cls(op):
    b = op.split(".")[0]
    if b in ("FFMA2", "FMUL2", "FADD2"): return 1
    if b in ("FFMA", "FMUL", "FADD", "FSETP", "FMNMX", "FSEL", "MUFU", "FMNMX3", "DFMA", "DMUL", "DADD", "DSETP"): return 2
    if b in ("LDS", "STS", "LDSM"): return 3
    if b in ("LDL", "STL"): return 4
    if b in ("LDG", "STG", "LDGSTS", "ATOMG", "RED", "ATOM"): return 5
    return 6
agg = collections.defaultdict(lambda: [0] * 7)
focus = os.environ.get("FOCUS")     # callee name: print its opcode histogram too
fops = collections.Counter()
group, fresh = [], True
for l in dis[start + 1:]:
    if l.startswith("//--------------------- .text."): break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if fresh: group = []; fresh = False
        group.append((m.group(1).split("/")[-1], int(m.group(2))))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d\s+)?(\S+)", l)
    if m:
        fresh = True
        if not group: key = ("?", "?")
        else:
            outer = group[-1]
            # outermost inlined callee below the kernel body
            callee = func_of(group[-2]) if len(group) >= 2 else "-"
            key = ("%s:%d" % outer, callee)
        a = agg[key]; a[0] += 1; a[cls(m.group(2))] += 1
        if focus and key[1] == focus: fops[m.group(2).split(".")[0]] += 1
print("%-24s %-24s %6s %6s %6s %6s %6s %6s %6s" % ("call site", "inlined callee", "instr", "packed", "fp", "smem", "local", "global", "other"))
tot = [0] * 7
for (site, callee), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[: int(os.environ.get("TOP", "40"))]:
    print("%-24s %-24s %6d %6d %6d %6d %6d %6d %6d" % ((site, callee) + tuple(a)))
for a in agg.values():
    for i in range(7): tot[i] += a[i]
print("%-24s %-24s %6d %6d %6d %6d %6d %6d %6d" % (("TOTAL", "") + tuple(tot)))
if focus: print(focus, dict(fops.most_common(40)))
