#!/usr/bin/env python
"""DRAM traffic of one captured launch -> profiles/r02/ncu_traffic.json (read by bench.py for roofline.traffic).
usage: ncu_traffic.py <capture.ncu-rep> <key> <envs per launch> [note] [digest file under profiles/]      key e.g. k_step_mono_f32_fused128"""
import csv, io, json, os, subprocess, sys

rep, key, envs = sys.argv[1], sys.argv[2], int(sys.argv[3])
note = sys.argv[4] if len(sys.argv) > 4 else ""
src = sys.argv[5] if len(sys.argv) > 5 else os.path.basename(rep)
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (float(v.replace(",", "")), u) for h, u, v in zip(hdr, units, vals) if v.replace(",", "").replace(".", "", 1).replace("-", "", 1).isdigit()}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
rd = d["dram__bytes_read.sum"][0] * scale[d["dram__bytes_read.sum"][1]]
wr = d["dram__bytes_write.sum"][0] * scale[d["dram__bytes_write.sum"][1]]
dur = d["gpu__time_duration.sum"]
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02", "ncu_traffic.json")
tj = json.load(open(path)) if os.path.exists(path) else {}
tj[key] = {"envs": envs, "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes": rd + wr,
           "duration": "%g %s" % dur, "source": src + " (ncu --set full)", "note": note}
json.dump(tj, open(path, "w"), indent=1, sort_keys=True)
print(key, tj[key])
