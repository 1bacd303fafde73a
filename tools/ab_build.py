#!/usr/bin/env python
"""Build kernel variants for A/B timing: python tools/ab_build.py name:DEF1=1,DEF2=0 ...  -> gpurun_ab/lib_<name>.so
Run one with QR_LIB_PATH=gpurun_ab/lib_<name>.so python bench.py ...
Extra nvcc flags for all variants of one invocation: QR_NVCC_EXTRA="-ftz=true -prec-div=false" python tools/ab_build.py ..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_rotor_b200 import build
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.makedirs(os.path.join(root, "gpurun_ab"), exist_ok=True)
for spec in sys.argv[1:]:
    name, _, defs = spec.partition(":")
    out = os.path.join(root, "gpurun_ab", "lib_%s.so" % name)
    build.build(force=True, out=out, defines=[d for d in defs.split(",") if d])
    print("built", out)
