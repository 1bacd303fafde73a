#!/bin/bash
# A/B timing of kernel variants built by tools/ab_build.py: bash tools/ab_run.sh <name> [<name> ...]  (run from the repo root on the GPU box)
# per variant: the 128-step rollout of the headline and the K = 1 launch, twice each
cd /root/repo
for v in "$@"; do
  for rep in 1 2; do
    for mode in "--steps 8 --warmup 3" "--fused 1 --steps 200 --warmup 20"; do
      QR_LIB_PATH=gpurun_ab/lib_$v.so timeout 150 python bench.py --no-cpu --no-extra $mode 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', '$mode', '%.4f G  ms/step %.4f kernel_ms %.4f frac %.4f'%(d['value']/1e9, d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac']), d.get('clocks'))"
    done
  done
done
