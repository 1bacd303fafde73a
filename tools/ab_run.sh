#!/bin/bash
# A/B timing of kernel variants built by tools/ab_build.py: bash tools/ab_run.sh <name> [<name> ...]  (run from the repo root on the GPU box)
cd /root/repo
for v in "$@"; do
  for rep in 1 2; do
    QR_LIB_PATH=gpurun_ab/lib_$v.so timeout 120 python bench.py --steps 200 --warmup 20 --no-cpu 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', '%.4f G  ms/step %.4f kernel_ms %.4f'%(d['value']/1e9, d['ms_per_step'], d['roofline']['kernel_ms']), d.get('clocks'), d.get('gpu_launches'))"
  done
done
