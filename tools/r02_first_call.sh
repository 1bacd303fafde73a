#!/bin/bash
# First GPU call of the next round: A/B of the experiments that were implemented but never timed (profiles/r01_next_steps.md)
# plus the two micro-benchmarks.  Run from /root/repo IN THE BUILD CONTAINER; it builds everything here (nvcc cross-compiles)
# and prints the gpurun command (about 3 GPU-minutes).
set -e
cd /root/repo
python tools/ab_build.py base: pre:QR_TILE_PREFETCH=1 e3:QR_E3_FROM_B=1 ss:QR_STREAM_STORES=1 sl:QR_STREAM_LOADS=1 \
    pad:QR_OBS_PAD=1 w11:QR_STEP_THREADS_F32=352 all4:QR_TILE_PREFETCH=1,QR_E3_FROM_B=1,QR_STREAM_STORES=1,QR_STREAM_LOADS=1 all5:QR_OBS_PAD=1,QR_TILE_PREFETCH=1,QR_E3_FROM_B=1,QR_STREAM_STORES=1,QR_STREAM_LOADS=1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_ab/ubench_layout tools/ubench_layout.cu
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Igym_rotor_b200/csrc -o gpurun_ab/ubench_attempt tools/ubench_attempt.cu
cat <<'MSG'
built.  Now:
/usr/local/graft/bin/gpurun --timeout 600 -- 'timeout 120 python -m pytest tests/test_zz_late_gpu.py -m gpu -x -q 2>&1 | tail -3; bash tools/ab_run.sh base pre e3 ss sl pad w11 all4 all5; timeout 60 gpurun_ab/ubench_layout; timeout 60 gpurun_ab/ubench_attempt; for f in "--policy" "--policy --fused 64" "--policy --goal eight"; do timeout 60 python bench.py --steps 100 --warmup 10 --no-cpu $f 2>/dev/null | tail -1 | cut -c1-160; done'
MSG
