#!/bin/bash
# The workload matrix of DESIGN.md section 4 (MONO/MODUL/QUAD, f32/f64, policy, tracking, one step or 128 per launch): bash tools/bench_matrix.sh
cd /root/repo
mkdir -p gpurun_out
p() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', '%.4f G  ms/step %.4f frac %.4f'%(d['value']/1e9, d['ms_per_step'], d['roofline']['frac']), 'att', d['config'].get('mean_dop853_attempts'), 'eplen', d['config'].get('mean_episode_length'))"; }
b() { timeout 300 python bench.py --no-cpu --no-extra "$@" 2>/dev/null | tail -1; }
b --steps 8 --warmup 3 | p mono_f32_rollout128
b --fused 1 --steps 200 --warmup 20 | p mono_f32_k1
b --steps 8 --warmup 3 --framework MODUL --envs-per-gpu 1048576 | p modul_f32_rollout128
b --fused 1 --steps 200 --warmup 20 --framework MODUL --envs-per-gpu 1048576 | p modul_f32_k1
b --fused 1 --steps 60 --warmup 150 --dtype f64 --envs-per-gpu 1048576 | p mono_f64_k1
b --steps 8 --warmup 3 --policy --fused 64 --envs-per-gpu 1048576 | p policy_rollout64
b --fused 1 --steps 100 --warmup 10 --policy --envs-per-gpu 1048576 | p policy_two_kernels
for g in hover circle eight; do b --steps 6 --warmup 3 --goal $g | p tracking_$g; done
b --fused 1 --steps 200 --warmup 20 --actions zero | p zero_actions_k1
b --fused 1 --steps 100 --warmup 10 --framework QUAD | p quad_f32_k1
