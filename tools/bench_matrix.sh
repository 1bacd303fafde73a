#!/bin/bash
# The workload matrix of DESIGN.md section 4 (MONO/MODUL/QUAD, f32/f64, policy, fused rollouts): bash tools/bench_matrix.sh
cd /root/repo
p() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', '%.4f G  ms/step %.4f'%(d['value']/1e9, d['ms_per_step']), 'e2e', d.get('e2e',{}).get('value'), 'att', d['config'].get('mean_dop853_attempts'), 'eplen', d['config'].get('mean_episode_length'))"; }
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu 2>/dev/null | tail -1 | tee gpurun_out/bench_mono_f32.json | p mono_f32
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu --framework MODUL --envs-per-gpu 1048576 2>/dev/null | tail -1 | tee gpurun_out/bench_modul_f32.json | p modul_f32_1M
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu --dtype f64 --envs-per-gpu 1048576 2>/dev/null | tail -1 | tee gpurun_out/bench_mono_f64.json | p mono_f64_1M
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu --policy --envs-per-gpu 1048576 2>/dev/null | tail -1 | tee gpurun_out/bench_config5.json | p config5_policy_1M
timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu --fused 64 2>/dev/null | tail -1 | tee gpurun_out/bench_fused64.json | p fused64
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu --actions zero 2>/dev/null | tail -1 | tee gpurun_out/bench_zero.json | p zero_actions
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu --framework QUAD 2>/dev/null | tail -1 | tee gpurun_out/bench_quad.json | p quad_f32
