#!/usr/bin/env python
"""env-steps/sec of the batched quadrotor step on N B200s (BASELINE.json metric), one JSON line on rank 0.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...        # the reference's CPU path (scipy port) on the host cores

Headline workload (the configuration the 1e10 steps/s target is quoted on): CoupledWrapper,
float32 state arithmetic, 2^21 envs per GPU (= 2^24 over 8 GPUs, weak scaling), on-device trajectory-generator goals
(mode 0), in-kernel auto reset with the trainer's 4000-step limit.  One bench "step" = one ROLLOUT of 128 env.step()
calls per env in one launch through the C ABI (qr_rollout: in-kernel Philox U(-1,1) actions, state resident in
registers) followed by the sum all-reduce of the episode statistics (NCCL) -- so every timed step contains the
path's one collective, resets occur at their true rate, and 20 steps are 2560 env-steps per env.

Also in the line: `k1` (one env.step() per launch, actions read from HBM, all per-step outputs written -- the figure
of round 1), `tracking` (the same rollout with figure-eight goals generated in the kernel), `e2e` (qr_step_host with
pinned host buffers) with a bare-copy ceiling -- all three at every N --, `configs` (BASELINE.json configs 3 and 5,
float64 mode; N=1 only), the CPU legs (reference cost structure on all host cores).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

ALG_BYTES = {"MONO_f32": 366, "MONO_f64": 614, "MODUL_f32": 363, "MODUL_f64": 611, "QUAD_f32": 366, "QUAD_f64": 614}   # SURVEY 8(d), per env-step at K = 1
ALG_FLOPS = 5450            # one env-step with one DOP853 attempt (the figure roofline.achieved is computed from)
ALG_FLOPS_EXTRA = 4780      # each further attempt
STATS_EVERY_K1 = 128        # K = 1 launches: statistics all-reduce every 128 steps
FP32_LANES_PER_SM = 128     # public B200 figure; SMs and clocks come from the device query
TRAFFIC_JSON = os.path.join(ROOT, "profiles", "r02", "ncu_traffic.json")   # written by tools/ncu_traffic.py from the .ncu-rep captures


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), d.get("sm_max_mhz", 1965.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1965.0, "fallback (B200_PROFILING.md)"


def _cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for l in f:
                if l.startswith("model name"):
                    return l.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


class ClockSampler(threading.Thread):
    """Samples nvidia-smi SM clocks and throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()

    def run(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.samples.append(float(f[0])); self.max_mhz = float(f[1])
                for nme, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nme)
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------------------
# CPU legs (the only place bench.py executes oracle/)
# ------------------------------------------------------------------------------------------------------

def _port_worker(args):
    framework, seconds, seed = args
    import numpy as np
    import quad_oracle as qo
    env = qo.ScipyPort(framework)
    rng = np.random.default_rng(seed)
    env.reset("train", rng)
    A = 4 if framework == "MONO" else 5
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        obs, rew, done, _, _ = env.step(rng.uniform(-1, 1, A))
        n += 1
        if any(done):
            env.reset("train", rng)
    return n, time.perf_counter() - t0


def cpu_port_baseline(framework="MONO", seconds=15.0, procs=None):
    """The reference's cost structure (Python RHS handed to scipy DOP853), one env per host core."""
    import multiprocessing as mp
    procs = procs or os.cpu_count() or 1
    import quad_oracle as qo
    qo.build()
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(procs) as pool:
        res = pool.map(_port_worker, [(framework, seconds, 1000 + i) for i in range(procs)])
    wall = time.perf_counter() - t0
    steps = sum(r[0] for r in res)
    return {"value": steps / max(r[1] for r in res), "unit": "env-steps/s", "cores": procs, "kind": "port", "cpu_model": _cpu_model(),
            "sample": "%d procs x %.0f s of %s ScipyPort.step (numpy RHS + scipy DOP853), U(-1,1) actions, train resets; "
                      "%d steps, wall %.1f s" % (procs, seconds, framework, steps, wall)}


def cpu_quad_v0_config1(steps=1000):
    """BASELINE.json config 1: Quad-v0, one env, 1000 steps of U(-1,1)^4 rotor-thrust actions, one core (quad.py:142-168, 225-242)."""
    import numpy as np
    import quad_oracle as qo
    env = qo.ScipyPort("QUAD")
    rng = np.random.default_rng(0)
    env.reset("train", rng)
    t0 = time.perf_counter()
    n = 0
    for _ in range(steps):
        obs, rew, done, _, _ = env.step(rng.uniform(-1, 1, 4))
        n += 1
        if any(np.atleast_1d(done)):
            env.reset("train", rng)
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "env-steps/s", "cores": 1, "kind": "port", "cpu_model": _cpu_model(),
            "sample": "Quad-v0 ScipyPort, 1 env x %d steps, U(-1,1)^4 (T1..T4) actions, train resets, %.2f s" % (n, dt)}


def cpu_c_baseline(framework="MONO", n=1 << 16, reps=3):
    """The plain-C restatement on all host threads (a far stronger CPU baseline than the reference's Python)."""
    import numpy as np
    import quad_oracle as qo
    thr = qo.lib().qo_get_max_threads()
    orc = qo.COracle(framework, threads=thr)
    rng = np.random.default_rng(0)
    st, ig, par = orc.reset_from_uniforms(rng.random((n, 20)))
    goal = np.zeros((n, 12)); goal[:, 6] = 1.0
    best = 0.0
    for _ in range(reps):
        a = rng.uniform(-1, 1, (n, orc.act_dim))
        t0 = time.perf_counter()
        orc.step(st, ig, par, goal, a)
        best = max(best, n / (time.perf_counter() - t0))
    return {"value": best, "unit": "env-steps/s", "cores": thr, "kind": "port", "cpu_model": _cpu_model(),
            "sample": "C oracle (gcc -O2, float64), %d envs x 1 step, best of %d, %d pthreads" % (n, reps, thr)}


# ------------------------------------------------------------------------------------------------------

def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fw = args.framework if args.framework != "QUAD" else "MONO"
    seconds = max(20.0, args.cpu_seconds)   # a fixed wall time of at least 20 s
    base = cpu_port_baseline(fw, seconds=seconds)
    line = {"impl": "reference", "metric": "env-steps/sec", "value": base["value"], "unit": "env-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / base["value"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            # same workload as the CUDA arm (CoupledWrapper env.step under U(-1,1) actions with the trainer's reset
            # protocol and mode-0 goals); the envs are stepped one per host core instead of 2^21 per GPU
            "config": {"workload": "%s env.step, U(-1,1) actions, reset on termination: the reference's CPU "
                                   "path (numpy RHS + scipy solve_ivp DOP853), one env per host core, bounded sample of %.0f s "
                                   "(fixed goal: the trajectory generator's per-step goal update is not in the sample)" % (
                                       {"MONO": "CoupledWrapper", "MODUL": "DecoupledWrapper"}[fw], seconds),
                       "framework": fw, "actions": "random", "envs_per_gpu": args.envs_per_gpu, "host_cores": base["cores"],
                       "cpu_model": base["cpu_model"]},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def _make_env(vec_env, n, fw, dtype, dev, seed, goal, policy, offset, limit=None):
    env = vec_env.BatchedQuadEnv(n, framework=fw, dtype=dtype, device=dev, seed=seed, autoreset=True,
                                 goal_mode=(goal if fw != "QUAD" else "external"),
                                 env_type="eval" if policy else "train",
                                 max_episode_steps=limit if limit is not None else (1000 if policy else 4000),
                                 env_id_offset=offset, diagnostics=False)
    env.reset(env_type="eval" if policy else "train")
    if fw != "QUAD":
        env.init_goal()
    env.get_norm_error_state()
    return env


def _timed(torch, dev, fn, steps, warmup):
    """fn(i) `steps` times after `warmup`, CUDA events on the current stream; returns ms per call."""
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / steps


def _attempts(stats):
    """Mean DOP853 attempts per env-step from the RHS-evaluation count as scipy keeps it (nfev = 2 + 12 per attempt); the
    attempt histogram (statistics 10..13) is only accumulated by handles created with diagnostics on."""
    return float((stats[9] / max(1.0, stats[7]) - 2.0) / 12.0)


def sub_configs(torch, vec_env, dev, seed, quick):
    """BASELINE.json configs 3 and 5, the float64 mode and trajectory tracking, each on its own handle (N = 1 runs only)."""
    import numpy as np
    out = {}
    s = 0.25 if quick else 1.0

    def steps(k):
        return max(3, int(k * s))
    # config 3: DecoupledWrapper, 2^20 envs, float32, Philox actions drawn on the device every step, auto reset
    n = 1 << 20
    env = _make_env(vec_env, n, "MODUL", torch.float32, dev, seed, "traj0", False, 0)
    ms = _timed(torch, dev, lambda i: env.rollout(1), steps(200), steps(20))
    st = env.stats()
    out["config3_modul_f32_2^20"] = {"value": n / (ms * 1e-3), "unit": "env-steps/s", "ms_per_step": ms, "steps": steps(200),
                                     "workload": "DecoupledWrapper env.step (o1/o2, two rewards), 2^20 envs, one step per launch, Philox U(-1,1)^5 actions "
                                                 "drawn in-kernel every step, auto reset, mode-0 goals", "mean_dop853_attempts": _attempts(st)}
    ms = _timed(torch, dev, lambda i: env.rollout(128), steps(8), 2)
    out["config3_modul_f32_2^20"]["rollout128_value"] = n * 128 / (ms * 1e-3)
    env.close()
    # config 5: the shipped TD3 actor in the loop, 2^20 envs, eval resets, 1000-step episodes: mean return vs KAT-2 (989.3 MONO)
    env = _make_env(vec_env, n, "MONO", torch.float32, dev, seed, "traj0", True, 0)
    env.stats()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nl = 4 if quick else 16
    e0.record()
    for i in range(nl):
        env.rollout(64, actions="policy")   # obs -> actor -> env.step, 64 times, one launch
    e1.record()
    torch.cuda.synchronize(dev)
    st = env.stats()
    msf = e0.elapsed_time(e1) / nl
    c5 = {"value": n * 64 / (msf * 1e-3), "unit": "env-steps/s", "ms_per_launch": msf, "launches": nl,
          "workload": "CoupledWrapper, 2^20 envs, the reference's shipped TD3 actor evaluated in the step kernel (qr_rollout, QR_ACT_POLICY), "
                      "64 env.step per launch, eval resets, 1000-step episodes",
          "episodes": st[0], "mean_return": float(st[1] / max(1.0, st[0])) if st[0] else None, "crashed": st[4],
          "kat2_reference_return": 989.3, "mean_dop853_attempts": _attempts(st)}
    act = torch.empty((n, env.act_dim), dtype=torch.float32, device=dev)
    ms2 = _timed(torch, dev, lambda i: env.step(env.policy_td3(out=act)), steps(100), steps(10))
    c5["two_kernel_value"] = n / (ms2 * 1e-3)
    out["config5_td3_mono_2^20"] = c5
    env.close()
    # float64 mode (the parity mode), CoupledWrapper 2^20
    env = _make_env(vec_env, n, "MONO", torch.float64, dev, seed, "traj0", False, 0)
    pool = [torch.rand((n, 4), device=dev, dtype=torch.float32) * 2 - 1 for _ in range(8)]
    ms = _timed(torch, dev, lambda i: env.step(pool[i % 8]), steps(40), steps(8))
    st = env.stats()
    out["mono_f64_2^20"] = {"value": n / (ms * 1e-3), "unit": "env-steps/s", "ms_per_step": ms,
                            "workload": "CoupledWrapper env.step, float64 arithmetic and storage, 2^20 envs, one step per launch, actions from HBM",
                            "mean_dop853_attempts": _attempts(st),
                            "fp64_issue_frac": ALG_FLOPS * n / (ms * 1e-3) / (torch.cuda.get_device_properties(dev).multi_processor_count * 64 * 2 * _peaks()[1] * 1e6)}
    env.close()
    del pool
    return out


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from gym_rotor_b200 import vec_env
    from gym_rotor_b200.dist import allreduce_stats

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the simulator has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    fw = args.framework
    dtype = torch.float64 if args.dtype == "f64" else torch.float32
    n = args.envs_per_gpu
    fused = max(1, args.fused)
    if fw != "QUAD" and args.goal in ("hover", "circle", "eight") and fused > 1:
        pass   # trajectory modes are evaluated inside the step kernel
    env = _make_env(vec_env, n, fw, dtype, dev, args.seed, args.goal, args.policy, rank * n)
    A = env.act_dim
    gen = torch.Generator(device=dev); gen.manual_seed(1234 + rank)

    def action_pool(count):
        if args.actions == "zero":   # hover-like thrust, no torque: the single-attempt, no-reset regime of a trained policy
            pool = [torch.zeros((n, A), device=dev, dtype=torch.float32) for _ in range(2)]
            for p in pool:
                p[:, 0] = -0.06
            return pool
        return [torch.rand((n, A), device=dev, dtype=torch.float32, generator=gen) * 2 - 1 for _ in range(count)]

    pool = action_pool(128) if (fused == 1 and not args.policy) else None
    actors = torch.empty((n, A), dtype=torch.float32, device=dev) if args.policy else None
    stats_total = np.zeros(len(env.stats(reset=False)))

    def read_stats():
        return allreduce_stats(env, dev) if world > 1 else env.stats()

    def launch(i):
        if args.policy and fused > 1:
            env.rollout(fused, actions="policy")   # obs -> shipped actor -> env.step, `fused` times, in ONE launch
        elif args.policy:
            env.step(env.policy_td3(out=actors))   # compiled actor kernel + step kernel: two launches per env.step
        elif fused > 1:
            env.rollout(fused)      # `fused` env.step() calls in one launch, Philox actions drawn in-kernel
        else:
            env.step(pool[i % len(pool)])

    def one_step(i, ev=None):
        if ev is not None:
            ev[0].record()
        launch(i)
        if ev is not None:
            ev[1].record()
        if fused > 1 or (i + 1) % STATS_EVERY_K1 == 0:
            return read_stats()   # one all-reduce per rollout (the path's only collective)
        return None

    sampler = ClockSampler(local); sampler.start()   # samples under load: warm-up + timed region
    for i in range(args.warmup):
        one_step(i)
    read_stats()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    launches0 = env.launch_count()
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    n_allreduce = 0
    for i in range(args.steps):
        s = one_step(i, kev[i])
        if s is not None:
            stats_total += s; n_allreduce += 1
    ev1.record()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    ms = ev0.elapsed_time(ev1)
    kernel_ms = sum(a.elapsed_time(b) for a, b in kev) / args.steps   # the launch(es) of one step alone, same stream
    launches = env.launch_count() - launches0
    clocks = sampler.stop()
    stats_total += read_stats()
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    total_steps = float(n) * world * args.steps * fused
    value = total_steps / (ms * 1e-3)

    # ---- K = 1: one env.step() per launch, actions from HBM, every per-step output written (round 1's headline) ----
    k1 = None
    if fused > 1 and not args.policy and not args.no_extra:
        kpool = action_pool(128)
        k1_steps = 100
        for i in range(10):
            env.step(kpool[i % len(kpool)])
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for i in range(k1_steps):
            env.step(kpool[(10 + i) % len(kpool)])
        a1.record()
        torch.cuda.synchronize(dev)
        tk = torch.tensor([a0.elapsed_time(a1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tk, op=dist.ReduceOp.MAX)
        sk = read_stats()
        k1_ms = float(tk.item()) / k1_steps
        k1 = {"value": float(n) * world / (k1_ms * 1e-3), "unit": "env-steps/s", "ms_per_step": k1_ms, "steps": k1_steps,
              "workload": "one env.step per launch (qr_step), 128 distinct U(-1,1) action tensors resident in HBM, obs/reward/done/state written every step, steady state",
              "mean_dop853_attempts": _attempts(sk), "mean_episode_length": float(sk[3] / max(1.0, sk[0])),
              "hbm_gbs": ALG_BYTES.get("%s_%s" % (fw, args.dtype), 366) * n / (k1_ms * 1e-3) / 1e9,
              "fp_issue_frac": None}
        del kpool

    # ---- trajectory tracking (north_star: "random-action and trajectory-tracking workloads at 1, 2, 4 and 8 GPUs"): the same
    # rollout with figure-eight goals generated in the step kernel before every step (trajectory_generator mode 6)
    tracking = None
    if fused > 1 and not args.policy and not args.no_extra and fw != "QUAD" and args.goal == "traj0":
        env_t = _make_env(vec_env, n, fw, dtype, dev, args.seed, "eight", False, rank * n)
        for i in range(2):
            env_t.rollout(fused)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        t_steps = 4
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        for i in range(t_steps):
            env_t.rollout(fused)
        b1.record()
        torch.cuda.synchronize(dev)
        tt = torch.tensor([b0.elapsed_time(b1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        st_t = allreduce_stats(env_t, dev) if world > 1 else env_t.stats()
        t_ms = float(tt.item()) / t_steps
        tracking = {"value": float(n) * world * fused / (t_ms * 1e-3), "unit": "env-steps/s", "ms_per_launch": t_ms, "launches": t_steps,
                    "workload": "same rollout (%d env.step per launch, Philox actions, auto reset) with figure-eight goals generated in the step "
                                "kernel before every step (trajectory_generator mode 6, main.py:145-147)" % fused,
                    "mean_dop853_attempts": _attempts(st_t), "mean_episode_length": float(st_t[3] / max(1.0, st_t[0]))}
        env_t.close()

    # ---- end to end through the host-buffer entry point (qr_step_host): pinned host actions in, obs/reward/done out
    e2e_steps = max(3, min(args.steps, 20))
    act_h = [torch.empty((n, A), dtype=torch.float32).uniform_(-1, 1).pin_memory() for _ in range(2)]
    obs_h = torch.empty((n, env.obs_dim), dtype=torch.float32).pin_memory()
    rew_h = torch.empty((n, env.n_agents), dtype=dtype).pin_memory()
    done_h = torch.empty((n, env.n_agents), dtype=torch.uint8).pin_memory()
    for i in range(2):
        env.step_host(act_h[i % 2], obs_h, rew_h, done_h)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        env.step_host(act_h[i % 2], obs_h, rew_h, done_h)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = float(n) * world * e2e_steps / float(te.item())
    h2d = n * A * 4
    d2h = n * (env.obs_dim * 4 + env.n_agents * (8 if dtype == torch.float64 else 4) + env.n_agents)
    # bare-copy ceiling: the same bytes per step over the same two directions, no kernel (all ranks at once)
    d_act = torch.empty((n, A), dtype=torch.float32, device=dev)
    d_obs = torch.empty((n, env.obs_dim), dtype=torch.float32, device=dev)
    d_rew = torch.empty((n, env.n_agents), dtype=dtype, device=dev)
    d_done = torch.empty((n, env.n_agents), dtype=torch.uint8, device=dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def copies():
        with torch.cuda.stream(s_in):
            d_act.copy_(act_h[0], non_blocking=True)
        with torch.cuda.stream(s_out):
            obs_h.copy_(d_obs, non_blocking=True); rew_h.copy_(d_rew, non_blocking=True); done_h.copy_(d_done, non_blocking=True)
    copies(); torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        copies()
    torch.cuda.synchronize(dev)
    tc = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tc, op=dist.ReduceOp.MAX)
    copy_ceiling = float(n) * world * e2e_steps / float(tc.item())

    if rank == 0:
        hbm_peak, sm_max, which = _peaks()
        props = torch.cuda.get_device_properties(dev)
        sms = props.multi_processor_count
        if clocks.get("sm_max_mhz"):
            sm_max = clocks["sm_max_mhz"]
        key = "%s_%s" % (fw, args.dtype)
        bytes_per = ALG_BYTES.get(key, 366)
        launches_per_step = 2 if (args.policy and fused == 1) else 1
        env_steps_per_launch = n * fused
        per_gpu_steps = env_steps_per_launch / (kernel_ms * 1e-3)
        lanes = FP32_LANES_PER_SM * (0.5 if args.dtype == "f64" else 1.0)
        fp_peak = sms * lanes * 2 * sm_max * 1e6 / 1e12
        mean_att = _attempts(stats_total)
        ach_tf = ALG_FLOPS * per_gpu_steps / 1e12
        ach_tf_att = (ALG_FLOPS + ALG_FLOPS_EXTRA * (mean_att - 1.0)) * per_gpu_steps / 1e12
        # algorithmic HBM bytes of one launch: every env is read and written once per launch (K fused sub-steps add nothing
        # without per-sub-step outputs; the in-kernel Philox actions are not read from memory)
        alg_bytes_launch = bytes_per * n - (16 * n if (fused > 1 and not args.policy) else 0)
        traffic = None
        tkey = "k_step_%s_%s_%s" % (fw.lower(), args.dtype, "fused%d" % fused if fused > 1 else "k1")
        if os.path.exists(TRAFFIC_JSON):
            with open(TRAFFIC_JSON) as f:
                tj = json.load(f)
            if tkey in tj and tj[tkey].get("envs") == n:
                traffic = tj[tkey]
        goal_txt = {"traj0": "on-device trajectory-generator mode-0 goals", "external": "external goals"}.get(
            args.goal if fw != "QUAD" else "external", "on-device trajectory-generator '%s' goals (trajectory tracking)" % args.goal)
        if args.policy:
            act_txt = "the reference's shipped TD3 actor in the loop (%s)" % ("evaluated inside the step kernel" if fused > 1 else "qr_policy_td3 kernel + qr_step")
        elif fused > 1:
            act_txt = "Philox U(-1,1) actions drawn in-kernel"
        else:
            act_txt = "%s actions read from HBM (%d distinct tensors)" % ("U(-1,1)" if args.actions == "random" else "constant hover-like", len(pool))
        workload = "%s env.step, %d envs/GPU x %d GPU(s) = %d envs, float%s, %s, %s, in-kernel auto reset (%s resets, %d-step limit), %s" % (
            {"MONO": "CoupledWrapper", "MODUL": "DecoupledWrapper", "QUAD": "Quad-v0"}[fw], n, world, n * world, args.dtype[1:], goal_txt, act_txt,
            "eval" if args.policy else "train", 1000 if args.policy else 4000,
            ("one bench step = one rollout of %d env.step per env in one launch (qr_rollout) + one statistics all-reduce" % fused) if fused > 1
            else "one bench step = one launch (qr_step), statistics all-reduce every %d steps" % STATS_EVERY_K1)
        line = {
            "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": workload, "envs_per_gpu": n, "framework": fw, "goal": args.goal if fw != "QUAD" else "external",
                       "fused_steps_per_launch": fused, "env_steps_per_bench_step": n * world * fused,
                       "stats_allreduces_in_timed_region": n_allreduce,
                       "l2": "per-launch working set %.0f MB > 126 MB L2 (inputs larger than L2)" % (bytes_per * n / 1e6),
                       "mean_dop853_attempts": mean_att, "mean_rhs_evaluations": float(stats_total[9] / max(1.0, stats_total[7])),
                       "episodes": stats_total[0], "mean_episode_length": float(stats_total[3] / max(1.0, stats_total[0])),
                       "mean_return": float(stats_total[1] / max(1.0, stats_total[0]))},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "api": "qr_step_host (pinned host actions in; obs, reward, done out; one env.step per call)",
                    "copy_ceiling": copy_ceiling,
                    "copy_ceiling_note": "the same bytes per step copied both ways on two streams with no kernel, all ranks at once: "
                                         "what the host links allow (%.1f GB/s aggregate)" % ((h2d + d2h) * world * e2e_steps / float(tc.item()) / 1e9)},
            "gpu_launches": launches,
            "roofline": {"bound": "fp%s_issue" % ("64" if args.dtype == "f64" else "32"), "achieved": ach_tf, "peak": fp_peak, "unit": "TFLOP/s",
                         "frac": ach_tf / fp_peak,
                         "achieved_with_measured_attempts": ach_tf_att, "frac_with_measured_attempts": ach_tf_att / fp_peak,
                         "algorithmic_flops_per_env_step": ALG_FLOPS, "env_steps_per_launch": env_steps_per_launch,
                         "kernel": "qr::k_step<%s, %s, %s>" % ("double" if args.dtype == "f64" else "float", fw, "multi-step" if (fused > 1 or args.policy) else "single-step"),
                         "kernel_ms": kernel_ms / launches_per_step if launches_per_step == 1 else kernel_ms,
                         "peak_source": "device query: %d SMs (cudaDevAttrMultiProcessorCount) x %d FP32 lanes x 2 flop x %.0f MHz (nvidia-smi clocks.max.sm); "
                                        "measured issue ceilings on this part: FFMA 100, FFMA2 116.5 lane-FMA/clk/SM of 128 (profiles/r02/r02a_ubench_pipes.txt)" % (
                                            sms, int(lanes), sm_max),
                         "traffic": (traffic["dram_bytes"] / 1e9) if traffic else None,
                         "traffic_unit": "GB per launch, dram__bytes_read.sum + dram__bytes_write.sum (%s)" % (traffic["source"] if traffic else "no capture of this configuration"),
                         "algorithmic_bytes_per_launch": alg_bytes_launch},
            "roofline_hbm": {"bound": "hbm", "achieved": alg_bytes_launch / (kernel_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                             "frac": alg_bytes_launch / (kernel_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": which,
                             "note": "not the binding bound: the K=1 launch (key k1) moves %d B per env-step" % bytes_per},
            "device": {"name": props.name, "sms": sms, "sm_max_mhz": sm_max, "total_memory_gb": props.total_memory / 2 ** 30},
        }
        if tracking is not None:
            tracking["fraction_of_headline"] = tracking["value"] / value
            line["tracking"] = tracking
        if k1 is not None:
            k1["fp_issue_frac"] = ALG_FLOPS * (k1["value"] / world) / 1e12 / fp_peak
            k1["hbm_frac"] = k1["hbm_gbs"] / hbm_peak
            line["k1"] = k1
    env.close()
    if rank == 0:
        if world == 1 and not args.no_extra:
            try:
                line["configs"] = sub_configs(torch, vec_env, dev, args.seed, args.steps < 10)
            except Exception as ex:  # pragma: no cover
                line["configs"] = {"error": repr(ex)}
        if world == 1 and not args.no_cpu:
            cfw = fw if fw != "QUAD" else "MONO"
            line["cpu_baseline"] = cpu_port_baseline(cfw, seconds=args.cpu_seconds)
            extra = {}
            try:
                extra["quad_v0_config1"] = cpu_quad_v0_config1()
                extra["modul" if cfw == "MONO" else "mono"] = cpu_port_baseline("MODUL" if cfw == "MONO" else "MONO", seconds=max(3.0, args.cpu_seconds / 3))
                extra["c_oracle"] = cpu_c_baseline(cfw)
            except Exception as ex:  # pragma: no cover
                extra["error"] = repr(ex)
            line["cpu_baselines"] = extra
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--framework", default="MONO", choices=["MONO", "MODUL", "QUAD"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--envs-per-gpu", type=int, default=1 << 21)
    ap.add_argument("--seed", type=int, default=1992)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU legs")
    ap.add_argument("--no-extra", action="store_true", help="skip the K=1 figure and the sub-configurations (profiling runs)")
    ap.add_argument("--actions", default="random", choices=["random", "zero"])
    ap.add_argument("--policy", action="store_true", help="config 5: the shipped TD3 actor in the loop")
    ap.add_argument("--goal", default="traj0", choices=["traj0", "hover", "circle", "eight"],
                    help="on-device trajectory generator mode (config 4: traj0; 'eight' etc. = trajectory tracking)")
    ap.add_argument("--fused", type=int, default=128,
                    help="env.step() calls per launch: 128 = the rollout of config 4 (qr_rollout, in-kernel actions); 1 = one step per launch, actions from HBM")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
