#!/usr/bin/env python
"""env-steps/sec of the batched quadrotor step on N B200s (BASELINE.json metric), one JSON line on rank 0.

    python bench.py --gpus 1 --steps 200 --warmup 20
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...        # the reference's CPU path (scipy port) on the host cores

Workload (the configuration the 1e10 steps/s target is quoted on): CoupledWrapper,
float32 state arithmetic, 2^21 envs per GPU (= 2^24 over 8 GPUs, weak scaling), U(-1,1) actions resident in
HBM, on-device trajectory-generator goals (mode 0), in-kernel auto reset with the trainer's 4000-step
limit, one all-reduce of the 16 episode statistics every 128 steps.  One bench "step" = one env.step() of
every env = one kernel launch through the C ABI (qr_step).
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

ALG_BYTES = {"MONO_f32": 366, "MONO_f64": 614, "MODUL_f32": 363}   # SURVEY 8(d), per env-step
ALG_FLOPS = 5450                                                   # SURVEY 8(d), one DOP853 attempt
STATS_EVERY = 128
DRAM_BYTES_PER_ENV_STEP_NCU = 428.5   # ncu --set full capture r01u (steady state, ~32 k resets in the launch): 898.6 MB per launch of 2^21 env-steps


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), d.get("sm_max_mhz", 1965.0), "measured"
    return 6650.0, 1965.0, "fallback"


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
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------------------
# CPU baselines (the only place bench.py executes oracle/)
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
    return {"value": steps / max(r[1] for r in res), "unit": "env-steps/s", "cores": procs, "kind": "port",
            "sample": "%d procs x %.0f s of %s ScipyPort.step (numpy RHS + scipy DOP853), U(-1,1) actions, train resets; "
                      "%d steps, wall %.1f s" % (procs, seconds, framework, steps, wall)}


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
    return {"value": best, "unit": "env-steps/s", "cores": thr, "kind": "port",
            "sample": "C oracle (gcc -O2, float64), %d envs x 1 step, best of %d, %d pthreads" % (n, reps, thr)}


# ------------------------------------------------------------------------------------------------------

def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    seconds = max(5.0, min(30.0, 1.5 * (args.steps + args.warmup) / 10.0))
    base = cpu_port_baseline("MONO", seconds=seconds)
    line = {"impl": "reference", "metric": "env-steps/sec", "value": base["value"], "unit": "env-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / base["value"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            # same workload as the CUDA arm (CoupledWrapper env.step under U(-1,1) actions with the trainer's reset
            # protocol and mode-0 goals); the envs are stepped one per host core instead of 2^21 per GPU
            "config": {"workload": "CoupledWrapper env.step, U(-1,1) actions, reset on termination: the reference's CPU "
                                   "path (numpy RHS + scipy solve_ivp DOP853), one env per host core, bounded sample "
                                   "(fixed goal: the trajectory generator's per-step goal update is not in the sample)",
                       "framework": "MONO", "actions": "random", "envs_per_gpu": args.envs_per_gpu},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


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
    env = vec_env.BatchedQuadEnv(n, framework=fw, dtype=dtype, device=dev, seed=args.seed, autoreset=True,
                                 goal_mode=(args.goal if fw != "QUAD" else "external"),
                                 env_type="eval" if args.policy else "train", max_episode_steps=1000 if args.policy else 4000,
                                 env_id_offset=rank * n, diagnostics=False)
    env.reset(env_type="eval" if args.policy else "train")
    if fw != "QUAD":
        env.init_goal()
    env.get_norm_error_state()
    A = env.act_dim
    gen = torch.Generator(device=dev); gen.manual_seed(1234 + rank)
    if args.actions == "zero":   # hover-like thrust, no torque: the single-attempt, no-reset regime of a trained policy
        pool = [torch.zeros((n, A), device=dev, dtype=torch.float32) for _ in range(2)]
        for p in pool:
            p[:, 0] = -0.06
    else:
        pool = [torch.rand((n, A), device=dev, dtype=torch.float32, generator=gen) * 2 - 1 for _ in range(4)]
    stats_total = np.zeros(16)
    fused = max(1, args.fused)
    actors = None
    if args.policy:   # BASELINE config 5: the reference's shipped TD3 actor in the loop (obs -> action on device)
        actors = torch.empty((n, env.act_dim), dtype=torch.float32, device=dev)   # action buffer of qr_policy_td3

    def one_step(i):
        if actors is not None and fused > 1:
            env.rollout(fused, actions="policy")   # obs -> shipped actor -> env.step, `fused` times, in ONE launch
        elif actors is not None:
            env.step(env.policy_td3(out=actors))   # compiled actor kernel + step kernel: two launches per env.step
        elif fused > 1:
            env.rollout(fused)      # `fused` env.step() calls in one launch, Philox actions drawn in-kernel
        else:
            env.step(pool[i % len(pool)])
        if (i + 1) % STATS_EVERY == 0:
            return allreduce_stats(env, dev) if world > 1 else env.stats()
        return None

    sampler = ClockSampler(local); sampler.start()   # samples under load: warm-up + timed region
    for i in range(args.warmup):
        one_step(i)
    env.stats()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    launches0 = env.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        s = one_step(i)
        if s is not None:
            stats_total += s
    ev1.record()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    ms = ev0.elapsed_time(ev1)
    launches = env.launch_count() - launches0
    clocks = sampler.stop()
    stats_total += (allreduce_stats(env, dev) if world > 1 else env.stats())
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    total_steps = float(n) * world * args.steps * fused
    value = total_steps / (ms * 1e-3)

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

    if rank == 0:
        hbm_peak, sm_max, which = _peaks()
        key = "%s_%s" % (fw, args.dtype)
        bytes_per = ALG_BYTES.get(key, 366)
        kernel_ms = ms / args.steps
        ach_gbs = bytes_per * n * fused / (kernel_ms * 1e-3) / 1e9
        per_gpu_steps = n * fused / (kernel_ms * 1e-3)
        fp_peak = 148 * 128 * 2 * sm_max * 1e6 / 1e12 * (0.5 if args.dtype == "f64" else 1.0)
        attempts = stats_total[10:14]
        mean_att = float((attempts * np.array([1, 2, 3, 4])).sum() / max(1.0, attempts.sum()))
        line = {
            "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": "%s env.step, %d envs/GPU (2^24 over 8 GPUs), U(-1,1) actions resident in HBM, "
                                   "on-device trajgen %s goals, in-kernel auto reset (4000-step limit), "
                                   "stats all-reduce every %d steps" % (
                                       {"MONO": "CoupledWrapper", "MODUL": "DecoupledWrapper", "QUAD": "Quad-v0"}[fw], n,
                                       {"traj0": "mode-0"}.get(args.goal, args.goal), STATS_EVERY),
                       "envs_per_gpu": n, "framework": fw, "actions": ("shipped TD3 actor in the loop (qr_policy_td3, compiled effective weights)" if args.policy else args.actions),
                       "fused_steps_per_launch": fused,
                       "l2": "per-step working set %.0f MB > 126 MB L2 (inputs larger than L2)" % (bytes_per * n / 1e6),
                       "mean_dop853_attempts": mean_att, "episodes": stats_total[0],
                       "mean_episode_length": float(stats_total[3] / max(1.0, stats_total[0]))},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "api": "qr_step_host (pinned host actions in; obs, reward, done out)"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                         "traffic": (DRAM_BYTES_PER_ENV_STEP_NCU * n * fused / 1e9) if (fw == "MONO" and args.dtype == "f32") else None,
                         "traffic_unit": "GB per launch (dram__bytes_read.sum + dram__bytes_write.sum, profiles/r01u_kstep_f32_ncu_digest.txt)",
                         "peak_source": which, "kernel": "qr::k_step<%s>" % ("double" if args.dtype == "f64" else "float"),
                         "algorithmic_bytes_per_env_step": bytes_per, "kernel_ms": kernel_ms},
            "roofline_fp": {"bound": "fp%s issue" % ("64" if args.dtype == "f64" else "32"),
                            "achieved": ALG_FLOPS * mean_att * per_gpu_steps / 1e12, "peak": fp_peak, "unit": "TFLOP/s",
                            "frac": ALG_FLOPS * mean_att * per_gpu_steps / 1e12 / fp_peak,
                            "note": "algorithmic flops = 5450 per DOP853 attempt x measured mean attempts; peak = 148 SM x 128 lanes x 2 x max SM clock"},
        }
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_port_baseline(fw if fw != "QUAD" else "MONO", seconds=args.cpu_seconds)
            try:
                line["cpu_baseline_c"] = cpu_c_baseline(fw if fw != "QUAD" else "MONO")
            except Exception as ex:  # pragma: no cover
                line["cpu_baseline_c"] = {"error": str(ex)}
        print(json.dumps(line), flush=True)
    env.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--framework", default="MONO", choices=["MONO", "MODUL", "QUAD"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--envs-per-gpu", type=int, default=1 << 21)
    ap.add_argument("--seed", type=int, default=1992)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--actions", default="random", choices=["random", "zero"])
    ap.add_argument("--policy", action="store_true", help="config 5: the shipped TD3 actor in the loop")
    ap.add_argument("--goal", default="traj0", choices=["traj0", "hover", "circle", "eight"],
                    help="on-device trajectory generator mode (config 4: traj0; 'eight' etc. = trajectory tracking, one extra "
                         "goal-update kernel per step)")
    ap.add_argument("--fused", type=int, default=1, help="env.step() calls fused per launch (qr_rollout, in-kernel actions)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
