#!/usr/bin/env python
"""float32 mode against float64 mode ON THE DEVICE over a free-running horizon (north_star: "the fp32 mode agrees to 1e-5
after one step, with observation and reward divergence reported over the horizon").  Both handles start from the same
reset states and receive the same float32 actions; envs are compared while alive in both runs.
usage: python tools/fp32_divergence_gpu.py [n_envs] [horizon] [action scale] > profiles/r02/fp32_divergence_gpu.md"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import quad_oracle as qo
from gym_rotor_b200 import vec_env

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
H = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
scale = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
marks = [1, 2, 5, 10, 20, 50, 100, 200, 500, 1000, 2000, 4000]
print("# float32 mode vs float64 mode on the device (%s), free running\n" % torch.cuda.get_device_name(0))
for fw, A in (("MONO", 4), ("MODUL", 5)):
    rng = np.random.default_rng(7)
    orc = qo.COracle(fw)
    st, ig, par = orc.reset_from_uniforms(rng.random((n, 20)))
    goal = np.zeros((n, 12)); goal[:, 6] = 1.0
    e64 = vec_env.BatchedQuadEnv(n, framework=fw, dtype=torch.float64)
    e32 = vec_env.BatchedQuadEnv(n, framework=fw, dtype=torch.float32)
    e64.set_state(st, ig, par, goal); e32.set_state(st, ig, par, goal)
    alive = torch.ones(n, dtype=torch.bool, device="cuda:0")
    done_flips = att_diff = steps_alive = 0
    print("## %s, %d envs, actions U(-%.2g, %.2g), train-reset initial states, fixed goal\n" % (fw, n, scale, scale))
    print("| step | envs alive in both | max abs state diff | max abs obs diff | max abs reward diff | done flags differing so far | steps with a different attempt count so far |")
    print("|---|---|---|---|---|---|---|")
    for t in range(1, H + 1):
        a = torch.as_tensor(rng.uniform(-scale, scale, (n, A)), dtype=torch.float32, device="cuda:0")
        o64, r64, d64, _, _ = e64.step(a)
        o32, r32, d32, _, _ = e32.step(a)
        d_any64, d_any32 = d64.any(dim=1), d32.any(dim=1)
        done_flips += int((d_any64 != d_any32)[alive].sum())
        att_diff += int((e64.nfev != e32.nfev)[alive].sum())
        steps_alive += int(alive.sum())
        alive &= ~(d_any64 | d_any32)
        if t in marks and int(alive.sum()) > 0:
            ds = (e64.state_soa.float() - e32.state_soa).abs().max(dim=0).values[alive].max()
            do = (torch.cat(o64, dim=1) - torch.cat(o32, dim=1)).abs().max(dim=1).values[alive].max()
            dr = (r64.float() - r32).abs().max(dim=1).values[alive].max()
            print("| %d | %d | %.2e | %.2e | %.2e | %d | %d (%.3f %%) |" % (t, int(alive.sum()), float(ds), float(do), float(dr), done_flips, att_diff,
                                                                      100.0 * att_diff / max(1, steps_alive)))
    print()
    e64.close(); e32.close()

# ---- closed loop: the shipped TD3 actor flies both runs for a whole 1000-step evaluation episode -------------------------
for fw, A in (("MONO", 4), ("MODUL", 5)):
    kw = dict(framework=fw, goal_mode="traj0", env_type="eval", seed=5)
    e64 = vec_env.BatchedQuadEnv(n, dtype=torch.float64, **kw)
    e32 = vec_env.BatchedQuadEnv(n, dtype=torch.float32, **kw)
    for e in (e64, e32):
        e.reset(env_type="eval"); e.init_goal()
    # identical initial conditions: the float32 run starts from the float32-rounded state of the float64 run
    st, ig, par, gl = e64.get_state()
    st, ig, par, gl = (x.astype(np.float32).astype(np.float64) for x in (st, ig, par, gl))
    e64.set_state(st, ig, par, gl); e32.set_state(st, ig, par, gl)
    for e in (e64, e32):
        e.get_norm_error_state()
    print("## %s, %d envs, closed loop: each run is flown by the reference's shipped TD3 actor on its own observations (eval resets, mode-0 goals)\n" % (fw, n))
    print("| step | envs alive in both | max abs state diff | max abs obs diff | max abs reward diff | max abs return diff | done flags differing so far |")
    print("|---|---|---|---|---|---|---|")
    alive = torch.ones(n, dtype=torch.bool, device="cuda:0")
    ret64 = torch.zeros(n, dtype=torch.float64, device="cuda:0"); ret32 = torch.zeros_like(ret64)
    flips = 0
    for t in range(1, min(H, 1000) + 1):
        o64, r64, d64, _, _ = e64.step(e64.policy_td3())
        o32, r32, d32, _, _ = e32.step(e32.policy_td3())
        ret64 += r64[:, 0]; ret32 += r32[:, 0].double()
        flips += int((d64.any(dim=1) != d32.any(dim=1))[alive].sum())
        alive &= ~(d64.any(dim=1) | d32.any(dim=1))
        if t in marks and int(alive.sum()) > 0:
            ds = (e64.state_soa.float() - e32.state_soa).abs().max(dim=0).values[alive].max()
            do = (torch.cat(o64, dim=1) - torch.cat(o32, dim=1)).abs().max(dim=1).values[alive].max()
            dr = (r64.float() - r32).abs().max(dim=1).values[alive].max()
            print("| %d | %d | %.2e | %.2e | %.2e | %.2e | %d |" % (t, int(alive.sum()), float(ds), float(do), float(dr),
                                                               float((ret64 - ret32).abs()[alive].max()), flips))
    print("\nmean return over %d envs: float64 %.4f, float32 %.4f\n" % (n, float(ret64.mean()), float(ret32.mean())))
    e64.close(); e32.close()
