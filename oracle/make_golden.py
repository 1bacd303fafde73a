#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference (/root/reference).

TEST INFRASTRUCTURE.  Runs only in the build container (the reference tree does not travel to the GPU
box); the produced .npz files are committed so that every test can run without the reference.

    python oracle/make_golden.py            # all fixtures
    python oracle/make_golden.py step kat1  # a subset

Fixtures (float64 unless noted; every array is [T, ...] with one row per env.step() call):
  step_{mono,modul}_{a64,a32}.npz  free-running episodes driven exactly like main.py:126-129,140-164,212-230
        (reset -> trajgen.mark_traj_start/get_desired(mode 0) -> set_goal_state -> get_norm_error_state ->
        loop) with U(-1,1) actions; per step: state_in, integ_in, params, goal, action, state_out,
        integ_out, obs (f32), reward, done, nfev.  a32 = the action array is float32 (numpy then computes the
        thrust in float32, coupled_yaw_wrapper.py:46-48).
  kat1_modul_log.npz   the reference's own flight log results/MODUL_log_20250303_120200.dat (3600 x 40).
  reset_samples.npz    reference reset('train') / reset('eval') draws (state + parameters), float32.
  quad_v0.npz          base Quad-v0 env (T1..T4 actions), DOP853 and Euler integrators.
  batch512.npz         512 envs x 100 steps, no resets: initial state/params, final state/integrals and the
                       full reward/done history (SURVEY 8(d) config 2 at a size the reference finishes in ~2 min).
"""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _save(name, **arrs):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **arrs)
    print("wrote %s (%.0f KB)" % (path, os.path.getsize(path) / 1024))


def gen_step(framework, act32, n_steps, seed):
    env = rh.make_env(framework)
    tg = rh.make_trajgen(env)
    cnt = rh.RhsCounter(env)
    rh.seed_all(seed)
    A = 4 if framework == "MONO" else 5
    rec = {k: [] for k in ("state_in", "integ_in", "params", "goal", "action", "state_out", "integ_out", "obs",
                           "reward", "done", "nfev", "reset_state32", "episode_start")}
    # main.py:126-129
    state32 = env.reset(env_type='train', seed=seed)
    xd, vd, b1d, b1d_dot, Wd = tg.get_desired(state32, 0)
    env.set_goal_state(xd, vd, b1d, b1d_dot, Wd)
    env.get_norm_error_state(framework)
    new_episode = True
    ep_steps = 0
    for _ in range(n_steps):
        ep_steps += 1
        # main.py:145-147
        st = env.get_current_state()
        xd, vd, b1d, b1d_dot, Wd = tg.get_desired(st, 0)
        env.set_goal_state(xd, vd, b1d, b1d_dot, Wd)
        a = np.random.rand(A) * 2 - 1  # main.py:155
        if act32:
            a = a.astype(np.float32)
        rec["state_in"].append(np.array(env.state, dtype=np.float64))
        rec["integ_in"].append(rh.get_integ(env))
        rec["params"].append(rh.get_params(env))
        rec["goal"].append(rh.get_goal(env))
        rec["action"].append(a.astype(np.float64))
        rec["episode_start"].append(new_episode)
        rec["reset_state32"].append(np.array(state32, dtype=np.float32))
        new_episode = False
        obs, rew, done, _, _ = env.step(a.copy())
        rec["nfev"].append(cnt.take())
        rec["state_out"].append(np.array(env.state, dtype=np.float64))
        rec["integ_out"].append(rh.get_integ(env))
        rec["obs"].append(np.concatenate(obs).astype(np.float32))
        rec["reward"].append(np.array(rew, dtype=np.float64))
        rec["done"].append(np.array(done, dtype=bool))
        if any(done) or ep_steps == 400:  # main.py:212-230 (shorter time limit to see more resets)
            state32 = env.reset(env_type='train', seed=seed)
            tg.mark_traj_start(state32)
            xd, vd, b1d, b1d_dot, Wd = tg.get_desired(state32, 0)
            env.set_goal_state(xd, vd, b1d, b1d_dot, Wd)
            env.get_norm_error_state(framework)
            cnt.take()
            new_episode = True
            ep_steps = 0
    return {k: np.array(v) for k, v in rec.items()}


def make_step():
    for fw, tag in (("MONO", "mono"), ("MODUL", "modul")):
        _save("step_%s_a64.npz" % tag, **gen_step(fw, False, 1200, seed=11))
        _save("step_%s_a32.npz" % tag, **gen_step(fw, True, 400, seed=12))


def make_kat1():
    rows = np.loadtxt(os.path.join(rh.REFERENCE_ROOT, "results", "MODUL_log_20250303_120200.dat"))
    assert rows.shape == (3600, 40), rows.shape
    _save("kat1_modul_log.npz", rows=rows)


def make_reset():
    out = {}
    for fw in ("MONO",):
        env = rh.make_env(fw)
        rh.seed_all(2024)
        for env_type, n in (("train", 8192), ("eval", 1024)):
            st = np.empty((n, 18), np.float32); par = np.empty((n, 6), np.float32)
            for i in range(n):
                env.reset(env_type=env_type)
                st[i] = env.state; par[i] = rh.get_params(env)
            out["state_" + env_type] = st
            out["params_" + env_type] = par
    _save("reset_samples.npz", **out)


def make_quad():
    out = {}
    for integ in ("solve_ivp", "euler"):
        env = rh.make_base_env()
        env.ode_integrator = integ
        rh.seed_all(5)
        env.reset(env_type='train')
        rec = {k: [] for k in ("state_in", "params", "goal", "action", "state_out", "reward", "done")}
        rng = np.random.default_rng(6)
        for t in range(300):
            goal = np.concatenate([rng.uniform(-.3, .3, 3), rng.uniform(-.3, .3, 3),
                                   [np.cos(0.7), np.sin(0.7), 0.], np.zeros(3)])
            rh.set_goal(env, goal)
            a = rng.uniform(-1, 1, 4)
            rec["state_in"].append(np.array(env.state, np.float64)); rec["params"].append(rh.get_params(env))
            rec["goal"].append(goal); rec["action"].append(a)
            obs, rew, done, _, _ = env.step(a.copy())
            rec["state_out"].append(np.array(env.state, np.float64))
            rec["reward"].append(np.array(rew, np.float64)); rec["done"].append(np.array(done, bool))
            if done[0]:
                env.reset(env_type='train')
        for k, v in rec.items():
            out["%s_%s" % (integ, k)] = np.array(v)
    _save("quad_v0.npz", **out)


def make_batch():
    N, H = 512, 100
    env = rh.make_env("MONO")
    st0 = np.empty((N, 18)); par = np.empty((N, 6))
    for s in range(N):
        rh.seed_all(s)
        env.reset(env_type='train')
        st0[s] = env.state; par[s] = rh.get_params(env)
    actions = np.random.default_rng(1).uniform(-1, 1, size=(H, N, 4))
    goal = np.zeros(12); goal[6] = 1.0
    stT = np.empty((N, 18)); igT = np.empty((N, 8))
    reward = np.empty((H, N)); done = np.empty((H, N), bool); nfev = np.empty((H, N), np.int16)
    obs_crc = np.empty(H, np.uint32)
    obs_all = np.empty((H, N, 23), np.float32)
    cnt = rh.RhsCounter(env)
    for s in range(N):
        rh.set_params(env, par[s]); env.state = st0[s].copy(); rh.set_integ(env, np.zeros(8)); rh.set_goal(env, goal)
        cnt.take()
        for t in range(H):
            o, r, d, _, _ = env.step(actions[t, s].copy())
            reward[t, s] = r[0]; done[t, s] = d[0]; nfev[t, s] = cnt.take(); obs_all[t, s] = o[0]
        stT[s] = env.state; igT[s] = rh.get_integ(env)
        if s % 64 == 0:
            print("batch env", s, flush=True)
    for t in range(H):
        obs_crc[t] = zlib.crc32(obs_all[t].tobytes())
    _save("batch512.npz", state0=st0, params=par, goal=goal, stateT=stT, integT=igT, reward=reward, done=done,
          nfev=nfev, obs_crc=obs_crc, obs_last=obs_all[-1], actions_crc=np.uint32(zlib.crc32(actions.tobytes())))


def make_traj():
    """Trajectory generator modes 1 (hover), 2 (take-off), 3 (land), 4 (stay), 5 (circle), 6 (figure eight) and the
    manual-mode fallback after a trajectory completes (utils/trajectory_generator.py:113-173, 232-505), driven like
    main.py:304-331 with the shipped
    MONO actor keeping the vehicle in the air.  One row per get_desired() call: input state, outputs, clock."""
    import make_policy_fixture as mpf
    args, agents = mpf.build_agents("MONO")
    out = {}
    only = os.environ.get("QR_TRAJ_ONLY")    # e.g. "takeoff,land,stay": regenerate a subset, keep the other arrays
    if only and os.path.exists(os.path.join(OUT, "traj_modes.npz")):
        out.update(dict(np.load(os.path.join(OUT, "traj_modes.npz"))))
    for name, mode, steps, tweak in (("hover", 1, 600, None), ("circle", 5, 600, None), ("eight", 6, 700, None),
                                     ("circle_manual", 5, 500, "short"), ("takeoff", 2, 1700, "snap"), ("land", 3, 400, "high"),
                                     ("land_low", 3, 60, None), ("stay", 4, 120, None)):
        if only and name not in only.split(","):
            continue
        env = rh.make_env("MONO")
        tg = rh.make_trajgen(env)
        if tweak == "short":
            tg.num_circles = 0          # t_traj = 1.75 s: the circle ends after its straight segment -> manual mode
        rh.seed_all(21 + mode)
        state32 = env.reset(env_type="eval")
        if tweak == "high":             # landing from above the cut-off height: t_traj > 0, all three branches of land()
            for sd in range(100, 400):
                rh.seed_all(sd)
                state32 = env.reset(env_type="eval")
                if state32[2] < -0.33:
                    break
        tg.mark_traj_start(state32)
        rec = {k: [] for k in ("state", "goal", "b1d_dot", "t", "manual")}
        xd, vd, b1d, b1d_dot, Wd = tg.get_desired(state32, mode)
        rec["state"].append(np.array(state32, np.float64)); rec["goal"].append(np.concatenate([xd, vd, b1d, Wd]))
        rec["b1d_dot"].append(np.array(b1d_dot, np.float64)); rec["t"].append(tg.t); rec["manual"].append(tg.manual_mode)
        env.set_goal_state(xd, vd, b1d, b1d_dot, Wd)
        obs_n = env.get_norm_error_state("MONO")
        for _ in range(steps):
            st = env.get_current_state()
            if tweak == "snap" and tg.t > tg.t_traj + 0.5:
                # the shipped actor holds ~6 cm of steady-state error, the take-off only completes within 4 cm of the
                # waypoint: hand the generator a state that is there (it only consumes states)
                st = np.array(st, np.float64); st[0:3] = np.asarray(tg.xd, np.float64) + np.array([0.01, -0.005, 0.01])
            xd, vd, b1d, b1d_dot, Wd = tg.get_desired(st, mode)
            rec["state"].append(np.array(st, np.float64)); rec["goal"].append(np.concatenate([xd, vd, b1d, Wd]))
            rec["b1d_dot"].append(np.array(b1d_dot, np.float64)); rec["t"].append(tg.t); rec["manual"].append(tg.manual_mode)
            env.set_goal_state(xd, vd, b1d, b1d_dot, Wd)
            act = agents[0].choose_action(obs_n[0], explor_noise_std=0.)
            obs_n, rew, done, _, _ = env.step(act.copy())
            if done[0]:
                break
        for k, v in rec.items():
            out["%s_%s" % (name, k)] = np.array(v)
        out["%s_draws" % name] = np.array([float(getattr(tg, "t_traj", 0.0)), float(getattr(tg, "w_b1d", 0.0)),
                                           float(getattr(tg, "smooth_term", 0.0)), float(tg.theta_init)])
        print(name, "calls", len(rec["t"]), "manual at end", rec["manual"][-1], "draws", out["%s_draws" % name])
    _save("traj_modes.npz", **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["step", "kat1", "reset", "quad", "batch", "traj"]
    for w in which:
        {"step": make_step, "kat1": make_kat1, "reset": make_reset, "quad": make_quad, "batch": make_batch,
         "traj": make_traj}[w]()
