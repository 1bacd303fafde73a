"""GPU tests of round 2: in-kernel trajectory goals, the env swap on episode end in multi-step launches at size, the new
statistics, the trainer's return rounding, config overrides with derived values, qr_step_host ordering / validation, the
float32 horizon against the float64 mode, and the vector env's spaces.  Their kernel logic is also covered on the CPU
emulator (tests/test_host_twin_kernel.py)."""
import ctypes as C
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import quad_oracle as qo  # noqa: E402
from test_gpu_parity import _env, _t  # noqa: E402

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _start(e, env_type="train"):
    e.reset(env_type=env_type)
    e.init_goal()
    e.get_norm_error_state()
    e.stats()


@pytest.mark.parametrize("fw,goal", [("MONO", "hover"), ("MONO", "circle"), ("MONO", "eight"), ("MODUL", "eight")])
def test_trajectory_goal_inside_the_step_kernel_rollout_equals_single_steps(fw, goal):
    """Tracking workloads: the goal of every step is generated in the step kernel (get_desired before env.step,
    main.py:145-147).  One 24-step launch == 24 single-step launches == goal-update kernel + external-goal step, with auto
    reset restarting the trajectories (float64: bit-identical)."""
    n, K = 4096, 24
    kw = dict(seed=8, autoreset=True, goal_mode=goal, max_episode_steps=10)
    ea, eb = _env(n, fw, torch.float64, **kw), _env(n, fw, torch.float64, **kw)
    for e in (ea, eb):
        _start(e)
    g = torch.Generator(device="cuda:0"); g.manual_seed(4)
    acts = (torch.rand((K, n, ea.act_dim), device="cuda:0", generator=g, dtype=torch.float64) * 2 - 1) * 0.4
    obs_r, rew_r, done_r = ea.rollout(K, actions=acts, store=True)
    for k in range(K):
        obs, rew, done, _, _ = eb.step(acts[k])
        assert torch.equal(torch.cat(obs, dim=1), obs_r[k]) and torch.equal(rew, rew_r[k]) and torch.equal(done, done_r[k].bool()), k
    for name in ("state_soa", "integ_soa", "goal_soa", "traj_soa", "params_soa", "ep_length"):
        assert torch.equal(getattr(ea, name), getattr(eb, name)), name
    sa, sb = ea.stats(), eb.stats()
    assert sa[0] == sb[0] >= 2 * n and sa[7] == K * n
    assert float(ea.goal_soa[0:3].abs().max()) > 0 and float(ea.traj_soa[0].min()) > 0
    # the standalone generator call advances the same clock by one dt
    t0 = ea.traj_soa[0].clone()
    ea.goal_update()
    assert torch.allclose(ea.traj_soa[0], t0 + ea.dt, rtol=0, atol=1e-12)
    ea.close(); eb.close()


def test_decoupled_full_size_invariants_and_env_swap():
    """BASELINE config 3 size: DecoupledWrapper, 2^20 envs, float32, Philox actions, auto reset.  A 48-step launch in which
    every env ends two episodes (20-step limit): envs leave their lane, are reset in batches and go on in another lane --
    step accounting, episode accounting, bounded observations, R on SO(3), and agreement with single-step launches."""
    n, K = 1 << 20, 48
    kw = dict(seed=12, autoreset=True, goal_mode="traj0", max_episode_steps=20, diagnostics=False)
    ea, eb = _env(n, "MODUL", torch.float32, **kw), _env(n, "MODUL", torch.float32, **kw)
    for e in (ea, eb):
        _start(e)
    ea.rollout(K)
    for _ in range(K):
        eb.rollout(1)
    for name in ("state_soa", "integ_soa", "params_soa", "goal_soa", "obs", "reward", "done", "ep_length"):
        assert torch.equal(getattr(ea, name), getattr(eb, name)), name
    sa, sb = ea.stats(), eb.stats()
    assert sa[7] == sb[7] == K * n and sa[0] == sb[0] >= 2 * n and sa[3] == sb[3] and sa[5] == sb[5]
    assert sa[9] == sb[9] >= 14 * sa[7] and sa[10:14].sum() == 0 and int(ea.status.max()) == 0   # nfev = 2 + 12 per attempt; no histogram without diagnostics
    o = ea.obs
    assert bool(torch.isfinite(o).all()) and bool((o[:, 0:3].abs() < 1).all()) and bool((o[:, 6:9].abs() < 1).all())
    R = ea.state_soa[6:15].t().reshape(n, 3, 3)
    assert float((R @ R.transpose(1, 2) - torch.eye(3, device="cuda:0")).abs().max()) < 1e-3
    assert sa[16] == sb[16] == 0 and sa[14] == 0                                 # per-step sums are kept with diagnostics only
    ea.close(); eb.close()


def test_benchmark_reward_solved_count_and_rounded_returns():
    """Statistics 16 / 17 against the torch helpers (utils/utils.py:21-47, main.py:169-173) and round_returns against the
    trainer's float('{:.4f}'.format(ret + r)) (main.py:180)."""
    from gym_rotor_b200 import vec_env
    n, steps = 4096, 5
    env = _env(n, "MONO", torch.float64, seed=4, goal_mode="external", max_episode_steps=steps, round_returns=1)
    rng = np.random.default_rng(0)
    orc = qo.COracle("MONO")
    st, ig, par = orc.reset_from_uniforms(rng.random((n, 20)))
    st[: n // 2, 0:6] = 0.0; st[: n // 2, 6:15] = np.eye(3).reshape(-1); st[: n // 2, 15:18] = 0.0
    goal = np.zeros((n, 12)); goal[:, 6] = 1.0
    env.set_state(st, ig, par, goal)
    env.stats()
    ret = np.zeros(n); br = 0.0
    for k in range(steps):
        act = rng.uniform(-0.05, 0.05, (n, 4)); act[:, 0] = -0.06
        obs, rew, done, _, _ = env.step(_t(act, torch.float64))
        br += float(vec_env.benchmark_reward(obs, "MONO").sum())
        ret = np.array([float("{:.4f}".format(a + b)) for a, b in zip(ret, rew[:, 0].cpu().numpy())])
        assert np.array_equal(env.ep_return[0].cpu().numpy(), ret), k
    s = env.stats()
    assert abs(s[16] - br) < 1e-5 * br
    rel = vec_env.time_limit_relabel(obs, rew, done, "MONO")[:, 0]
    assert bool(env.truncated.bool().all()) and s[17] == int(rel.sum()) and 0 < s[17] < n
    env.close()


def test_config_overrides_recompute_derived_values():
    """A non-default Cx changes reward_min the way the reference derives it (quad.py:80-88): rewards against the oracle."""
    n = 2048
    rng = np.random.default_rng(3)
    orc = qo.COracle("MONO")
    orc.cfg.Cx = 9.5
    orc.cfg.reward_min = -np.ceil(9.5 + orc.cfg.CIx + orc.cfg.Cv + orc.cfg.Cb1 + orc.cfg.CIb1 + orc.cfg.CW)
    st, ig, par = orc.reset_from_uniforms(rng.random((n, 20)))
    goal = np.zeros((n, 12)); goal[:, 6] = 1.0
    act = rng.uniform(-1, 1, (n, 4))
    env = _env(n, "MONO", torch.float64, Cx=9.5)
    assert env.cfg.reward_min == orc.cfg.reward_min == -17.0
    env.set_state(st, ig, par, goal)
    obs, rew, done, _, _ = env.step(_t(act, torch.float64))
    o_ref, r_ref, d_ref, _, _ = orc.step(st.copy(), ig.copy(), par, goal, act)
    assert np.abs(rew.cpu().numpy() - r_ref).max() < 2e-7 and np.array_equal(done.cpu().numpy(), d_ref)
    env.close()


def test_step_host_is_ordered_after_the_callers_stream_and_validates_buffers():
    """qr_step_host runs on the handle's own streams: it must wait for what the caller enqueued before it (here: state
    injection kernels and a device step on the current torch stream), and later device steps must see its result.
    Trajectory goals are generated inside it as well; malformed host buffers are refused."""
    n = 50000
    rng = np.random.default_rng(5)
    kw = dict(seed=2, goal_mode="eight")
    e1, e2 = _env(n, "MONO", torch.float32, **kw), _env(n, "MONO", torch.float32, **kw)
    acts = [torch.as_tensor(rng.uniform(-0.5, 0.5, (n, 4)), dtype=torch.float32).pin_memory() for _ in range(3)]
    obs_h = torch.empty((n, 23), dtype=torch.float32).pin_memory()
    rew_h = torch.empty((n, 1), dtype=torch.float32).pin_memory()
    done_h = torch.empty((n, 1), dtype=torch.uint8).pin_memory()
    for e in (e1, e2):
        e.reset(); e.init_goal(); e.get_norm_error_state()
    e1.step(acts[0].cuda()); e1.step_host(acts[1], obs_h, rew_h, done_h); o1 = e1.step(acts[2].cuda())[0][0]
    for a in acts:
        o2, r2, d2, _, _ = e2.step(a.cuda())
    assert torch.equal(o1, o2[0]) and torch.equal(e1.state_soa, e2.state_soa) and torch.equal(e1.goal_soa, e2.goal_soa)
    with pytest.raises(ValueError):
        e1.step_host(acts[0][: n // 2], obs_h, rew_h, done_h)                  # short action array
    with pytest.raises(ValueError):
        e1.step_host(acts[0], obs_h.double(), rew_h, done_h)                   # wrong dtype
    with pytest.raises(ValueError):
        e1.step_host(acts[0], obs_h.t(), rew_h, done_h)                        # wrong shape / not contiguous
    with pytest.raises(ValueError):
        e1.step_host(acts[0].cuda(), obs_h, rew_h, done_h)                     # device memory
    e1.close(); e2.close()


def test_fp32_mode_follows_fp64_mode_over_a_horizon():
    """north_star: "the fp32 mode agrees to 1e-5 after one step, with observation and reward divergence reported over the
    horizon".  Both precisions on the device from identical states and actions, free running (tools/fp32_divergence_gpu.py
    writes the full table to profiles/); loose regression bounds on envs alive in both runs."""
    n, H = 4096, 100
    rng = np.random.default_rng(7)
    orc = qo.COracle("MONO")
    st, ig, par = orc.reset_from_uniforms(rng.random((n, 20)))
    goal = np.zeros((n, 12)); goal[:, 6] = 1.0
    e64, e32 = _env(n, "MONO", torch.float64), _env(n, "MONO", torch.float32)
    e64.set_state(st, ig, par, goal); e32.set_state(st, ig, par, goal)
    alive = torch.ones(n, dtype=torch.bool, device="cuda:0")
    worst = {1: None, 50: None, 100: None}
    flips = 0
    for t in range(1, H + 1):
        a = _t(rng.uniform(-0.3, 0.3, (n, 4)), torch.float32)
        o64, r64, d64, _, _ = e64.step(a)
        o32, r32, d32, _, _ = e32.step(a)
        flips += int((d64[:, 0] != d32[:, 0])[alive].sum())
        alive &= ~(d64[:, 0] | d32[:, 0])
        if t in worst:
            ds = (e64.state_soa.float() - e32.state_soa).abs().max(dim=0).values[alive].max()
            do = (o64[0] - o32[0]).abs().max(dim=1).values[alive].max()
            dr = (r64[:, 0].float() - r32[:, 0]).abs()[alive].max()
            worst[t] = (float(ds), float(do), float(dr))
    # observed on the B200 (profiles/r02/fp32_divergence_gpu.md): 2.4e-7 / 1.2e-7 / 1.5e-7 after one step, 5.3e-6 / 5.0e-6 /
    # 2.3e-6 after 100 steps with 1922 envs still flying, no done flag differing
    assert int(alive.sum()) > n // 4
    assert max(worst[1]) < 1e-5
    assert max(worst[50]) < 5e-5 and max(worst[100]) < 1e-4, worst
    assert flips <= 4, flips
    e64.close(); e32.close()


def test_vector_env_spaces_and_stats_summary():
    """N1: the vector env declares gymnasium-style spaces (quad.py:120-132 for Quad-v0; normalised boxes for the wrappers)
    and its samples step."""
    from gym_rotor_b200 import vec_env
    from gym_rotor_b200.dist import summarize
    for fw, O, A in (("QUAD", 18, 4), ("MONO", 23, 4)):
        ve = vec_env.QuadVectorEnv(256, framework=fw, max_episode_steps=50, dtype=torch.float32, seed=3)
        assert ve.single_observation_space.shape == (O,) and ve.single_action_space.shape == (A,)
        assert ve.observation_space.shape == (256, O) and ve.action_space.shape == (256, A)
        if fw == "QUAD":
            hi = ve.single_observation_space.high
            assert hi[0] == 1.0 and hi[3] == 4.0 and abs(hi[15] - 2 * np.pi) < 1e-6 and hi[6] == 1.0
        obs, info = ve.reset()
        assert tuple(obs.shape) == (256, O)
        a = torch.as_tensor(ve.action_space.sample(), device="cuda:0")
        obs, rew, term, trunc, info = ve.step(a)
        assert tuple(obs.shape) == (256, O) and term.dtype == torch.bool and "final_obs" in info
        if fw == "MONO":
            assert float(obs.abs().max()) <= 1.0 + 1e-6        # inside the declared normalised box
            s = summarize(ve.env.stats())
            assert 0.0 <= s["mean_benchmark_reward"] <= 1.0 and s["steps"] == 256
        ve.close()
    ve = vec_env.QuadVectorEnv(64, framework="MODUL", dtype=torch.float32)
    assert [sp.shape for sp in ve.single_observation_space.spaces] == [(15,), (3,)] and ve.single_action_space.shape == (5,)
    obs, _ = ve.reset()
    assert tuple(obs[0].shape) == (64, 15) and tuple(obs[1].shape) == (64, 3)
    ve.close()


@pytest.mark.parametrize("scale", [20.0, 150.0])
def test_fp32_mode_at_extreme_angular_rates(scale):
    """The float32 mode's shortcuts around ensure_SO3 (closed-form Euler probe, det R from the trace of R^T R - I, running
    maxima of the stage defects) where they matter most: |W| far beyond the termination limit, where stage matrices do leave
    SO(3) and attempts are redone with per-stage re-projection.  One step from the same states in both precisions.
    Observed on the B200: relative state difference 2.1e-7 (|W| <= 20) / 5.3e-7 (|W| <= 150), no done flag differing."""
    n = 4096
    rng = np.random.default_rng(0)
    orc = qo.COracle("MONO")
    st, ig, par = orc.reset_from_uniforms(rng.random((n, 20)))
    st[:, 15:18] = rng.uniform(-scale, scale, (n, 3))
    goal = np.zeros((n, 12)); goal[:, 6] = 1.0
    act = _t(rng.uniform(-1, 1, (n, 4)), torch.float32)
    e64, e32 = _env(n, "MONO", torch.float64), _env(n, "MONO", torch.float32)
    e64.set_state(st, ig, par, goal); e32.set_state(st, ig, par, goal)
    e64.stats(); e32.stats()
    o64, r64, d64, _, _ = e64.step(act)
    o32, r32, d32, _, _ = e32.step(act)
    s64, s32 = e64.state_soa.float(), e32.state_soa
    assert bool(torch.isfinite(s32).all()) and int(e32.status.max()) == 0 and int(e64.status.max()) == 0
    rel = (s64 - s32).abs().max(dim=0).values / (1 + s64.abs().max(dim=0).values)
    assert float(rel.max()) < 1e-5, float(rel.max())
    assert int((d64 != d32).sum()) <= 2
    assert e64.stats()[15] > 1000 and e32.stats()[15] > 1000       # re-projections happened in both modes
    e64.close(); e32.close()


@pytest.mark.parametrize("fw,dtype", [("MONO", torch.float32), ("MODUL", torch.float64)])
def test_one_step_rollout_with_storage_equals_step(fw, dtype):
    """qr_rollout(n_steps = 1) into the caller's arrays is served by the multi-step kernel (the single-step kernel carries no
    code for rollout storage): rows, rewards, dones and every env array against a plain step, resets in both."""
    n = 4099
    kw = dict(autoreset=True, goal_mode="traj0", max_episode_steps=3, seed=8)
    e1, e2 = _env(n, fw, dtype, **kw), _env(n, fw, dtype, **kw)
    _start(e1); _start(e2)
    rng = np.random.default_rng(21)
    for t in range(5):
        act = _t(rng.uniform(-1, 1, (n, e1.act_dim)), dtype)
        obs_r, rew_r, done_r = e1.rollout(1, act[None].contiguous(), store=True)
        obs, rew, done, _, _ = e2.step(act)
        assert torch.equal(e1.obs, e2.obs) and torch.equal(e1.final_obs, e2.final_obs), t
        assert torch.equal(rew_r[0], rew) and torch.equal(done_r[0].bool(), done)
        keep = ~(e2.terminated.bool() | e2.truncated.bool())
        assert torch.equal(obs_r[0][keep], torch.cat(obs, dim=1)[keep])
        assert torch.equal(e1.state_soa, e2.state_soa) and torch.equal(e1.integ_soa, e2.integ_soa)
    s1, s2 = e1.stats(), e2.stats()
    assert np.array_equal(s1, s2) and s1[0] >= n
    e1.close(); e2.close()


def test_create_refuses_more_envs_than_the_kernel_indexes():
    """Env indices are 32-bit inside the step kernel: qr_create caps a handle at 2^30 envs (before it allocates anything)."""
    from gym_rotor_b200 import _native as nat
    L = nat.load()
    cfg = nat.QrConfig()
    nat.check(L.qr_default_config(C.byref(cfg), 1, nat.F32))
    cfg.n_envs = (1 << 30) + 1
    h = C.c_void_p()
    rc = L.qr_create(C.byref(cfg), 0, C.byref(h))
    assert rc != 0 and not h.value
    assert b"2^30" in L.qr_last_error()
