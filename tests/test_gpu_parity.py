"""GPU parity tests: the sm_100a step kernel (through the C ABI) against the reference's golden vectors and
the CPU oracle.  Run on the B200 box with `pytest -m gpu`.

Tolerances (BASELINE.json north_star):
  * fp64 mode: state / integrals within 1e-12 relative per step of the reference; <= 1e-9 after a free-running
    horizon on envs still alive;
  * fp32 mode: within 1e-5 of the float64 reference after one step;
  * float32 observations of the fp64 mode: bit-exact except for double-rounding flips (a 1e-16 state
    difference landing on a float32 rounding boundary; expected ~2e-8 per element) -- at most a handful;
  * done flags and RHS-evaluation counts: identical; reward: exact given identical observations.
"""
import os
import zlib

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import quad_oracle as qo  # noqa: E402

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _env(n, fw, dtype=torch.float64, **kw):
    from gym_rotor_b200 import vec_env
    return vec_env.BatchedQuadEnv(n, framework=fw, dtype=dtype, **kw)


def _load(name):
    return np.load(os.path.join(G, name))


def _t(a, dtype):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device="cuda:0")


def _relerr(a, b):
    return np.abs(a - b).max() / max(1.0, np.abs(b).max())


@pytest.mark.parametrize("fw,tag,a32", [("MONO", "mono", False), ("MONO", "mono", True),
                                        ("MODUL", "modul", False), ("MODUL", "modul", True)])
def test_fp64_step_matches_reference_golden(fw, tag, a32):
    g = _load("step_%s_%s.npz" % (tag, "a32" if a32 else "a64"))
    n = g["action"].shape[0]
    env = _env(n, fw)
    env.set_state(g["state_in"], g["integ_in"], g["params"], g["goal"])
    act = _t(g["action"], torch.float32 if a32 else torch.float64)
    obs, rew, done, _, _ = env.step(act)
    st, ig, _, _ = env.get_state()
    assert _relerr(st, g["state_out"]) <= 1e-12
    assert np.abs(ig - g["integ_out"]).max() <= 1e-12
    o = torch.cat(obs, dim=1).cpu().numpy()
    flips = int((o.view(np.uint32) != g["obs"].view(np.uint32)).sum())
    assert flips <= 3, flips
    assert np.abs(o - g["obs"]).max() <= 1.2e-7
    assert (done.cpu().numpy() == g["done"]).all()
    assert (env.nfev.cpu().numpy() == g["nfev"]).all()
    r = rew.cpu().numpy()
    assert np.abs(r - g["reward"]).max() <= 2e-7 and (r != g["reward"]).mean() <= 5e-3
    assert int(env.status.max()) == 0
    env.close()


def test_fp64_kat1_flight_log():
    rows = _load("kat1_modul_log.npz")["rows"]
    n = rows.shape[0] - 1
    env = _env(n, "MODUL")
    par = np.tile(np.array([2.15, 0.23, 0.022, 0.035, 0.0135, 2.2]), (n, 1))
    goal = np.zeros((n, 12)); goal[:, 6] = 1.0
    env.set_state(rows[:-1, 5:23], np.zeros((n, 8)), par, goal)
    env.step(_t(rows[:-1, 0:5], torch.float64))
    st = env.get_state()[0]
    assert np.abs(st - rows[1:, 5:23]).max() < 2e-10      # the log is printed with %.10f
    assert (env.nfev.cpu().numpy() == 14).all()
    env.close()


@pytest.mark.parametrize("fw,tag", [("MONO", "mono"), ("MODUL", "modul")])
def test_fp32_step_within_1e5_of_reference(fw, tag):
    g = _load("step_%s_a64.npz" % tag)
    n = g["action"].shape[0]
    env = _env(n, fw, torch.float32)
    env.set_state(g["state_in"], g["integ_in"], g["params"], g["goal"])
    obs, rew, done, _, _ = env.step(_t(g["action"], torch.float32))
    st, ig, _, _ = env.get_state()
    assert np.abs(st - g["state_out"]).max() <= 1e-5
    assert np.abs(ig - g["integ_out"]).max() <= 1e-5
    o = torch.cat(obs, dim=1).cpu().numpy()
    assert np.abs(o - g["obs"]).max() <= 1e-5
    assert np.abs(rew.cpu().numpy() - g["reward"]).max() <= 1e-4
    # done flags may only differ where an error is within float32 rounding of the limit
    mism = done.cpu().numpy() != g["done"]
    assert mism.mean() <= 2e-3
    attempts_ref = (g["nfev"] - 2) // 12
    attempts = (env.nfev.cpu().numpy() - 2) // 12
    assert (attempts != attempts_ref).mean() <= 0.03
    env.close()


def test_fp64_free_running_4096x1000_vs_oracle():
    """4096 envs, train-reset states, U(-1,1) actions, 1000 steps, no re-sync."""
    n, H = 4096, 1000
    rng = np.random.default_rng(123)
    orc = qo.COracle("MONO", threads=qo.lib().qo_get_max_threads())
    st, ig, par = orc.reset_from_uniforms(rng.random((n, 20)))
    goal = np.zeros((n, 12)); goal[:, 6] = 1.0
    env = _env(n, "MONO")
    env.set_state(st, ig, par, goal)
    alive = np.ones(n, bool)
    first_done_ref = np.full(n, -1); first_done_gpu = np.full(n, -1)
    worst = 0.0
    arng = np.random.default_rng(1)
    for t in range(H):
        a = arng.uniform(-1, 1, (n, 4))
        o_ref, r_ref, d_ref, nf_ref, _ = orc.step(st, ig, par, goal, a)
        obs, rew, done, _, _ = env.step(_t(a, torch.float64))
        d_gpu = done.cpu().numpy()[:, 0]
        first_done_ref[(first_done_ref < 0) & d_ref[:, 0]] = t
        first_done_gpu[(first_done_gpu < 0) & d_gpu] = t
        if t % 100 == 99 or t == H - 1:
            sg = env.get_state()[0]
            ok = alive & ~d_ref[:, 0]
            if ok.any():
                worst = max(worst, np.abs(sg[ok] - st[ok]).max())
        alive &= ~d_ref[:, 0]
        if t == 99:
            assert (d_gpu == d_ref[:, 0]).all()
    assert (first_done_ref == first_done_gpu).all(), "first-done step indices must be identical"
    assert worst <= 1e-9, worst
    env.close()


def test_fp64_batch512_reference_free_run():
    """512 envs x 100 steps run by the reference itself (no resets): final state, rewards and done history."""
    g = _load("batch512.npz")
    N, H = 512, 100
    actions = np.random.default_rng(1).uniform(-1, 1, size=(H, N, 4))
    assert np.uint32(zlib.crc32(actions.tobytes())) == g["actions_crc"]
    env = _env(N, "MONO")
    goal = np.tile(g["goal"], (N, 1))
    env.set_state(g["state0"], np.zeros((N, 8)), g["params"], goal)
    obs_crc_ok = 0
    for t in range(H):
        obs, rew, done, _, _ = env.step(_t(actions[t], torch.float64))
        assert (done.cpu().numpy()[:, 0] == g["done"][t]).all(), t
        assert (env.nfev.cpu().numpy() == g["nfev"][t]).all(), t
        assert np.abs(rew.cpu().numpy()[:, 0] - g["reward"][t]).max() <= 2e-7
        obs_crc_ok += int(np.uint32(zlib.crc32(obs[0].cpu().numpy().tobytes())) == g["obs_crc"][t])
    st, ig, _, _ = env.get_state()
    assert _relerr(st, g["stateT"]) <= 1e-9
    assert np.abs(ig - g["integT"]).max() <= 1e-9
    assert np.abs(obs[0].cpu().numpy() - g["obs_last"]).max() <= 1.2e-7
    assert obs_crc_ok >= 0.9 * H        # whole-batch observations bit-exact on nearly every step
    env.close()


@pytest.mark.parametrize("integ", ["solve_ivp", "euler"])
def test_quad_v0_base_env(integ):
    g = _load("quad_v0.npz")
    n = g[integ + "_action"].shape[0]
    env = _env(n, "QUAD", integrator=integ)
    env.set_state(g[integ + "_state_in"], np.zeros((n, 8)), g[integ + "_params"], g[integ + "_goal"])
    obs, rew, done, _, _ = env.step(_t(g[integ + "_action"], torch.float64))
    st = env.get_state()[0]
    assert np.abs(st - g[integ + "_state_out"]).max() <= 1e-12
    assert (done.cpu().numpy() == g[integ + "_done"]).all()
    assert np.abs(rew.cpu().numpy() - g[integ + "_reward"]).max() <= 1e-12
    env.close()


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_fused_rollout_equals_single_steps(dtype):
    n, K = 1000, 12
    rng = np.random.default_rng(5)
    orc = qo.COracle("MODUL")
    st, ig, par = orc.reset_from_uniforms(rng.random((n, 20)))
    goal = np.zeros((n, 12)); goal[:, 6] = 1.0
    acts = _t(rng.uniform(-1, 1, (K, n, 5)), dtype)
    e1 = _env(n, "MODUL", dtype); e2 = _env(n, "MODUL", dtype)
    e1.set_state(st, ig, par, goal); e2.set_state(st, ig, par, goal)
    obs_r, rew_r, done_r = e1.rollout(K, acts, store=True)
    for k in range(K):
        obs, rew, done, _, _ = e2.step(acts[k])
        assert torch.equal(torch.cat(obs, dim=1), obs_r[k]) and torch.equal(rew, rew_r[k])
        assert torch.equal(done, done_r[k].bool())
    assert torch.equal(e1.state_soa, e2.state_soa) and torch.equal(e1.integ_soa, e2.integ_soa)
    assert torch.equal(e1.obs, e2.obs)
    e1.close(); e2.close()


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_fused_rollout_with_autoreset_equals_single_steps(dtype):
    """Multi-step launches reset an env inside the kernel (the env keeps stepping in its lane), single-step launches
    queue the env and reset a batch at a time: same episodes, observations and states.  A 7-step time limit makes
    every env truncate at once (a burst larger than the queue batch); n is odd, so the rows of the rollout storage
    are not 16-byte aligned (the coalesced observation path must fall back to per-lane stores)."""
    n, K = 1001, 20
    rng = np.random.default_rng(15)
    acts = _t(rng.uniform(-1, 1, (K, n, 4)), dtype)
    kw = dict(seed=5, autoreset=True, goal_mode="traj0", max_episode_steps=7)
    e1 = _env(n, "MONO", dtype, **kw); e2 = _env(n, "MONO", dtype, **kw)
    for e in (e1, e2):
        e.reset(); e.init_goal(); e.get_norm_error_state()
    obs_r, rew_r, done_r = e1.rollout(K, acts, store=True)
    for k in range(K):
        obs, rew, done, _, _ = e2.step(acts[k])
        assert torch.equal(obs[0], obs_r[k]), k
        assert torch.equal(rew, rew_r[k]) and torch.equal(done, done_r[k].bool())
    assert torch.equal(e1.state_soa, e2.state_soa) and torch.equal(e1.integ_soa, e2.integ_soa)
    assert torch.equal(e1.params_soa, e2.params_soa) and torch.equal(e1.goal_soa, e2.goal_soa)
    assert torch.equal(e1.obs, e2.obs)
    s1, s2 = e1.stats(), e2.stats()
    assert s1[0] == s2[0] and s1[0] >= 2 * n and s1[7] == K * n      # episodes ended, env-steps
    e1.close(); e2.close()


def _philox_uniforms(seed, gids, episode):
    u = np.empty((len(gids), 20))
    for i, gid in enumerate(gids):
        for j in range(5):
            w = qo.philox4x32_10([gid & 0xFFFFFFFF, gid >> 32, episode, j], [seed & 0xFFFFFFFF, seed >> 32])
            u[i, 4 * j:4 * j + 4] = [(k + 0.5) * 2.0 ** -32 for k in w]
    return u


@pytest.mark.parametrize("env_type", ["train", "eval"])
def test_reset_matches_oracle_and_is_sharding_independent(env_type):
    n, seed = 512, 77
    env = _env(n, "MONO", seed=seed)
    s32 = env.reset(env_type=env_type)
    st, ig, par, _ = env.get_state()
    u = _philox_uniforms(seed, list(range(n)), 1)
    st_o, ig_o, par_o = qo.COracle("MONO").reset_from_uniforms(u, qo.ENV_TRAIN if env_type == "train" else qo.ENV_EVAL)
    assert np.abs(st - st_o).max() <= 1e-14 and np.abs(par - par_o).max() <= 1e-15 and (ig == 0).all()
    assert s32.dtype == torch.float32 and np.abs(s32.cpu().numpy() - st).max() < 1e-6
    # two shards of 256 with env_id_offset reproduce the same envs (Philox key = global env id)
    a = _env(256, "MONO", seed=seed, env_id_offset=0); b = _env(256, "MONO", seed=seed, env_id_offset=256)
    a.reset(env_type=env_type); b.reset(env_type=env_type)
    assert np.array_equal(np.concatenate([a.get_state()[0], b.get_state()[0]]), st)
    # masked reset only touches the selected envs and advances their episode index
    mask = torch.zeros(n, dtype=torch.uint8, device="cuda:0"); mask[::3] = 1
    env.reset(env_type=env_type, mask=mask)
    st2 = env.get_state()[0]
    m = mask.cpu().numpy().astype(bool)
    assert np.array_equal(st2[~m], st[~m]) and (np.abs(st2[m] - st[m]).max(axis=1) > 0).all()
    for e in (env, a, b):
        e.close()


def test_goal_mode0_matches_reference_trajgen():
    """On-device trajectory generator mode 0: Wd from the pre-step state equals the reference's goal."""
    g = _load("step_mono_a64.npz")
    n = g["action"].shape[0]
    env = _env(n, "MONO", goal_mode="traj0")
    goal_in = g["goal"].copy(); goal_in[:, 9:12] = 7.0      # Wd must be recomputed in-kernel
    env.set_state(g["state_in"], g["integ_in"], g["params"], goal_in)
    obs, rew, done, _, _ = env.step(_t(g["action"], torch.float64))
    st, ig, _, gl = env.get_state()
    ok = np.ones(n, bool)
    assert np.abs(gl[ok, 9:12] - g["goal"][ok, 9:12]).max() <= 1e-12
    assert _relerr(st[ok], g["state_out"][ok]) <= 1e-12
    o = obs[0].cpu().numpy()
    assert np.abs(o[ok] - g["obs"][ok]).max() <= 1.2e-7
    env.close()


def test_autoreset_equals_manual_protocol():
    """In-kernel auto reset == step -> reset(mask) -> init_goal(mask) -> get_norm_error_state(mask) (main.py:212-230)."""
    n, seed = 2048, 3
    rng = np.random.default_rng(9)
    a = _env(n, "MONO", seed=seed, autoreset=True, goal_mode="traj0", max_episode_steps=25)
    b = _env(n, "MONO", seed=seed, autoreset=False, goal_mode="traj0")
    for e in (a, b):
        e.reset(); e.init_goal(); e.get_norm_error_state()
    assert torch.equal(a.state_soa, b.state_soa) and torch.equal(a.obs, b.obs)
    steps_b = torch.zeros(n, dtype=torch.int32, device="cuda:0")
    n_resets = 0
    for t in range(60):
        act = _t(rng.uniform(-1, 1, (n, 4)), torch.float64)
        oa, ra, da, _, _ = a.step(act)
        ob, rb, db, _, _ = b.step(act)
        steps_b += 1
        assert torch.equal(ra, rb) and torch.equal(da, db)
        need = (db[:, 0] | (steps_b >= 25))
        assert torch.equal(a.terminated.bool(), db[:, 0]) and torch.equal(a.truncated.bool(), steps_b >= 25)
        if need.any():
            term_obs = b.obs.clone()
            b.reset(mask=need); b.init_goal(mask=need); b.get_norm_error_state(mask=need)
            steps_b[need] = 0
            assert torch.equal(a.final_obs[need], term_obs[need])
            n_resets += int(need.sum())
        assert torch.equal(a.obs, b.obs), t
        assert torch.equal(a.state_soa, b.state_soa) and torch.equal(a.goal_soa, b.goal_soa)
        assert torch.equal(a.params_soa, b.params_soa) and torch.equal(a.integ_soa, b.integ_soa)
    s = a.stats()
    assert n_resets > n and s[0] == n_resets and s[7] == 60 * n
    a.close(); b.close()


def test_step_host_equals_device_step():
    n = 40000
    rng = np.random.default_rng(2)
    orc = qo.COracle("MONO")
    st, ig, par = orc.reset_from_uniforms(rng.random((n, 20)))
    goal = np.zeros((n, 12)); goal[:, 6] = 1.0
    e1 = _env(n, "MONO", torch.float32); e2 = _env(n, "MONO", torch.float32)
    e1.set_state(st, ig, par, goal); e2.set_state(st, ig, par, goal)
    act = torch.as_tensor(rng.uniform(-1, 1, (n, 4)), dtype=torch.float32).pin_memory()
    obs_h = torch.empty((n, 23), dtype=torch.float32).pin_memory()
    rew_h = torch.empty((n, 1), dtype=torch.float32).pin_memory()
    done_h = torch.empty((n, 1), dtype=torch.uint8).pin_memory()
    e1.step_host(act, obs_h, rew_h, done_h)
    obs, rew, done, _, _ = e2.step(act.cuda())
    assert torch.equal(obs_h, obs[0].cpu()) and torch.equal(rew_h, rew.cpu()) and torch.equal(done_h.bool(), done.cpu())
    assert torch.equal(e1.state_soa, e2.state_soa)
    e1.close(); e2.close()


def test_random_action_rollout_statistics():
    """In-kernel Philox actions + auto reset: the DOP853 attempt histogram and episode length look like the
    reference's under random actions (~94 % single attempt, mean episode ~110 steps)."""
    n = 16384
    env = _env(n, "MONO", torch.float32, autoreset=True, goal_mode="traj0", max_episode_steps=4000, seed=11)
    env.reset(); env.init_goal(); env.get_norm_error_state()
    env.stats()
    env.rollout(400)
    s = env.stats()
    steps = s[7]
    assert steps == 400 * n
    assert 0.90 < s[10] / steps < 0.98
    assert s[0] > 0 and 60 < s[3] / s[0] < 200
    assert s[8] == 0
    env.close()


def test_missing_library_fails_loudly(monkeypatch):
    from gym_rotor_b200 import _native
    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setattr(_native, "LIB_PATH", "/nonexistent/libquadrotor_b200.so")
    with pytest.raises(_native.NativeError):
        _native.load()


def test_init_goal_matches_oracle_trajgen():
    """qr_init_goal == mark_traj_start + get_desired(mode 0) on the float32-cast reset state (main.py:226-229),
    with theta taken from the env's Philox stream (uniform #19 of the episode)."""
    n, seed = 256, 5
    env = _env(n, "MONO", seed=seed, goal_mode="traj0")
    env.reset(); env.init_goal()
    st, _, _, gl = env.get_state()
    u = _philox_uniforms(seed, list(range(n)), 1)
    theta = np.deg2rad(-25.0 + 50.0 * u[:, 19])
    st32 = st.astype(np.float32).astype(np.float64)
    b1d = qo.traj_init_mode0(st32, theta)
    assert np.abs(gl[:, 6:9] - b1d).max() <= 1e-14
    assert np.abs(gl[:, 9:12] - qo.traj_wd(st32, b1d)).max() <= 1e-13
    assert np.abs(gl[:, 0:6]).max() == 0
    env.close()


def test_fp32_reset_is_a_valid_state():
    n = 4096
    env = _env(n, "MODUL", torch.float32, seed=9)
    env.reset()
    st, ig, par, _ = env.get_state()
    R = st[:, 6:15].reshape(n, 3, 3).transpose(0, 2, 1)
    assert np.abs(R @ R.transpose(0, 2, 1) - np.eye(3)).max() < 5e-6 and np.abs(np.linalg.det(R) - 1).max() < 5e-6
    assert np.abs(st[:, 0:3]).max() <= 0.6 + 1e-6 and np.abs(st[:, 3:6]).max() <= 2.0 + 1e-6
    assert np.abs(st[:, 15:18]).max() <= np.pi + 1e-5 and (ig == 0).all()
    still = np.abs(st[:, 0:6]).sum(axis=1) == 0
    assert 0.15 < still.mean() < 0.25                      # 20 % spawn at the origin (quad.py:342)
    assert np.abs(par[:, 0] - 2.15).max() <= 0.2151 and par[:, 0].std() > 0.05
    env.close()


def test_vector_env_facade():
    from gym_rotor_b200 import vec_env
    n = 2048
    ve = vec_env.QuadVectorEnv(n, framework="MODUL", max_episode_steps=90, dtype=torch.float32, seed=1)
    obs, info = ve.reset()
    assert obs[0].shape == (n, 15) and obs[1].shape == (n, 3) and obs[0].dtype == torch.float32
    g = torch.Generator(device="cuda:0"); g.manual_seed(0)
    n_trunc = n_term = 0
    for t in range(120):
        a = torch.rand((n, 5), device="cuda:0", generator=g) * 2 - 1
        obs, rew, term, trunc, info = ve.step(a)
        assert rew.shape == (n, 2) and term.shape == (n,) and trunc.shape == (n,)
        assert torch.isfinite(obs[0]).all() and torch.isfinite(rew).all()
        assert ((rew >= 0) & (rew <= 1) | (rew == -1)).all()
        n_trunc += int(trunc.sum()); n_term += int(term.sum())
        # a finished env already shows the first observation of its next episode: integral terms restart near 0
        fin = term | trunc
        if fin.any():
            assert obs[0][fin, 3:6].abs().max() < 0.01
    s = ve.env.stats()
    assert n_trunc > 0 and n_term > 0 and s[0] == n_term + int(s[5]) and s[7] == 120 * n
    assert abs(s[10:14].sum() - s[7]) < 0.5
    ve.close()


def test_status_flags_nonfinite_state():
    """scipy raises on a non-finite y0; the kernel flags the env instead and leaves the others untouched."""
    n = 64
    rng = np.random.default_rng(0)
    st, ig, par = qo.COracle("MONO").reset_from_uniforms(rng.random((n, 20)))
    st[5, 2] = np.nan
    goal = np.zeros((n, 12)); goal[:, 6] = 1.0
    env = _env(n, "MONO")
    env.set_state(st, ig, par, goal)
    env.step(_t(rng.uniform(-1, 1, (n, 4)), torch.float64))
    status = env.status.cpu().numpy()
    assert status[5] & 1 and (np.delete(status, 5) == 0).all()
    assert np.isfinite(np.delete(env.get_state()[0], 5, axis=0)).all()
    env.close()


TRAJ_CASES = [("hover", "hover"), ("circle", "circle"), ("eight", "eight"), ("circle_manual", "circle")]


@pytest.mark.parametrize("name,gm", TRAJ_CASES)
def test_trajectory_modes_match_reference(name, gm):
    """qr_init_goal + qr_goal_update (modes 1 - 6 and the manual fallback) call by call against the reference's
    TrajectoryGenerator driven along a real flight (tests/golden/traj_modes.npz)."""
    g = _load("traj_modes.npz")
    st, goal_ref, bdd_ref, t_ref = g[name + "_state"], g[name + "_goal"], g[name + "_b1d_dot"], g[name + "_t"]
    t_traj, w, smooth, theta0 = g[name + "_draws"]
    env = _env(1, "MONO", goal_mode=gm)
    par = np.array([[2.15, 0.23, 0.022, 0.035, 0.0135, 2.2]])
    env.set_state(st[0:1], np.zeros((1, 8)), par, None)
    env.init_goal()                                 # mark_traj_start + first get_desired on the float32 reset state
    ts = env.traj_soa
    if name == "hover":                             # inject the reference run's two random draws
        ts[8, 0] = t_traj; ts[7, 0] = smooth; ts[6, 0] = w
    if name == "circle_manual":
        ts[8, 0] = 1.75                             # the golden run shortened the circle to reach manual mode
    start = 1 if name == "hover" else 0
    if start == 0:
        assert np.abs(env.get_state()[3][0] - goal_ref[0]).max() < 1e-11
    worst = 0.0
    for i in range(1, len(t_ref)):
        env.set_state(st[i:i + 1], None, None, None)
        env.goal_update()
        gl = env.get_state()[3][0]
        worst = max(worst, np.abs(gl - goal_ref[i]).max(), float((ts[9:11, 0].cpu() - torch.as_tensor(bdd_ref[i, 0:2])).abs().max()))
        if i % 50 == 0:
            assert abs(float(ts[0, 0]) - t_ref[i]) < 1e-12
    assert worst < 1e-11, worst
    assert bool(int(ts[1, 0]) & 2) == bool(g[name + "_manual"][-1])
    env.close()


def test_trajectory_modes_batch_vs_oracle_with_philox_draws():
    """Hover mode on 512 envs: the per-env random draws come from the env's Philox stream (block 5 of the episode)."""
    n, seed = 512, 13
    env = _env(n, "MODUL", seed=seed, goal_mode="hover")
    env.reset(); env.init_goal()
    st = env.get_state()[0]
    draws = np.empty((n, 2))
    for i in range(n):
        wds = qo.philox4x32_10([i, 0, 1, 5], [seed, 0])
        draws[i] = [(wds[0] + 0.5) * 2.0 ** -32, (wds[1] + 0.5) * 2.0 ** -32]
    st32 = st.astype(np.float32).astype(np.float64)
    ts = qo.traj_start(st32)
    goal = np.zeros((n, 12)); goal[:, 6] = 1.0
    qo.traj_desired(1, st32, ts, goal, draws)
    assert np.abs(env.get_state()[3] - goal).max() < 1e-12
    rng = np.random.default_rng(0)
    for _ in range(5):
        obs, rew, done, _, _ = env.step(_t(rng.uniform(-0.2, 0.2, (n, 5)), torch.float64))   # step() runs goal_update first
        # the oracle sees the same pre-step state the goal kernel saw
        qo.traj_desired(1, st, ts, goal, draws)
        assert np.abs(env.get_state()[3] - goal).max() < 1e-12
        st = env.get_state()[0]
    assert np.abs(env.traj_soa.t().cpu().numpy() - ts).max() < 1e-12
    env.close()


def test_trajectory_mode_autoreset_restarts_the_trajectory():
    n = 1024
    env = _env(n, "MONO", torch.float32, seed=4, goal_mode="eight", autoreset=True, max_episode_steps=40)
    env.reset(); env.init_goal(); env.get_norm_error_state()
    g = torch.Generator(device="cuda:0"); g.manual_seed(1)
    for t in range(45):
        env.step(torch.rand((n, 4), device="cuda:0", generator=g) * 0.2 - 0.1)
    tclock = env.traj_soa[0].cpu().numpy()
    # every env was reset at step 40 (time limit) at the latest; a restart makes one get_desired call (t = dt) and
    # every later step one more, so the trajectory clock follows the episode length
    eplen = env.ep_length.cpu().numpy()
    assert eplen.max() <= 40 and np.abs(tclock - (1 + eplen) * 0.005).max() < 1e-6
    assert int(env.stats()[0]) >= n
    env.close()


@pytest.mark.parametrize("scale", [20.0, 150.0])
def test_extreme_angular_rates_exercise_stage_reprojection(scale):
    """Far outside the termination limits (|W| <= 2 pi) DOP853 stage matrices do leave SO(3) by more than 1e-5 and
    the reference re-projects them inside EoM.  The kernel runs its stages speculatively and redoes such an attempt
    in checked mode: results and RHS counts must still equal the oracle's stage-by-stage evaluation."""
    n = 2048
    rng = np.random.default_rng(0)
    orc = qo.COracle("MONO", threads=qo.lib().qo_get_max_threads())
    st, ig, par = orc.reset_from_uniforms(rng.random((n, 20)))
    st[:, 15:18] = rng.uniform(-scale, scale, (n, 3))
    goal = np.zeros((n, 12)); goal[:, 6] = 1.0
    act = rng.uniform(-1, 1, (n, 4))
    env = _env(n, "MONO")
    env.set_state(st, ig, par, goal)
    qo.lib().qo_stage_projection_count(1)
    o_ref, r_ref, d_ref, nfev, status = orc.step(st, ig, par, goal, act)
    assert qo.lib().qo_stage_projection_count(1) > 100          # the slow path really is exercised
    obs, rew, done, _, _ = env.step(_t(act, torch.float64))
    sg, igg, _, _ = env.get_state()
    assert (env.nfev.cpu().numpy() == nfev).all()
    assert _relerr(sg, st) <= 1e-11 and np.abs(igg - ig).max() <= 1e-11
    assert (done.cpu().numpy() == d_ref).all()
    env.close()


def test_cabi_error_paths_and_independent_handles():
    """int return codes + qr_last_error instead of exceptions; two handles on one device do not interact."""
    import ctypes as C
    from gym_rotor_b200 import _native as nat
    L = nat.load()
    cfg = nat.QrConfig(); assert L.qr_default_config(C.byref(cfg), nat.MODE_COUPLED, nat.F32) == 0
    h = C.c_void_p()
    cfg.n_envs = 0
    assert L.qr_create(C.byref(cfg), 0, C.byref(h)) == 1 and b"n_envs" in L.qr_last_error()
    cfg.n_envs = 64; cfg.integrator = nat.INT_EULER
    assert L.qr_create(C.byref(cfg), 0, C.byref(h)) == 1 and b"Euler" in L.qr_last_error()
    cfg.integrator = nat.INT_DOP853
    assert L.qr_create(C.byref(cfg), 99, C.byref(h)) == 1 and b"device" in L.qr_last_error()
    assert L.qr_create(C.byref(cfg), 0, C.byref(h)) == 0
    assert L.qr_step(h, None, nat.F32, None) == 1 and b"null actions" in L.qr_last_error()
    act = torch.zeros((64, 4), device="cuda:0")
    assert L.qr_step(h, C.c_void_p(act.data_ptr()), 7, None) == 1
    assert L.qr_rollout(h, 0, None, nat.F32, None, None, None, None) == 1
    assert L.qr_stats(h, None, 0, None) == 1
    assert L.qr_step(None, C.c_void_p(act.data_ptr()), nat.F32, None) == 1
    assert L.qr_destroy(h) == 0 and L.qr_destroy(None) == 0
    # two independent handles, interleaved on different streams
    a = _env(512, "MONO", torch.float32, seed=1); b = _env(512, "MONO", torch.float32, seed=1)
    a.reset(); b.reset()
    act = torch.rand((512, 4), device="cuda:0") * 2 - 1
    s1 = torch.cuda.Stream()
    torch.cuda.synchronize()
    with torch.cuda.stream(s1):
        a.step(act)
    b.step(act)
    torch.cuda.synchronize()
    assert torch.equal(a.state_soa, b.state_soa) and torch.equal(a.obs, b.obs)
    a.close(); b.close()
