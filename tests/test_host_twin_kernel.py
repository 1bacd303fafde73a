"""The REAL step kernel, run WITHOUT a GPU.

tests/host_twin/twin_kernel.cpp compiles qr::k_step itself (gym_rotor_b200/csrc/qr_kernels.cuh: the per-lane state
machine, the shared-memory stash, the fetch-ahead of the next env, the reset queue and the parked reset call, both
observation paths, the statistics) for the host and runs it on the one-warp SIMT emulator of tests/host_twin/simt.h
(32 fibers, every warp collective a rendezvous; a collective that not all lanes reach aborts as a deadlock).  These
tests drive it over host arrays laid out like the device buffers, against the reference's golden vectors and against
the per-env functions of tests/host_twin/twin.cpp.  Test infrastructure only: the package never loads the twin and
has no CPU path; timing and the float32 MUFU approximations remain GPU-only matters (tests/test_gpu_parity.py).
"""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")
TWIN_DIR = os.path.join(ROOT, "tests", "host_twin")
CUDA_INC = "/usr/local/cuda/include"

pytestmark = pytest.mark.skipif(shutil.which("g++") is None or not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")),
                                reason="needs g++ and the CUDA headers")


class TwArrays(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("state", "integ", "params", "goal", "traj", "obs", "reward", "done", "terminated",
                                          "truncated", "final_obs", "nfev", "status", "ep_return", "ep_length", "ep_index",
                                          "stats", "actions")] + [("act_f32", C.c_int)] + \
               [(n, C.c_void_p) for n in ("obs_roll", "reward_roll", "done_roll")]


def _build(src, tmp, name, opt="-O1", defines=()):
    out = os.path.join(tmp, name)
    cmd = ["g++", "-std=c++17", opt, "-ffp-contract=off", "-fPIC", "-shared"] + ["-D" + d for d in defines] + ["-I" + CUDA_INC, "-I" + TWIN_DIR,
           "-I" + os.path.join(ROOT, "gym_rotor_b200", "csrc"), "-o", out, os.path.join(TWIN_DIR, src)]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0, res.stdout
    return C.CDLL(out)


@pytest.fixture(scope="module")
def libs(tmp_path_factory):
    tmp = str(tmp_path_factory.mktemp("twink"))
    # QR_TWIN_DEFINES="NAME=1,...": run the whole file against a kernel variant built with extra macros
    extra = tuple(d for d in os.environ.get("QR_TWIN_DEFINES", "").split(",") if d)
    K = _build("twin_kernel.cpp", tmp, "libtwink.so", defines=extra)
    K.tw_kstep.argtypes = [C.c_void_p, C.POINTER(TwArrays), C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int]
    K.tw_actor.argtypes = [C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    K.tw_companion.argtypes = [C.c_void_p, C.POINTER(TwArrays), C.c_int, C.c_void_p, C.c_int]
    F = _build("twin.cpp", tmp, "libtwin.so", "-O2")
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    F.tw_step.argtypes = [C.c_void_p, dp, dp, dp, dp, dp, C.c_int, dp, dp, C.POINTER(C.c_float), dp, ip, ip, ip]
    F.tw_reset.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, dp, dp, dp, dp]
    F.tw_norm_error_state.argtypes = [C.c_void_p, dp, dp, dp, C.POINTER(C.c_float)]
    return K, F


def _config(mode, dtype64=True, **kw):
    from gym_rotor_b200 import _native as nat
    cfg = nat.QrConfig()
    nat.check(nat.load().qr_default_config(C.byref(cfg), mode, nat.F64 if dtype64 else nat.F32))
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


class HostEnv:
    """Host arrays with the layout of qr_buffers + one emulated k_step launch per call."""

    def __init__(self, K, cfg, warps=3):
        self.K, self.cfg, self.warps = K, cfg, warps
        n = self.n = int(cfg.n_envs)
        self.T = np.float64 if cfg.dtype == 1 else np.float32
        self.O = 23 if cfg.mode == 1 else 18
        self.A = 5 if cfg.mode == 2 else 4
        self.G = 2 if cfg.mode == 2 else 1
        T = self.T
        self.state = np.zeros((18, n), T); self.state[6] = 1; self.state[10] = 1; self.state[14] = 1
        self.integ = np.zeros((8, n), T); self.params = np.zeros((6, n), T); self.goal = np.zeros((12, n), T); self.goal[6] = 1
        self.traj = np.zeros((12, n), T)
        S = int(K.tw_obs_stride(int(cfg.mode)))        # padded row stride (obs_stride_of)
        self._obs_base = np.zeros((n, S), np.float32); self._final_obs_base = np.zeros((n, S), np.float32)
        self.obs = self._obs_base[:, :self.O]; self.final_obs = self._final_obs_base[:, :self.O]
        self.reward = np.zeros((n, self.G), T); self.done = np.zeros((n, self.G), np.uint8)
        self.terminated = np.zeros(n, np.uint8); self.truncated = np.zeros(n, np.uint8)
        self.nfev = np.zeros(n, np.int32); self.status = np.zeros(n, np.uint8)
        self.ep_return = np.zeros((2, n), T); self.ep_length = np.zeros(n, np.int32); self.ep_index = np.zeros(n, np.uint32)
        self.stats = np.zeros(20, np.float64)   # QR_NUM_STATS

    def set_state(self, st, ig, par, goal):          # [n,18] [n,8] [n,6] [n,12] like vec_env.set_state
        self.state[:] = np.asarray(st).T; self.integ[:] = np.asarray(ig).T
        self.params[:] = np.asarray(par).T; self.goal[:] = np.asarray(goal).T

    def _arrays(self):
        b = TwArrays()
        for name in ("state", "integ", "params", "goal", "traj", "obs", "reward", "done", "terminated", "truncated", "final_obs",
                     "nfev", "status", "ep_return", "ep_length", "ep_index", "stats"):
            setattr(b, name, getattr(self, name).ctypes.data)
        return b

    def companion(self, which, mask=None, env_type=0):
        """k_reset / k_init_goal / k_goal_update / k_norm_error_state (qr_reset, qr_init_goal, qr_goal_update, ...)."""
        b = self._arrays()
        mp = None if mask is None else np.ascontiguousarray(mask, np.uint8).ctypes.data
        assert self.K.tw_companion(C.byref(self.cfg), C.byref(b), {"reset": 0, "init_goal": 1, "goal_update": 2,
                                                                 "norm_error_state": 3}[which], mp, env_type) == 0

    def launch(self, actions=None, n_steps=1, store=False, lo=0, hi=None):
        b = self._arrays()
        keep = []
        policy = isinstance(actions, str) and actions == "policy"
        if actions is not None and not policy:
            actions = np.ascontiguousarray(actions)
            assert actions.dtype in (np.float32, np.float64) and actions.shape[-2:] == (self.n, self.A)
            keep.append(actions)
            b.actions = actions.ctypes.data; b.act_f32 = int(actions.dtype == np.float32)
        out = None
        if store:
            out = (np.zeros((n_steps, self.n, self.O), np.float32), np.zeros((n_steps, self.n, self.G), self.T),
                   np.zeros((n_steps, self.n, self.G), np.uint8))
            b.obs_roll, b.reward_roll, b.done_roll = (x.ctypes.data for x in out)
        rc = self.K.tw_kstep(C.byref(self.cfg), C.byref(b), int(lo), self.n if hi is None else int(hi), int(n_steps), self.warps, int(policy))
        assert rc == 0
        return out


def _relerr(a, b):
    return np.abs(a - b).max() / max(1.0, np.abs(b).max())


@pytest.mark.parametrize("fw,tag,a32", [("MONO", "mono", False), ("MONO", "mono", True),
                                        ("MODUL", "modul", False), ("MODUL", "modul", True)])
def test_kernel_fp64_step_matches_reference_golden(libs, fw, tag, a32):
    """test_gpu_parity.py::test_fp64_step_matches_reference_golden on the emulated kernel."""
    K, _ = libs
    g = np.load(os.path.join(G, "step_%s_%s.npz" % (tag, "a32" if a32 else "a64")))
    n = g["action"].shape[0]
    env = HostEnv(K, _config(1 if fw == "MONO" else 2, n_envs=n))
    env.set_state(g["state_in"], g["integ_in"], g["params"], g["goal"])
    env.launch(g["action"].astype(np.float32 if a32 else np.float64))
    assert _relerr(env.state.T, g["state_out"]) <= 1e-12
    assert np.abs(env.integ.T - g["integ_out"]).max() <= 1e-12
    flips = int((env.obs.view(np.uint32) != g["obs"].view(np.uint32)).sum())
    assert flips <= 3, flips
    assert (env.done.astype(bool) == g["done"]).all()
    assert (env.nfev == g["nfev"]).all()
    assert np.abs(env.reward - g["reward"]).max() <= 2e-7
    assert int(env.status.max()) == 0 and env.stats[7] == n


@pytest.mark.parametrize("fw,tag", [("MONO", "mono"), ("MODUL", "modul")])
def test_kernel_fp32_step_within_1e5(libs, fw, tag):
    K, _ = libs
    g = np.load(os.path.join(G, "step_%s_a64.npz" % tag))
    n = g["action"].shape[0]
    env = HostEnv(K, _config(1 if fw == "MONO" else 2, dtype64=False, n_envs=n), warps=12)
    env.set_state(g["state_in"], g["integ_in"], g["params"], g["goal"])
    env.launch(g["action"].astype(np.float32))
    assert np.abs(env.state.T - g["state_out"]).max() <= 1e-5
    assert np.abs(env.obs - g["obs"]).max() <= 1e-5
    assert (env.done.astype(bool) != g["done"]).mean() <= 2e-3
    assert ((env.nfev - 2) // 12 != (g["nfev"] - 2) // 12).mean() <= 0.03


@pytest.mark.parametrize("integ", ["solve_ivp", "euler"])
def test_kernel_quad_v0_base_env(libs, integ):
    K, _ = libs
    g = np.load(os.path.join(G, "quad_v0.npz"))
    s_in = g[integ + "_state_in"]; n = s_in.shape[0]
    env = HostEnv(K, _config(0, n_envs=n, integrator=1 if integ == "euler" else 0))
    env.set_state(s_in, np.zeros((n, 8)), g[integ + "_params"], g[integ + "_goal"])
    env.launch(np.ascontiguousarray(g[integ + "_action"], np.float64))
    assert np.abs(env.state.T - g[integ + "_state_out"]).max() < 1e-12
    assert (env.done.astype(bool) == g[integ + "_done"]).all()
    assert np.abs(env.reward - g[integ + "_reward"]).max() < 1e-12


def test_kernel_extreme_angular_rates_take_the_checked_redo(libs):
    """|W| far outside the termination limits: stage matrices leave SO(3), the speculative attempt is thrown away and
    redone with per-stage re-projection; results and RHS counts must equal the oracle's stage-by-stage evaluation."""
    import quad_oracle as qo
    K, _ = libs
    n = 128
    rng = np.random.default_rng(0)
    orc = qo.COracle("MONO")
    st, ig, par = orc.reset_from_uniforms(rng.random((n, 20)))
    st[:, 15:18] = rng.uniform(-1, 1, (n, 3)) * 60.0
    goal = np.zeros((n, 12)); goal[:, 6] = 1.0
    act = rng.uniform(-1, 1, (n, 4))
    env = HostEnv(K, _config(1, n_envs=n))
    env.set_state(st, ig, par, goal)
    env.launch(act)
    st_o, ig_o = st.copy(), ig.copy()
    obs_o, rew_o, done_o, nfev_o, status_o = orc.step(st_o, ig_o, par, goal, act)
    assert (env.nfev == nfev_o).all()
    assert _relerr(env.state.T, st_o) <= 1e-11
    assert env.stats[15] > 0          # SO(3) re-projections were needed


def _reset_all(F, cfg, env, episode=1):
    """k_reset + k_init_goal for every env, through the per-env functions of twin.cpp."""
    n = env.n
    st = np.empty((n, 18)); ig = np.empty((n, 8)); par = np.empty((n, 6)); gl = np.empty((n, 12))
    dp = C.POINTER(C.c_double)
    for i in range(n):
        F.tw_reset(C.byref(cfg), cfg.env_id_offset + i, episode, cfg.env_type, st[i].ctypes.data_as(dp), ig[i].ctypes.data_as(dp),
                   par[i].ctypes.data_as(dp), gl[i].ctypes.data_as(dp))
    env.set_state(st, ig, par, gl)
    env.ep_index[:] = episode
    return st, ig, par, gl


@pytest.mark.parametrize("dtype64,act64", [(True, True), (False, False), (False, True), (True, False)])
def test_kernel_rollout_with_autoreset_equals_single_steps(libs, dtype64, act64):
    """Multi-step launch (an env whose episode ends leaves its lane, is reset in a batch and waits in CQ for a free lane to
    go on stepping) == single-step launches (queued resets, batches of >= 24 and bursts of a common time limit), ragged n,
    misaligned rollout rows.  The float32 cases start from staggered episode lengths, so that lanes run out of phase and
    envs are adopted from CQ and from the tile sequence in the same round."""
    K, F = libs
    n, steps = 203, 20
    kw = dict(n_envs=n, seed=5, autoreset=1, goal_mode=1, max_episode_steps=7)
    c1, c2 = _config(1, dtype64, **kw), _config(1, dtype64, **kw)
    e1, e2 = HostEnv(K, c1, warps=2), HostEnv(K, c2, warps=5)
    _reset_all(F, c1, e1); _reset_all(F, c2, e2)
    rng = np.random.default_rng(15)
    if not dtype64:
        e1.ep_length[:] = rng.integers(0, 7, n); e2.ep_length[:] = e1.ep_length
    acts = rng.uniform(-1, 1, (steps, n, 4)).astype(np.float64 if act64 else np.float32)   # staged or loaded at the step start
    obs_r, rew_r, done_r = e1.launch(acts, n_steps=steps, store=True)
    for k in range(steps):
        e2.launch(acts[k])
        assert np.array_equal(e2.obs, obs_r[k]), k
        assert np.array_equal(e2.reward, rew_r[k]) and np.array_equal(e2.done, done_r[k])
    for name in ("state", "integ", "params", "goal", "obs", "ep_length", "ep_index"):
        assert np.array_equal(getattr(e1, name), getattr(e2, name)), name
    assert e1.stats[0] == e2.stats[0] >= 2 * n and e1.stats[7] == e2.stats[7] == steps * n


@pytest.mark.parametrize("mode", [1, 2])
def test_kernel_one_step_rollout_with_storage_equals_step(libs, mode):
    """qr_rollout(n_steps = 1) into the caller's arrays (served by the multi-step kernel: the single-step kernel carries no
    code for rollout storage) against a plain step: rows, rewards, dones, every env array, with resets in both."""
    K, F = libs
    n = 131
    kw = dict(n_envs=n, seed=8, autoreset=1, goal_mode=1, max_episode_steps=3, diagnostics=1)
    c1, c2 = _config(mode, True, **kw), _config(mode, True, **kw)
    e1, e2 = HostEnv(K, c1, warps=2), HostEnv(K, c2, warps=3)
    _reset_all(F, c1, e1); _reset_all(F, c2, e2)
    rng = np.random.default_rng(21)
    for t in range(5):
        act = rng.uniform(-1, 1, (n, e1.A))
        obs_r, rew_r, done_r = e1.launch(act[None], n_steps=1, store=True)
        e2.launch(act)
        assert np.array_equal(e1.obs, e2.obs) and np.array_equal(e1.final_obs, e2.final_obs), t
        assert np.array_equal(rew_r[0], e2.reward) and np.array_equal(done_r[0], e2.done)
        # the stored row is the step's own observation; the handle's row of an env that was reset is its new episode's first
        keep = ~(e2.terminated | e2.truncated).astype(bool)
        assert np.array_equal(obs_r[0][keep], e2.obs[keep])
        for name in ("state", "integ", "params", "goal", "reward", "done", "terminated", "truncated", "nfev", "ep_length", "ep_index", "ep_return"):
            assert np.array_equal(getattr(e1, name), getattr(e2, name)), (name, t)
    assert np.array_equal(e1.stats, e2.stats) and e1.stats[0] >= n


def test_kernel_autoreset_equals_manual_protocol(libs):
    """The kernel's auto reset against step -> reset -> goal -> get_norm_error_state done env by env with the per-env
    functions (main.py:212-230): states, goals, parameters, first observations of the new episodes, final_obs."""
    K, F = libs
    n, steps, limit = 96, 30, 9
    cfg = _config(1, n_envs=n, seed=3, autoreset=1, goal_mode=1, max_episode_steps=limit)
    ref_cfg = _config(1, n_envs=n, seed=3, autoreset=0, goal_mode=1)
    env = HostEnv(K, cfg, warps=1)
    st, ig, par, gl = _reset_all(F, cfg, env)
    episode = np.ones(n, np.int64); eplen = np.zeros(n, np.int64)
    rng = np.random.default_rng(9)
    dp, ip, fp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_float)
    n_resets = 0
    for t in range(steps):
        act = rng.uniform(-1, 1, (n, 4))
        env.launch(act)
        for i in range(n):
            so = np.empty(18); io = np.empty(8); ob = np.empty(23, np.float32); rw = np.empty(2); dn = np.empty(2, np.int32)
            nf = np.empty(1, np.int32); npj = np.empty(1, np.int32)
            F.tw_step(C.byref(ref_cfg), st[i].ctypes.data_as(dp), ig[i].ctypes.data_as(dp), par[i].ctypes.data_as(dp),
                      gl[i].ctypes.data_as(dp), act[i].ctypes.data_as(dp), 0, so.ctypes.data_as(dp), io.ctypes.data_as(dp),
                      ob.ctypes.data_as(fp), rw.ctypes.data_as(dp), dn.ctypes.data_as(ip), nf.ctypes.data_as(ip), npj.ctypes.data_as(ip))
            st[i], ig[i] = so, io
            eplen[i] += 1
            assert env.reward[i, 0] == rw[0] and bool(env.done[i, 0]) == bool(dn[0])
            if dn[0] or eplen[i] >= limit:
                assert np.array_equal(env.final_obs[i], ob)
                episode[i] += 1; eplen[i] = 0; n_resets += 1
                F.tw_reset(C.byref(ref_cfg), i, int(episode[i]), 0, st[i].ctypes.data_as(dp), ig[i].ctypes.data_as(dp),
                           par[i].ctypes.data_as(dp), gl[i].ctypes.data_as(dp))
                F.tw_norm_error_state(C.byref(ref_cfg), st[i].ctypes.data_as(dp), ig[i].ctypes.data_as(dp), gl[i].ctypes.data_as(dp),
                                      ob.ctypes.data_as(fp))
            assert np.array_equal(env.obs[i], ob), (t, i)
        assert np.array_equal(env.state.T, st) and np.array_equal(env.integ.T, ig), t
        assert np.array_equal(env.params.T, par) and np.array_equal(env.goal.T, gl), t
    assert n_resets > n and env.stats[0] == n_resets and env.stats[7] == steps * n


def test_kernel_inkernel_actions_and_sharding(libs):
    """Philox actions drawn inside the kernel (single- and multi-step launches) and independence of the split into
    handles: one 160-env handle against two 80-env shards with env_id_offset."""
    K, F = libs
    kw = dict(seed=21, autoreset=1, goal_mode=1, max_episode_steps=11)
    whole = HostEnv(K, _config(1, n_envs=160, **kw), warps=2)
    parts = [HostEnv(K, _config(1, n_envs=80, env_id_offset=80 * i, **kw), warps=3) for i in range(2)]
    for e in [whole] + parts:
        _reset_all(F, e.cfg, e)
        e.launch(None, n_steps=15)
        for _ in range(4):
            e.launch(None)
    for name in ("state", "integ", "params", "goal"):
        assert np.array_equal(np.concatenate([getattr(p, name) for p in parts], axis=1), getattr(whole, name)), name
    for name in ("obs", "reward", "done", "ep_length", "ep_index"):
        assert np.array_equal(np.concatenate([getattr(p, name) for p in parts], axis=0), getattr(whole, name)), name
    assert whole.stats[0] == sum(p.stats[0] for p in parts) >= 160 and whole.stats[7] == 19 * 160


@pytest.mark.parametrize("name,gm", [("hover", 2), ("circle", 3), ("eight", 4), ("circle_manual", 3), ("takeoff", 5), ("land", 6),
                                     ("land_low", 6), ("stay", 7)])
def test_kernel_trajectory_modes_match_reference(libs, name, gm):
    """test_gpu_parity.py::test_trajectory_modes_match_reference through the emulated k_init_goal / k_goal_update:
    goal modes QR_GOAL_TRAJ_HOVER ... QR_GOAL_TRAJ_STAY call by call against the reference's TrajectoryGenerator."""
    K, _ = libs
    g = np.load(os.path.join(G, "traj_modes.npz"))
    st, goal_ref, bdd_ref, t_ref = g[name + "_state"], g[name + "_goal"], g[name + "_b1d_dot"], g[name + "_t"]
    t_traj, w, smooth, theta0 = g[name + "_draws"]
    env = HostEnv(K, _config(1, n_envs=1, goal_mode=gm), warps=1)
    par = np.array([[2.15, 0.23, 0.022, 0.035, 0.0135, 2.2]])
    gl0 = np.zeros((1, 12)); gl0[0, 6] = 1.0
    env.set_state(st[0:1], np.zeros((1, 8)), par, gl0)
    env.companion("init_goal")                      # mark_traj_start + first get_desired on the float32 reset state
    ts = env.traj
    if name == "hover":                             # inject the reference run's two random draws
        ts[8, 0] = t_traj; ts[7, 0] = smooth; ts[6, 0] = w
    if name == "circle_manual":
        ts[8, 0] = 1.75                             # the golden run shortened the circle to reach manual mode
    if name != "hover":
        assert np.abs(env.goal[:, 0] - goal_ref[0]).max() < 1e-11
    worst = 0.0
    for i in range(1, len(t_ref)):
        env.state[:, 0] = st[i]
        env.companion("goal_update")
        worst = max(worst, np.abs(env.goal[:, 0] - goal_ref[i]).max(), np.abs(ts[9:11, 0] - bdd_ref[i, 0:2]).max())
        assert abs(ts[0, 0] - t_ref[i]) < 1e-12
    assert worst < 1e-11, worst
    assert bool(int(ts[1, 0]) & 2) == bool(g[name + "_manual"][-1])


def test_kernel_autoreset_equals_companion_kernels(libs):
    """test_gpu_parity.py::test_autoreset_equals_manual_protocol on the emulator: auto reset inside k_step against
    k_step -> k_reset(mask) -> k_init_goal(mask) -> k_norm_error_state(mask) (main.py:212-230)."""
    K, _ = libs
    n, limit = 70, 9
    kw = dict(n_envs=n, seed=3, goal_mode=1)
    a = HostEnv(K, _config(1, autoreset=1, max_episode_steps=limit, **kw), warps=2)
    b = HostEnv(K, _config(1, autoreset=0, **kw), warps=2)
    for e in (a, b):
        e.companion("reset"); e.companion("init_goal"); e.companion("norm_error_state")
    assert np.array_equal(a.state, b.state) and np.array_equal(a.obs, b.obs) and (a.ep_index == 1).all()
    steps_b = np.zeros(n, np.int64)
    rng = np.random.default_rng(9)
    n_resets = 0
    for t in range(28):
        act = rng.uniform(-1, 1, (n, 4))
        a.launch(act); b.launch(act)
        steps_b += 1
        assert np.array_equal(a.reward, b.reward) and np.array_equal(a.done, b.done)
        need = b.done[:, 0].astype(bool) | (steps_b >= limit)
        assert np.array_equal(a.terminated.astype(bool), b.done[:, 0].astype(bool)) and np.array_equal(a.truncated.astype(bool), steps_b >= limit)
        if need.any():
            term_obs = b.obs.copy()
            b.companion("reset", mask=need); b.companion("init_goal", mask=need); b.companion("norm_error_state", mask=need)
            steps_b[need] = 0
            assert np.array_equal(a.final_obs[need], term_obs[need])
            n_resets += int(need.sum())
        assert np.array_equal(a.obs, b.obs), t
        for name in ("state", "goal", "params", "integ"):
            assert np.array_equal(getattr(a, name), getattr(b, name)), (t, name)
    assert n_resets > n and a.stats[0] == n_resets


@pytest.mark.parametrize("fw,tag", [("MONO", "mono"), ("MODUL", "modul")])
def test_kernel_fused_policy_rollout_flies_the_reference_eval_episode(libs, fw, tag):
    """qr_rollout(act_dtype = QR_ACT_POLICY): obs -> shipped TD3 actor -> env.step fused in the kernel, against the
    evaluation episode the reference flew with the same checkpoint (tests/golden/eval_*.npz, KAT-2), and bit for bit
    against stepping with the same actor applied outside the kernel."""
    K, _ = libs
    ep = np.load(os.path.join(G, "eval_%s.npz" % tag))
    H = len(ep["reward"])
    mode = 1 if fw == "MONO" else 2
    n = 3
    def fresh():
        env = HostEnv(K, _config(mode, n_envs=n, goal_mode=1), warps=1)
        env.set_state(np.tile(ep["state0"], (n, 1)), np.tile(ep["integ0"], (n, 1)), np.tile(ep["params"], (n, 1)),
                      np.tile(ep["goal0"], (n, 1)))
        env.obs[:] = np.tile(ep["obs0"], (n, 1))
        return env
    fused = fresh()
    obs_r, rew_r, done_r = fused.launch("policy", n_steps=H, store=True)
    assert not done_r.any()
    assert np.abs(fused.state.T[0] - ep["state"][-1]).max() < 1e-3
    ret = rew_r[:, 0, :].sum(axis=0)
    assert np.abs(ret - ep["reward"].sum(axis=0)).max() < 0.05 and ret[0] > 985
    assert np.abs(fused.state.T - fused.state.T[0]).max() == 0.0          # identical envs stay identical
    # the same loop with the actor outside the kernel (first 60 steps): identical bits
    loop = fresh()
    act = np.empty((n, loop.A), np.float32)
    fp = C.POINTER(C.c_float)
    for t in range(60):
        for i in range(n):
            K.tw_actor(mode, loop.obs[i].ctypes.data_as(fp), act[i].ctypes.data_as(fp))
        loop.launch(act)
        assert np.array_equal(loop.obs, obs_r[t]) and np.array_equal(loop.reward, rew_r[t]), t
    # a full warp in lock step (the coalesced observation path feeds the actor) agrees with the 3-env run
    full = HostEnv(K, _config(mode, n_envs=32, goal_mode=1), warps=1)
    full.set_state(np.tile(ep["state0"], (32, 1)), np.tile(ep["integ0"], (32, 1)), np.tile(ep["params"], (32, 1)),
                   np.tile(ep["goal0"], (32, 1)))
    full.obs[:] = np.tile(ep["obs0"], (32, 1))
    obs_f, rew_f, _ = full.launch("policy", n_steps=120, store=True)
    assert np.array_equal(obs_f[:, :3], obs_r[:120]) and np.array_equal(obs_f[:, 31], obs_r[:120, 0])
    # single-step policy launches continue a fused rollout seamlessly
    a2, b2 = fresh(), fresh()
    a2.launch("policy", n_steps=7)
    for _ in range(7):
        b2.launch("policy", n_steps=1)
    assert np.array_equal(a2.state, b2.state) and np.array_equal(a2.obs, b2.obs) and np.array_equal(a2.integ, b2.integ)


def test_kernel_modul_quad_rollouts_and_external_goals(libs):
    """DecoupledWrapper and Quad-v0 through multi-step launches with EXTERNAL goals (the goal is read from the goal
    buffer at the end of every step) == single-step launches; float32; auto reset off and on."""
    K, F = libs
    rng = np.random.default_rng(3)
    for mode, A in ((2, 5), (0, 4)):
        for autoreset in (0, 1):
            n, steps = 77, 9
            kw = dict(n_envs=n, seed=8, autoreset=autoreset, goal_mode=0, max_episode_steps=4 if autoreset else 0)
            e1, e2 = HostEnv(K, _config(mode, False, **kw), warps=2), HostEnv(K, _config(mode, False, **kw), warps=3)
            for e in (e1, e2):
                e.companion("reset"); e.goal[:] = 0; e.goal[6] = 1; e.goal[0] = 0.1; e.goal[4] = -0.05
            acts = rng.uniform(-1, 1, (steps, n, A)).astype(np.float32)
            obs_r, rew_r, done_r = e1.launch(acts, n_steps=steps, store=True)
            for k in range(steps):
                e2.launch(acts[k])
                assert np.array_equal(e2.obs, obs_r[k]) and np.array_equal(e2.reward, rew_r[k]) and np.array_equal(e2.done, done_r[k]), (mode, k)
            for name in ("state", "integ", "params", "ep_length", "ep_index"):
                assert np.array_equal(getattr(e1, name), getattr(e2, name)), (mode, name)


def test_kernel_edge_cases(libs):
    """Non-finite state flagged (scipy would raise), tiny handles, and launches over sub-ranges of the envs (what
    qr_step_host's chunked pipeline issues) == one launch over all of them."""
    import quad_oracle as qo
    K, _ = libs
    rng = np.random.default_rng(0)
    n = 64
    st, ig, par = qo.COracle("MONO").reset_from_uniforms(rng.random((n, 20)))
    goal = np.zeros((n, 12)); goal[:, 6] = 1.0
    act = rng.uniform(-1, 1, (n, 4))
    bad = st.copy(); bad[5, 2] = np.nan
    env = HostEnv(K, _config(1, n_envs=n))
    env.set_state(bad, ig, par, goal)
    env.launch(act)
    assert env.status[5] & 1 and (np.delete(env.status, 5) == 0).all()
    assert np.isfinite(np.delete(env.state.T, 5, axis=0)).all() and env.stats[8] == 1
    # the other envs are what a clean launch gives
    ref = HostEnv(K, _config(1, n_envs=n)); ref.set_state(st, ig, par, goal); ref.launch(act)
    keep = np.arange(n) != 5
    assert np.array_equal(env.state[:, keep], ref.state[:, keep]) and np.array_equal(env.obs[keep], ref.obs[keep])
    # sub-range launches (ragged boundaries, not multiples of 32)
    parts = HostEnv(K, _config(1, n_envs=n)); parts.set_state(st, ig, par, goal)
    for lo, hi in ((0, 19), (19, 50), (50, 64)):
        parts.launch(act, lo=lo, hi=hi)
    for name in ("state", "integ", "obs", "reward", "done", "nfev"):
        assert np.array_equal(getattr(parts, name), getattr(ref, name)), name
    # tiny handles
    for m in (1, 5):
        tiny = HostEnv(K, _config(1, n_envs=m), warps=1); tiny.set_state(st[:m], ig[:m], par[:m], goal[:m]); tiny.launch(act[:m])
        assert np.array_equal(tiny.state, ref.state[:, :m]) and np.array_equal(tiny.obs, ref.obs[:m])


def test_kernel_fp64_free_running_vs_c_oracle(libs):
    """State fed back for 300 steps without re-synchronisation (the north-star bar: <= 1e-9 after the horizon on envs
    still alive), attempt counts and done flags identical step by step."""
    import quad_oracle as qo
    K, _ = libs
    n = 48
    rng = np.random.default_rng(4)
    orc = qo.COracle("MONO")
    st_o, ig_o, par = orc.reset_from_uniforms(rng.random((n, 20)))
    goal = np.zeros((n, 12)); goal[:, 6] = 1.0
    env = HostEnv(K, _config(1, n_envs=n), warps=2)
    env.set_state(st_o, ig_o, par, goal)
    alive = np.ones(n, bool)
    for t in range(300):
        act = rng.uniform(-1, 1, (n, 4)) * 0.3
        obs_o, rew_o, done_o, nfev_o, _ = orc.step(st_o, ig_o, par, goal, act)
        env.launch(act)
        d = np.asarray(done_o).reshape(n, -1)[:, 0].astype(bool)
        assert (env.nfev[alive] == nfev_o[alive]).all() and (env.done[alive, 0].astype(bool) == d[alive]).all(), t
        alive &= ~d
        if not alive.any():
            break
        assert _relerr(env.state.T[alive], st_o[alive]) <= 1e-9, t
    assert t > 50


@pytest.mark.parametrize("gm", [2, 3, 4, 5])
def test_kernel_step_evaluates_the_trajectory_goal_itself(libs, gm):
    """Goal modes hover / circle / eight / take-off: the step kernel calls get_desired + set_goal_state on the pre-step state
    before every env.step (main.py:145-147).  Single-step and multi-step launches == k_goal_update followed by a step that
    takes the goal as an external one, step by step: state, integrals, goal, trajectory state, observations, rewards."""
    K, F = libs
    n, steps = 70, 6
    kw = dict(n_envs=n, seed=11, goal_mode=gm)
    ea, eb, ec = (HostEnv(K, _config(1, True, **kw), warps=w) for w in (2, 3, 1))
    for e in (ea, eb, ec):
        _reset_all(F, e.cfg, e)
        e.companion("init_goal")
    rng = np.random.default_rng(2)
    acts = rng.uniform(-0.3, 0.3, (steps, n, 4))
    obs_r, rew_r, _ = ec.launch(acts, n_steps=steps, store=True)       # goal generated inside a multi-step launch
    for k in range(steps):
        ea.launch(acts[k])                                             # ... inside single-step launches
        eb.companion("goal_update")                                    # ... by the companion kernel, then an external-goal step
        eb.cfg.goal_mode = 0
        eb.launch(acts[k])
        eb.cfg.goal_mode = gm
        for name in ("state", "integ", "goal", "traj", "obs", "reward", "done"):
            assert np.array_equal(getattr(ea, name), getattr(eb, name)), (k, name)
        assert np.array_equal(obs_r[k], ea.obs) and np.array_equal(rew_r[k], ea.reward), k
    for name in ("state", "integ", "goal", "traj"):
        assert np.array_equal(getattr(ec, name), getattr(ea, name)), name
    assert np.abs(ea.goal[0:3]).max() > 0 and ea.traj[0].min() > 0     # a moving position command, clocks advanced


def test_kernel_benchmark_reward_solved_count_and_rounded_returns(libs):
    """Statistics 16 / 17 (sum of benchmark_reward_func over env-steps, utils/utils.py:21-47; episodes the trainer relabels
    as solved at the time limit, main.py:169-173) and the trainer's 4-decimal running return (main.py:180, round_returns)."""
    K, F = libs
    n, steps = 64, 3
    cfg = _config(1, True, n_envs=n, seed=4, goal_mode=0, max_episode_steps=steps, round_returns=1)
    env = HostEnv(K, cfg, warps=2)
    st, ig, par, gl = _reset_all(F, cfg, env)
    st[: n // 2, 0:6] = 0.0                                  # half of the envs sit at the goal: x = xd = 0, v = 0
    st[: n // 2, 6:15] = np.eye(3).reshape(-1); st[: n // 2, 15:18] = 0.0
    gl[:] = 0.0; gl[:, 6] = 1.0
    env.set_state(st, ig, par, gl)
    rng = np.random.default_rng(6)
    br_sum, ret = 0.0, np.zeros(n)
    for k in range(steps):
        act = rng.uniform(-0.05, 0.05, (n, 4)); act[:, 0] = -0.06
        env.launch(act)
        o = env.obs.astype(np.float64)
        r = -np.linalg.norm(o[:, 0:3] * 1.0, axis=1) - np.abs(o[:, 18] * np.pi)
        br_sum += np.interp(r, [-2.0, 0.0], [0.0, 1.0]).sum()
        ret = np.array([float("{:.4f}".format(a + b)) for a, b in zip(ret, env.reward[:, 0])])
        assert np.array_equal(env.ep_return[0], ret), k
    assert abs(env.stats[16] - br_sum) < 1e-4 * br_sum
    solved = (np.abs(env.obs[:, 0:3].astype(np.float64)) <= 0.03).all(axis=1) & (env.reward[:, 0] != -1.0)
    assert env.truncated.all() and env.stats[17] == solved.sum() and 0 < solved.sum() < n
