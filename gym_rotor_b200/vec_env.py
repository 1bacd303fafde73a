"""Batched, device-resident mirror of gym-rotor's env objects.

`BatchedQuadEnv` keeps the reference's duck-typed surface -- reset / step / get_current_state /
set_goal_state / get_norm_error_state and the attributes the trainer reads (main.py:68-73,126-129,145-147,
164,226-230; policy_regularization.py:31-33) -- but every method works on N envs whose state lives on the
GPU.  All compute happens in the sm_100a library behind include/quadrotor_b200.h; torch is used for
device tensors and streams only.  There is no CPU path.

Shapes: obs MONO [N,23] / MODUL ([N,15],[N,3]) float32; reward, done [N,G]; state [N,18].
"""
import ctypes as C

import numpy as np
import torch

from . import _native as nat

_FRAMEWORKS = {"QUAD": nat.MODE_QUAD, "MONO": nat.MODE_COUPLED, "MODUL": nat.MODE_DECOUPLED}


class _DevArray:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def _view(ptr, shape, typestr, device):
    return torch.as_tensor(_DevArray(ptr, shape, typestr), device=device)


class Box:
    """Stand-in for gymnasium.spaces.Box with what the reference's callers use (quad.py:120-132; utils/utils.py:17-18
    seeds both spaces): low / high / shape / dtype, seed(), sample(), contains().  gymnasium is not a dependency."""

    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is None:
            shape = np.shape(low)
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self.low = np.broadcast_to(np.asarray(low, self.dtype), self.shape).copy()
        self.high = np.broadcast_to(np.asarray(high, self.dtype), self.shape).copy()
        self._rng = np.random.default_rng()

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)
        return [seed]

    def sample(self):
        return self._rng.uniform(self.low, self.high).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))


def forces_to_fM_matrices(d, c_tf):
    """forces_to_fM of quad.py:396-401 for per-env arm lengths d [N] and torque coefficients c_tf [N] -> [N,4,4]
    (rows f, M1, M2, M3; columns T1..T4), and its inverse fM_to_forces (quad.py:402)."""
    d = torch.as_tensor(d, dtype=torch.float64).reshape(-1)
    c = torch.as_tensor(c_tf, dtype=torch.float64, device=d.device).reshape(-1)
    one, zero = torch.ones_like(d), torch.zeros_like(d)
    A = torch.stack([torch.stack([one, one, one, one], dim=1),
                     torch.stack([zero, -d, zero, d], dim=1),
                     torch.stack([d, zero, -d, zero], dim=1),
                     torch.stack([-c, c, -c, c], dim=1)], dim=1)
    return A, torch.linalg.inv(A)


class BatchedQuadEnv:
    """N quadrotor envs on one GPU.  `framework`: 'MONO' (CoupledWrapper), 'MODUL' (DecoupledWrapper), 'QUAD' (Quad-v0)."""

    def __init__(self, num_envs, framework="MONO", dtype=torch.float32, device="cuda:0", seed=1992, autoreset=False,
                 goal_mode="external", env_type="train", max_episode_steps=0, env_id_offset=0, diagnostics=True,
                 integrator="solve_ivp", **overrides):
        if not torch.cuda.is_available():
            raise nat.NativeError("BatchedQuadEnv needs a CUDA device: the simulator has no CPU path")
        self._L = nat.load()
        self.framework = framework
        self.device = torch.device(device)
        self._dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.dtype = dtype
        cfg = nat.QrConfig()
        nat.check(self._L.qr_default_config(C.byref(cfg), _FRAMEWORKS[framework], nat.F64 if dtype == torch.float64 else nat.F32))
        cfg.n_envs = int(num_envs); cfg.env_id_offset = int(env_id_offset); cfg.seed = int(seed)
        cfg.autoreset = int(bool(autoreset))
        cfg.goal_mode = {"external": nat.GOAL_EXTERNAL, "traj0": nat.GOAL_TRAJ_MODE0, "hover": nat.GOAL_TRAJ_HOVER,
                         "circle": nat.GOAL_TRAJ_CIRCLE, "eight": nat.GOAL_TRAJ_EIGHT, "takeoff": nat.GOAL_TRAJ_TAKEOFF,
                         "land": nat.GOAL_TRAJ_LAND, "stay": nat.GOAL_TRAJ_STAY}[goal_mode]
        cfg.env_type = nat.ENV_TRAIN if env_type == "train" else nat.ENV_EVAL
        cfg.max_episode_steps = int(max_episode_steps)
        cfg.diagnostics = int(bool(diagnostics))
        cfg.integrator = nat.INT_EULER if integrator == "euler" else nat.INT_DOP853
        for k, v in overrides.items():
            if not hasattr(cfg, k):
                raise TypeError("unknown config field %r" % k)
            setattr(cfg, k, v)
        # values the reference derives from the coefficients (quad.py:80-88, decoupled_yaw_wrapper.py:29-33) follow the
        # overrides unless they were overridden themselves
        if "CW" not in overrides:
            cfg.CW = cfg.Cw12
        if "reward_min" not in overrides:
            cfg.reward_min = -np.ceil(cfg.Cx + cfg.CIx + cfg.Cv + cfg.Cb1 + cfg.CIb1 + cfg.CW)
        if "reward_min_1" not in overrides:
            cfg.reward_min_1 = -np.ceil(cfg.Cx + cfg.CIx + cfg.Cv + cfg.Cw12)
        if "reward_min_2" not in overrides:
            cfg.reward_min_2 = -np.ceil(cfg.Cb1 + cfg.CW3 + cfg.CIb1)
        self.cfg = cfg
        self._h = C.c_void_p()
        nat.check(self._L.qr_create(C.byref(cfg), self._dev_index, C.byref(self._h)))
        b = nat.QrBuffers()
        nat.check(self._L.qr_get_buffers(self._h, C.byref(b)))
        self._b = b
        N, O, G = int(num_envs), b.obs_dim, b.n_agents
        self.num_envs, self.obs_dim, self.act_dim, self.n_agents = N, O, b.act_dim, G
        ts = "<f8" if dtype == torch.float64 else "<f4"
        d = self.device
        # zero-copy views of the library-owned buffers ([component][env] for the state-side arrays)
        self.state_soa = _view(b.state, (18, N), ts, d)
        self.integ_soa = _view(b.integ, (8, N), ts, d)
        self.params_soa = _view(b.params, (6, N), ts, d)
        self.goal_soa = _view(b.goal, (12, N), ts, d)
        self.traj_soa = _view(b.traj, (12, N), ts, d)
        S = int(self._L.qr_obs_stride(self._h))   # row stride of the observation buffers (O unless built with padded rows)
        self.obs = _view(b.obs, (N, S), "<f4", d)[:, :O]
        self.reward = _view(b.reward, (N, G), ts, d)
        self.done = _view(b.done, (N, G), "|u1", d)
        self.terminated = _view(b.terminated, (N,), "|u1", d)
        self.truncated = _view(b.truncated, (N,), "|u1", d)
        self.final_obs = _view(b.final_obs, (N, S), "<f4", d)[:, :O]
        self.nfev = _view(b.nfev, (N,), "<i4", d)
        self.status = _view(b.status, (N,), "|u1", d)
        self.ep_return = _view(b.ep_return, (2, N), ts, d)
        self.ep_length = _view(b.ep_length, (N,), "<i4", d)
        self.ep_index = _view(b.ep_index, (N,), "<u4", d) if hasattr(torch, "uint32") else None
        self.stats_dev = _view(b.stats, (nat.NUM_STATS,), "<f8", d)
        # constants the reference exposes as attributes
        self.dt, self.g = cfg.dt, cfg.g
        self.x_lim, self.v_lim, self.W_lim = cfg.x_lim, cfg.v_lim, cfg.W_lim
        self.eIx_lim, self.eIb1_lim = cfg.eIx_lim, cfg.eIb1_lim
        self.min_force = cfg.min_force
        self.alpha, self.beta = cfg.alpha, cfg.beta
        # spaces as the reference declares them (quad.py:104-132): raw-state bounds, normalised actions
        hi = np.concatenate([cfg.x_lim * np.ones(3), cfg.v_lim * np.ones(3), np.ones(9), cfg.W_lim * np.ones(3)]).astype(np.float32)
        self.observation_space = Box(-hi, hi, dtype=np.float32)
        self.action_space = Box(-1.0, 1.0, shape=(self.act_dim,), dtype=np.float32)

    # ---- reference attribute surface (per-env where domain randomisation makes them per-env) ----
    J_nominal = np.diag([0.022, 0.022, 0.035])   # quad.py:30

    @property
    def forces_to_fM(self):
        """[N,4,4] float64: quad.py:396-401 with each env's (randomised) arm length and torque coefficient."""
        return forces_to_fM_matrices(self.params_soa[1], self.params_soa[4])[0]

    @property
    def fM_to_forces(self):
        """[N,4,4] float64: inverse of forces_to_fM (quad.py:402; read by draw_plot.py:63,70)."""
        return forces_to_fM_matrices(self.params_soa[1], self.params_soa[4])[1]

    @property
    def hover_force(self):
        return self.params_soa[0] * self.g / 4.0

    @property
    def max_force(self):
        return self.params_soa[5] * self.hover_force

    @property
    def avrg_act(self):
        return (self.min_force + self.max_force) / 2.0

    @property
    def scale_act(self):
        return self.max_force - self.avrg_act

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _mask_ptr(self, mask):
        if mask is None:
            return None, None
        m = torch.as_tensor(mask).to(device=self.device, dtype=torch.uint8).contiguous()   # the kernels read it on the device
        if tuple(m.shape) != (self.num_envs,):
            raise ValueError("mask must have shape (%d,)" % self.num_envs)
        return m, C.c_void_p(m.data_ptr())

    # ---- reference methods, batched ----
    def reset(self, env_type="train", seed=None, options=None, mask=None):
        """env.reset(env_type) for all envs (or those in `mask`); returns state [N,18] float32 like the reference."""
        keep, p = self._mask_ptr(mask)
        nat.check(self._L.qr_reset(self._h, p, nat.ENV_TRAIN if env_type == "train" else nat.ENV_EVAL, self._stream()))
        return self.state_soa.t().to(torch.float32)

    def init_goal(self, mask=None):
        """trajectory_generator mark_traj_start + get_desired(mode 0) (goal_mode='traj0')."""
        keep, p = self._mask_ptr(mask)
        nat.check(self._L.qr_init_goal(self._h, p, self._stream()))

    def goal_update(self):
        """ONE trajectory_generator.get_desired + set_goal_state call without stepping (hover ... stay goal modes; no-op
        otherwise).  step() / rollout() / step_host() make this call themselves before every env.step."""
        nat.check(self._L.qr_goal_update(self._h, self._stream()))

    def get_current_state(self):
        """Live [N,18] view of the device state (the reference returns a live alias too, quad.py:409-410)."""
        return self.state_soa.t()

    def set_goal_state(self, xd, vd, b1d, b1d_dot, Wd):
        g = self.goal_soa
        for k, v in ((0, xd), (3, vd), (6, b1d), (9, Wd)):
            v = torch.as_tensor(v, dtype=self.dtype, device=self.device)
            g[k:k + 3] = (v.t() if v.dim() == 2 else v.reshape(3, 1).expand(3, self.num_envs))

    def get_norm_error_state(self, framework=None, mask=None):
        """Mutating, like the reference: advances the integral terms once (quad.py:447-450)."""
        keep, p = self._mask_ptr(mask)
        nat.check(self._L.qr_norm_error_state(self._h, p, self._stream()))
        return self._split_obs(self.obs)

    def _split_obs(self, o):
        if self.framework == "MODUL":
            return [o[:, :15], o[:, 15:18]]
        return [o]

    def step(self, action):
        """env.step(action): action [N,A] float32/float64 CUDA tensor (a list of per-agent tensors is concatenated)."""
        if isinstance(action, (list, tuple)):
            action = torch.cat([a.reshape(self.num_envs, -1) for a in action], dim=1)
        if action.dtype not in (torch.float32, torch.float64):
            action = action.to(torch.float32)
        if action.device != self.device or not action.is_contiguous():
            action = action.to(self.device).contiguous()
        if tuple(action.shape) != (self.num_envs, self.act_dim):
            raise ValueError("action must have shape (%d, %d)" % (self.num_envs, self.act_dim))
        # (goal modes hover ... stay: the step kernel evaluates trajectory_generator.get_desired on the pre-step state
        #  itself, where the trainer calls it, main.py:145-147)
        nat.check(self._L.qr_step(self._h, C.c_void_p(action.data_ptr()),
                                  nat.F64 if action.dtype == torch.float64 else nat.F32, self._stream()))
        return self._split_obs(self.obs), self.reward, self.done.view(torch.bool), False, {}   # flags are 0/1 bytes: reinterpreted, not copied

    def policy_td3(self, out=None):
        """Actions of the reference's shipped TD3 actor(s) for the current observations, computed on device
        (compiled effective weights; agent.choose_action(obs, explor_noise_std=0), td3.py:93-96)."""
        if out is None:
            out = torch.empty((self.num_envs, self.act_dim), dtype=torch.float32, device=self.device)
        nat.check(self._L.qr_policy_td3(self._h, C.c_void_p(out.data_ptr()), self._stream()))
        return out

    def rollout(self, n_steps, actions=None, store=False):
        """n_steps fused env.step() calls in ONE kernel launch (state stays in registers).

        actions: [n_steps,N,A] tensor, None for in-kernel Philox U(-1,1) actions, or "policy": the reference's
        shipped TD3 actor evaluated inside the kernel on each env's latest observation (the evaluation loop
        obs -> choose_action -> step of main.py:304-365 in one launch; self.obs must be current).
        store=True returns per-step (obs, reward, done) tensors [n_steps,N,..]."""
        ap, ad = None, nat.F32
        if isinstance(actions, str):
            if actions != "policy":
                raise ValueError("actions must be a tensor, None or 'policy'")
            ad = nat.ACT_POLICY
        elif actions is not None:
            actions = actions.to(self.device).contiguous()
            assert tuple(actions.shape) == (n_steps, self.num_envs, self.act_dim)
            ap = C.c_void_p(actions.data_ptr()); ad = nat.F64 if actions.dtype == torch.float64 else nat.F32
        obs = rew = dn = None
        po = pr = pd = None
        if store:
            obs = torch.empty((n_steps, self.num_envs, self.obs_dim), dtype=torch.float32, device=self.device)
            rew = torch.empty((n_steps, self.num_envs, self.n_agents), dtype=self.dtype, device=self.device)
            dn = torch.empty((n_steps, self.num_envs, self.n_agents), dtype=torch.uint8, device=self.device)
            po, pr, pd = C.c_void_p(obs.data_ptr()), C.c_void_p(rew.data_ptr()), C.c_void_p(dn.data_ptr())
        nat.check(self._L.qr_rollout(self._h, int(n_steps), ap, ad, po, pr, pd, self._stream()))
        return obs, rew, dn

    def step_host(self, actions, obs_out=None, reward_out=None, done_out=None):
        """End-to-end step with HOST buffers (numpy arrays or pinned CPU tensors): H2D + step + D2H, ordered after the work
        enqueued on the current torch stream.  Buffers must be C-contiguous host memory of exactly the shapes / dtypes
        actions [N,A] float32|float64, obs [N,O] float32, reward [N,G] float32|float64 (the env's dtype), done [N,G] uint8|bool."""
        N = self.num_envs
        rdt = np.float64 if self.dtype == torch.float64 else np.float32

        def ptr(x, name, shape, dtypes):
            if x is None:
                return None
            if isinstance(x, torch.Tensor):
                if x.device.type != "cpu":
                    raise ValueError("step_host: %s must be host memory (use step() for device tensors)" % name)
                ok, dt, p = x.is_contiguous(), np.dtype(str(x.dtype).replace("torch.", "")), x.data_ptr()
            else:
                x = np.asarray(x) if not isinstance(x, np.ndarray) else x
                ok, dt, p = x.flags["C_CONTIGUOUS"], x.dtype, x.ctypes.data
            if tuple(x.shape) != shape or not ok or dt not in [np.dtype(d) for d in dtypes]:
                raise ValueError("step_host: %s must be C-contiguous, shape %s, dtype in %s" % (name, shape, [np.dtype(d).name for d in dtypes]))
            return C.c_void_p(p)
        pa = ptr(actions, "actions", (N, self.act_dim), (np.float32, np.float64))
        if pa is None:
            raise ValueError("step_host: actions are required")
        is64 = str(actions.dtype).endswith("float64")
        nat.check(self._L.qr_step_host(self._h, pa, nat.F64 if is64 else nat.F32,
                                       ptr(obs_out, "obs_out", (N, self.obs_dim), (np.float32,)),
                                       ptr(reward_out, "reward_out", (N, self.n_agents), (rdt,)),
                                       ptr(done_out, "done_out", (N, self.n_agents), (np.uint8, np.bool_)), self._stream()))

    # ---- state injection / extraction (row-major [N,..] float64 host arrays) ----
    def set_state(self, state=None, integ=None, params=None, goal=None):
        def arr(a, c):
            if a is None:
                return None, None
            a = np.ascontiguousarray(a, dtype=np.float64)
            assert a.shape == (self.num_envs, c), a.shape
            return a, C.c_void_p(a.ctypes.data)
        ks, ps = arr(state, 18); ki, pi = arr(integ, 8); kp, pp = arr(params, 6); kg, pg = arr(goal, 12)
        torch.cuda.synchronize(self.device)
        nat.check(self._L.qr_set_state_host(self._h, ps, pi, pp, pg))

    def get_state(self):
        N = self.num_envs
        st, ig, pa, gl = (np.empty((N, c), np.float64) for c in (18, 8, 6, 12))
        torch.cuda.synchronize(self.device)
        nat.check(self._L.qr_get_state_host(self._h, *(C.c_void_p(a.ctypes.data) for a in (st, ig, pa, gl))))
        return st, ig, pa, gl

    def stats(self, reset=True):
        out = (C.c_double * nat.NUM_STATS)()
        nat.check(self._L.qr_stats(self._h, out, int(reset), self._stream()))
        return np.array(out[:], dtype=np.float64)

    @staticmethod
    def launch_count():
        return int(nat.load().qr_launch_count())

    def render(self):
        raise NotImplementedError("render() (vpython viewer, quad.py:469-754) is out of scope")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            torch.cuda.synchronize(self.device)
            self._L.qr_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CoupledWrapper(BatchedQuadEnv):
    """Batched CoupledWrapper (monolithic agent: obs 23, act 4) -- wrappers/coupled_yaw_wrapper.py."""

    def __init__(self, num_envs=1, **kw):
        super().__init__(num_envs, framework="MONO", **kw)


class DecoupledWrapper(BatchedQuadEnv):
    """Batched DecoupledWrapper (two agents: obs 15+3, act 4+1) -- wrappers/decoupled_yaw_wrapper.py."""

    def __init__(self, num_envs=1, **kw):
        super().__init__(num_envs, framework="MODUL", **kw)


class QuadEnv(BatchedQuadEnv):
    """Batched base Quad-v0 env (T1..T4 actions, obs = state) -- envs/quad.py."""

    def __init__(self, num_envs=1, **kw):
        super().__init__(num_envs, framework="QUAD", **kw)


def benchmark_reward(obs_n, framework, x_lim=1.0):
    """utils/utils.py:21-47 batched: interp(-||ex|| - |eb1|, [-2, 0], [0, 1]) with ex = obs[0:3] * x_lim and
    eb1 = obs[18] * pi (MONO) or obs2[0] * pi (MODUL).  obs_n: list of observation tensors as step() returns them."""
    o = obs_n[0]
    ex = o[:, 0:3].double() * x_lim
    eb1 = (obs_n[1][:, 0] if framework == "MODUL" else o[:, 18]).double() * np.pi
    r = -torch.linalg.vector_norm(ex, dim=1) - eb1.abs()
    return torch.clamp((r + 2.0) / 2.0, 0.0, 1.0)


def time_limit_relabel(obs_n, reward, done, framework, x_lim=1.0):
    """The trainer's relabel when an episode hits max_steps (main.py:169-173): done_n[0] becomes "solved" =
    all(|ex| <= 0.03 m) and reward != -1; for MODUL done_n[1] = |eb1| <= 0.03 rad and reward != -1.
    Returns the relabelled done tensor [N, G] (to be applied to the truncated envs only)."""
    o = obs_n[0]
    ex = o[:, 0:3].double() * x_lim
    out = done.clone().bool()
    out[:, 0] = (ex.abs() <= 0.03).all(dim=1) & (reward[:, 0] != -1.0)
    if framework == "MODUL":
        eb1 = obs_n[1][:, 0].double() * np.pi
        out[:, 1] = (eb1.abs() <= 0.03) & (reward[:, 1] != -1.0)
    return out


def forces_from_fM(f, M, d=0.23, c_tf=0.0135, min_force=None, max_force=None):
    """(f, M) -> rotor thrusts T1..T4 = fM_to_forces @ [f; M] (the inverse of quad.py:396-401), optionally clipped
    to [min_force, max_force] as draw_plot.py:55-72 does.  Diagnostic output only: the wrappers' dynamics never
    clip individual rotor thrusts (coupled_yaw_wrapper.py:44-53)."""
    f = torch.as_tensor(f, dtype=torch.float64); M = torch.as_tensor(M, dtype=torch.float64)
    A = torch.tensor([[1.0, 1.0, 1.0, 1.0], [0.0, -d, 0.0, d], [d, 0.0, -d, 0.0], [-c_tf, c_tf, -c_tf, c_tf]],
                     dtype=torch.float64, device=f.device)
    fM = torch.cat([f.reshape(-1, 1), M.reshape(-1, 3)], dim=1)
    T = fM @ torch.linalg.inv(A).T
    if min_force is not None:
        T = T.clamp(min=min_force)
    if max_force is not None:
        T = torch.minimum(T, torch.as_tensor(max_force, dtype=torch.float64, device=T.device).reshape(-1, 1))
    return T


class FlightLog:
    """The evaluation flight log of main.py:344-352,382-389 for ONE env of a batch: per step a row
    `action | state[18] | eIx[3] | eb1 | eIb1 | xd[3] | vd[3] | b1d[3] | Wd[3]` (39 columns MONO, 40 MODUL), written with
    %.10f under the reference's two header lines, so that draw_plot.py reads it unchanged."""

    def __init__(self, env, index=0):
        self.env, self.i, self.rows = env, index, []

    def record(self, action, obs_n):
        e, i = self.env, self.i
        o = [t[i].double().cpu().numpy() for t in obs_n]
        if e.framework == "MODUL":
            eIx, eb1, eIb1 = o[0][3:6] * e.eIx_lim, o[1][0] * np.pi, o[1][1] * e.eIb1_lim
        else:
            eIx, eb1, eIb1 = o[0][3:6] * e.eIx_lim, o[0][18] * np.pi, o[0][19] * e.eIb1_lim
        st = e.state_soa[:, i].double().cpu().numpy()
        gl = e.goal_soa[:, i].double().cpu().numpy()
        act = action[i].double().cpu().numpy().ravel()
        self.rows.append(np.concatenate([act, st, eIx, [eb1], [eIb1], gl]))

    def save(self, path):
        with open(path, "w") as f:
            f.write("# Actions and States\n# action[0], ..., state[0], ..., command[0], ...\n")
            for r in self.rows:
                f.write(" ".join("%.10f" % v for v in r) + "\n")


def _gymnasium():
    """gymnasium if it is importable (it is not installed in the build image; the reference depends on it), else None."""
    try:
        import gymnasium
        import gymnasium.vector  # noqa: F401
        return gymnasium
    except Exception:
        return None


class TupleSpace:
    """Stand-in for gymnasium.spaces.Tuple (used only when gymnasium is absent)."""

    def __init__(self, spaces):
        self.spaces = tuple(spaces)

    def __len__(self):
        return len(self.spaces)

    def __getitem__(self, i):
        return self.spaces[i]

    def seed(self, seed=None):
        return [sp.seed(seed) for sp in self.spaces]

    def sample(self):
        return tuple(sp.sample() for sp in self.spaces)


def make_spaces(framework, num_envs, cfg):
    """(single_observation_space, single_action_space, observation_space, action_space) of a vector env.

    Quad-v0: exactly the boxes the reference declares (quad.py:104-132: raw-state bounds, actions in [-1, 1]^4).  The wrappers
    inherit that declaration in the reference although they return NORMALISED observations (23 values, or 15 + 3 for the
    two agents); a vector env has to describe what step() returns, so for them the observation space is the normalised box
    [-1, 1]^O (errors are divided by their limits, quad.py:421-466; R and b3 entries are direction cosines), a Tuple of the
    two agents' boxes for DecoupledWrapper.  gymnasium's classes are used when gymnasium is importable."""
    gym = _gymnasium()
    BoxT = gym.spaces.Box if gym else Box
    TupT = gym.spaces.Tuple if gym else TupleSpace
    A = 5 if framework == "MODUL" else 4

    def build(n):
        lead = () if n is None else (n,)
        if framework == "QUAD":
            hi = np.concatenate([cfg.x_lim * np.ones(3), cfg.v_lim * np.ones(3), np.ones(9), cfg.W_lim * np.ones(3)]).astype(np.float32)
            hi = np.broadcast_to(hi, lead + (18,)).copy()
            obs = BoxT(low=-hi, high=hi, dtype=np.float32)
        elif framework == "MONO":
            obs = BoxT(low=-1.0, high=1.0, shape=lead + (23,), dtype=np.float32)
        else:
            obs = TupT((BoxT(low=-1.0, high=1.0, shape=lead + (15,), dtype=np.float32),
                        BoxT(low=-1.0, high=1.0, shape=lead + (3,), dtype=np.float32)))
        act = BoxT(low=-1.0, high=1.0, shape=lead + (A,), dtype=np.float32)
        return obs, act
    so, sa = build(None)
    bo, ba = build(num_envs)
    return so, sa, bo, ba


_VecBase = (_gymnasium().vector.VectorEnv if _gymnasium() else object)


class QuadVectorEnv(_VecBase):
    """gymnasium.vector.VectorEnv over the device-resident batch (a subclass of it whenever gymnasium is importable; the
    build image does not have it, so there it is a plain class with the same attributes and methods).

    reset() -> (obs, info); step(actions) -> (obs, reward, terminated, truncated, info) with same-step auto reset done
    in-kernel (gymnasium's AutoresetMode.SAME_STEP): the returned obs of a finished env is the first observation of its
    next episode (main.py:226-230) and info['final_obs'] holds the terminal one.  Tensors stay on the device.
    Attributes: num_envs, single_observation_space, single_action_space, observation_space, action_space (make_spaces),
    spec-like `max_episode_steps`.
    """

    metadata = {"render_modes": [], "autoreset_mode": "SameStep"}

    def __init__(self, num_envs, framework="MONO", max_episode_steps=4000, goal_mode="traj0", **kw):
        self.env = BatchedQuadEnv(num_envs, framework=framework, autoreset=True,
                                  goal_mode=(goal_mode if framework != "QUAD" else "external"),
                                  max_episode_steps=max_episode_steps, **kw)
        self.num_envs = num_envs
        self.framework = framework
        self.max_episode_steps = max_episode_steps
        (self.single_observation_space, self.single_action_space,
         self.observation_space, self.action_space) = make_spaces(framework, num_envs, self.env.cfg)
        self.single_observation_shape = (self.env.obs_dim,)
        self.single_action_shape = (self.env.act_dim,)
        self.is_vector_env = True
        self.render_mode = None
        self.closed = False

    def reset(self, *, seed=None, options=None):
        e = self.env
        e.reset(env_type="train" if e.cfg.env_type == nat.ENV_TRAIN else "eval")
        if e.cfg.goal_mode != nat.GOAL_EXTERNAL:
            e.init_goal()
        obs = e.get_norm_error_state()
        return obs[0] if len(obs) == 1 else obs, {}

    def step(self, actions):
        e = self.env
        obs, rew, done, _, _ = e.step(actions)
        trunc = e.truncated.bool()
        info = {"final_obs": e.final_obs, "done_n": done}
        if bool(trunc.any()):   # main.py:169-173: what the trainer stores as done_n for episodes that hit the limit
            fin = e._split_obs(e.final_obs)
            info["time_limit_done_n"] = time_limit_relabel(fin, rew, done, e.framework, e.x_lim)
        return (obs[0] if len(obs) == 1 else obs), rew, e.terminated.bool(), e.truncated.bool(), info

    def close(self, **kwargs):
        if not getattr(self, "closed", True):
            self.closed = True
            self.env.close()

    def close_extras(self, **kwargs):
        pass


def register_envs():
    """gym_rotor/__init__.py:3-7 registers 'Quad-v0' (entry point QuadEnv, max_episode_steps 10000).  With gymnasium
    importable this registers the batched counterparts under the same naming: 'Quad-v0' is left to the reference package;
    'QuadB200-v0', 'CoupledWrapperB200-v0', 'DecoupledWrapperB200-v0' are created with
    gymnasium.make_vec(id, num_envs=N, vectorization_mode='vector_entry_point').  Returns the registered ids
    ([] without gymnasium)."""
    gym = _gymnasium()
    if gym is None:
        return []
    ids = []
    for name, fw, steps in (("QuadB200-v0", "QUAD", 10000), ("CoupledWrapperB200-v0", "MONO", 4000), ("DecoupledWrapperB200-v0", "MODUL", 4000)):
        if name not in gym.envs.registration.registry:
            gym.envs.registration.register(id=name, vector_entry_point=_VectorEntry(fw, steps), max_episode_steps=steps)
        ids.append(name)
    return ids


class _VectorEntry:
    def __init__(self, framework, steps):
        self.framework, self.steps = framework, steps

    def __call__(self, num_envs=1, **kw):
        kw.setdefault("max_episode_steps", self.steps)
        return QuadVectorEnv(num_envs, framework=self.framework, **kw)
