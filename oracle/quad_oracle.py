"""ctypes binding of the C oracle (oracle/quad_oracle.c) + a numpy/scipy port of the reference step.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product (gym_rotor_b200/) never imports this module.

Two restatements live here:
  * `COracle`   -- the plain-C restatement (own DOP853 driver), fast enough for 4096 x 1000-step parity runs;
  * `ScipyPort` -- a line-by-line numpy port that, like the reference, hands a Python RHS to
                   scipy.integrate.solve_ivp(method='DOP853') (coupled_yaw_wrapper.py:63).  It has the
                   reference's own cost structure (about 600 steps/s/core) and is what bench.py times as
                   the "reference CPU path" on the GPU box, where /root/reference does not exist.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libquad_oracle.so")

MODE_QUAD, MODE_COUPLED, MODE_DECOUPLED = 0, 1, 2
INT_DOP853, INT_EULER = 0, 1
ENV_TRAIN, ENV_EVAL = 0, 1
MODES = {"QUAD": MODE_QUAD, "MONO": MODE_COUPLED, "MODUL": MODE_DECOUPLED}


class QoConfig(C.Structure):
    _fields_ = [("mode", C.c_int32), ("integrator", C.c_int32), ("act_f32", C.c_int32), ("reserved", C.c_int32)] + [
        (n, C.c_double) for n in (
            "dt", "g", "rtol", "atol", "x_lim", "v_lim", "W_lim", "eIx_lim", "eIb1_lim", "sat_sigma", "alpha",
            "beta", "Cx", "CIx", "Cv", "Cb1", "CIb1", "CW", "Cw12", "CW3", "reward_min", "reward_min_1",
            "reward_min_2", "min_force", "euler_lim_deg")]


def build(force=False):
    """Compile oracle/libquad_oracle.so with gcc (make -C oracle)."""
    if force or not os.path.exists(_LIB_PATH) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
            for f in ("quad_oracle.c", "quad_oracle_impl.inc", "quad_oracle.h", "dop853_tableau.h")):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp, fp, u8p, i32p = (C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.POINTER(C.c_int32))
        L.qo_default_config.argtypes = [C.POINTER(QoConfig), C.c_int]
        L.qo_step_f64.argtypes = [C.POINTER(QoConfig), C.c_int64, dp, dp, dp, dp, dp, fp, dp, u8p, i32p, u8p]
        L.qo_step_f32.argtypes = [C.POINTER(QoConfig), C.c_int64, fp, fp, fp, fp, fp, fp, fp, u8p, i32p, u8p]
        L.qo_norm_error_state_f64.argtypes = [C.POINTER(QoConfig), C.c_int64, dp, dp, dp, fp]
        L.qo_rhs_f64.argtypes = [C.c_int64, dp, dp, dp, dp]
        L.qo_ensure_so3_f64.argtypes = [C.c_int64, dp]
        L.qo_ensure_so3_f64.restype = C.c_int64
        L.qo_reset_from_uniforms_f64.argtypes = [C.POINTER(QoConfig), C.c_int, C.c_double, C.c_int64, dp, dp, dp, dp]
        L.qo_traj_wd_f64.argtypes = [C.c_int64, dp, dp, dp]
        L.qo_traj_init_mode0_f64.argtypes = [C.c_int64, dp, dp, dp]
        L.qo_traj_start_f64.argtypes = [C.c_int64, dp, dp]
        L.qo_traj_desired_f64.argtypes = [C.c_int, C.c_int64, dp, dp, dp, dp, C.c_double]
        L.qo_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.qo_set_threads.argtypes = [C.c_int]
        L.qo_stage_projection_count.argtypes = [C.c_int]
        L.qo_stage_projection_count.restype = C.c_longlong
        L.qo_get_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct)) if a is not None else None


class COracle:
    """Batched env.step() through the C restatement.  Arrays are row-major [n, ...] (see quad_oracle.h)."""

    def __init__(self, framework="MONO", act_f32=False, integrator=INT_DOP853, threads=1):
        self.L = lib()
        self.mode = MODES[framework]
        self.cfg = QoConfig()
        self.L.qo_default_config(C.byref(self.cfg), self.mode)
        self.cfg.act_f32 = int(act_f32)
        self.cfg.integrator = integrator
        self.threads = threads
        self.obs_dim = 23 if self.mode == MODE_COUPLED else 18
        self.act_dim = 5 if self.mode == MODE_DECOUPLED else 4
        self.n_agents = 2 if self.mode == MODE_DECOUPLED else 1

    def step(self, state, integ, params, goal, action):
        """state[n,18], integ[n,8] are updated IN PLACE (float64, C-contiguous).  Returns obs, reward, done, nfev, status."""
        n = state.shape[0]
        for a in (state, integ):
            assert a.dtype == np.float64 and a.flags.c_contiguous
        params = np.ascontiguousarray(params, np.float64)
        goal = np.ascontiguousarray(goal, np.float64)
        action = np.ascontiguousarray(action, np.float64)
        assert action.shape == (n, self.act_dim)
        obs = np.empty((n, self.obs_dim), np.float32)
        reward = np.empty((n, self.n_agents), np.float64)
        done = np.empty((n, self.n_agents), np.uint8)
        nfev = np.empty(n, np.int32)
        status = np.empty(n, np.uint8)
        self.L.qo_set_threads(self.threads)
        self.L.qo_step_f64(C.byref(self.cfg), n, _p(state, C.c_double), _p(integ, C.c_double), _p(params, C.c_double),
                           _p(goal, C.c_double), _p(action, C.c_double), _p(obs, C.c_float), _p(reward, C.c_double),
                           _p(done, C.c_uint8), _p(nfev, C.c_int32), _p(status, C.c_uint8))
        return obs, reward, done.astype(bool), nfev, status

    def step_f32(self, state, integ, params, goal, action):
        n = state.shape[0]
        for a in (state, integ):
            assert a.dtype == np.float32 and a.flags.c_contiguous
        params = np.ascontiguousarray(params, np.float32)
        goal = np.ascontiguousarray(goal, np.float32)
        action = np.ascontiguousarray(action, np.float32)
        obs = np.empty((n, self.obs_dim), np.float32)
        reward = np.empty((n, self.n_agents), np.float32)
        done = np.empty((n, self.n_agents), np.uint8)
        nfev = np.empty(n, np.int32)
        status = np.empty(n, np.uint8)
        self.L.qo_set_threads(self.threads)
        self.L.qo_step_f32(C.byref(self.cfg), n, _p(state, C.c_float), _p(integ, C.c_float), _p(params, C.c_float),
                           _p(goal, C.c_float), _p(action, C.c_float), _p(obs, C.c_float), _p(reward, C.c_float),
                           _p(done, C.c_uint8), _p(nfev, C.c_int32), _p(status, C.c_uint8))
        return obs, reward, done.astype(bool), nfev, status

    def norm_error_state(self, state, integ, goal):
        n = state.shape[0]
        obs = np.empty((n, self.obs_dim), np.float32)
        state = np.ascontiguousarray(state, np.float64)
        goal = np.ascontiguousarray(goal, np.float64)
        self.L.qo_norm_error_state_f64(C.byref(self.cfg), n, _p(state, C.c_double), _p(integ, C.c_double),
                                       _p(goal, C.c_double), _p(obs, C.c_float))
        return obs

    def reset_from_uniforms(self, u, env_type=ENV_TRAIN, udm_pct=10.0):
        u = np.ascontiguousarray(u, np.float64)
        n = u.shape[0]
        assert u.shape == (n, 20)
        state = np.empty((n, 18)); integ = np.empty((n, 8)); params = np.empty((n, 6))
        self.L.qo_reset_from_uniforms_f64(C.byref(self.cfg), env_type, udm_pct, n, _p(u, C.c_double),
                                          _p(state, C.c_double), _p(integ, C.c_double), _p(params, C.c_double))
        return state, integ, params


def rhs(y, params, fM):
    y = np.ascontiguousarray(y, np.float64); params = np.ascontiguousarray(params, np.float64)
    fM = np.ascontiguousarray(fM, np.float64)
    out = np.empty_like(y)
    lib().qo_rhs_f64(y.shape[0], _p(y, C.c_double), _p(params, C.c_double), _p(fM, C.c_double), _p(out, C.c_double))
    return out


def ensure_so3(R):
    """R[n,9] column-major; returns (projected copy, number re-projected)."""
    R = np.array(R, dtype=np.float64, order="C", copy=True)
    k = lib().qo_ensure_so3_f64(R.shape[0], _p(R, C.c_double))
    return R, int(k)


def traj_wd(state, b1d):
    """Mode-0 goal generator: Wd[n,3] from the current state and the stored heading goal."""
    state = np.ascontiguousarray(state, np.float64); b1d = np.ascontiguousarray(b1d, np.float64)
    out = np.empty((state.shape[0], 3))
    lib().qo_traj_wd_f64(state.shape[0], _p(state, C.c_double), _p(b1d, C.c_double), _p(out, C.c_double))
    return out


def traj_init_mode0(state, theta):
    state = np.ascontiguousarray(state, np.float64); theta = np.ascontiguousarray(theta, np.float64)
    out = np.empty((state.shape[0], 3))
    lib().qo_traj_init_mode0_f64(state.shape[0], _p(state, C.c_double), _p(theta, C.c_double), _p(out, C.c_double))
    return out


def traj_start(state):
    """mark_traj_start: fresh per-env trajectory state ts[n,12] from the (float32-valued) reset state."""
    state = np.ascontiguousarray(state, np.float64)
    ts = np.zeros((state.shape[0], 12))
    lib().qo_traj_start_f64(state.shape[0], _p(state, C.c_double), _p(ts, C.c_double))
    return ts


def traj_desired(mode, state, ts, goal, draws, dt=1. / 200):
    """One get_desired(state, mode) call for mode 1 / 5 / 6; ts and goal are updated in place."""
    state = np.ascontiguousarray(state, np.float64); draws = np.ascontiguousarray(draws, np.float64)
    assert ts.dtype == np.float64 and goal.dtype == np.float64 and ts.flags.c_contiguous and goal.flags.c_contiguous
    lib().qo_traj_desired_f64(int(mode), state.shape[0], _p(state, C.c_double), _p(ts, C.c_double), _p(goal, C.c_double),
                              _p(draws, C.c_double), float(dt))


def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*[int(v) & 0xFFFFFFFF for v in ctr])
    k = (C.c_uint32 * 2)(*[int(v) & 0xFFFFFFFF for v in key])
    o = (C.c_uint32 * 4)()
    lib().qo_philox4x32_10(c, k, o)
    return [int(v) for v in o]


# ------------------------------------------------------------------------------------------------------
# numpy / scipy port (reference call structure: Python RHS handed to scipy's DOP853)
# ------------------------------------------------------------------------------------------------------

class ScipyPort:
    """One env, numpy state, stepping exactly like CoupledWrapper / DecoupledWrapper.step().

    Restates quad.py:142-168, 321-335, 421-466; coupled_yaw_wrapper.py:44-110;
    decoupled_yaw_wrapper.py:49-161; quad_utils.py:12-26, 38-63, 80-85, 123-142, 226-240.
    """

    def __init__(self, framework="MONO"):
        from scipy.integrate import solve_ivp
        self._solve_ivp = solve_ivp
        self.framework = framework
        self.g, self.dt = 9.81, 1. / 200
        self.x_lim, self.v_lim, self.W_lim = 1.0, 4.0, 2 * np.pi
        self.eIx_lim = self.eIb1_lim = 3.0
        self.alpha, self.beta, self.sat_sigma = 0.01, 0.05, 1.
        self.Cx, self.CIx, self.Cv, self.Cw12 = 6.0, 0.1, 0.4, 0.6
        self.Cb1, self.CIb1, self.CW3 = 6.0, 0.1, 0.1
        self.CW = self.Cw12
        self.reward_min = -np.ceil(self.Cx + self.CIx + self.Cv + self.Cb1 + self.CIb1 + self.CW)
        self.reward_min_1 = -np.ceil(self.Cx + self.CIx + self.Cv + self.Cw12)
        self.reward_min_2 = -np.ceil(self.Cb1 + self.CW3 + self.CIb1)
        self.e3 = np.array([0., 0., 1.])
        self.set_params([2.15, 0.23, 0.022, 0.035, 0.0135, 2.2])
        self.state = np.zeros(18); self.state[[6, 10, 14]] = 1.
        self.integ = np.zeros(8)
        self.goal = np.zeros(12); self.goal[6] = 1.
        self.nfev = 0

    def set_params(self, p):
        self.m, self.d, J1, J3, self.c_tf, self.c_tw = [float(v) for v in p]
        self.J = np.diag([J1, J1, J3])
        self.hover_force = self.m * self.g / 4.0
        self.min_force = 0.5
        self.max_force = self.c_tw * self.hover_force
        self.avrg_act = (self.min_force + self.max_force) / 2.0
        self.scale_act = self.max_force - self.avrg_act

    @staticmethod
    def _hat(x):
        return np.array([[0.0, -x[2], x[1]], [x[2], 0.0, -x[0]], [-x[1], x[0], 0.0]])

    @staticmethod
    def _ensure_SO3(R, tol=1e-5):
        if np.allclose(R.T @ R, np.eye(3), rtol=tol, atol=tol) and np.isclose(np.linalg.det(R), 1., rtol=tol):
            return R
        U, s, VT = np.linalg.svd(R)
        dU, dV = np.linalg.det(U), np.linalg.det(VT)
        U[:, 2] = U[:, 2] * dU
        VT[2, :] = VT[2, :] * dV
        return U @ VT

    def _eom(self, t, y):
        self.nfev += 1
        R = self._ensure_SO3(y[6:15].reshape(3, 3, order='F'))
        v, W = y[3:6], y[15:18]
        v_dot = self.g * self.e3 - self.f * R @ self.e3 / self.m
        R_dot = (R @ self._hat(W)).reshape(1, 9, order='F')
        W_dot = np.linalg.inv(self.J) @ (-self._hat(W) @ self.J @ W + self.M)
        return np.concatenate([v.flatten(), v_dot.flatten(), R_dot.flatten(), W_dot.flatten()])

    def norm_error_state(self):
        s, g = self.state, self.goal
        R = self._ensure_SO3(s[6:15].reshape(3, 3, order='F'))
        ex = s[0:3] / self.x_lim - g[0:3] / self.x_lim
        ev = s[3:6] / self.v_lim - g[3:6] / self.v_lim
        eW = s[15:18] / self.W_lim - g[9:12] / self.W_lim
        b1, b2, b3 = R[:, 0], R[:, 1], R[:, 2]
        b1d = g[6:9]
        b1c = b1d - np.dot(b1d, b3) * b3
        eb1_norm = np.arctan2(-np.dot(b1c, b2), np.dot(b1c, b1)) / np.pi
        I = self.integ
        gx = -self.alpha * I[0:3] + ex * self.x_lim
        I[0:3] = I[0:3] + (I[3:6] + gx) * self.dt / 2.0
        I[3:6] = gx
        eIx_n = np.clip(I[0:3] / self.eIx_lim, -self.sat_sigma, self.sat_sigma)
        gb = -self.beta * I[6] + eb1_norm * np.pi
        I[6] = I[6] + (I[7] + gb) * self.dt / 2.0
        I[7] = gb
        eIb1_n = np.clip(I[6] / self.eIb1_lim, -self.sat_sigma, self.sat_sigma)
        if self.framework == "MODUL":
            ew12 = eW[0] * b1 + eW[1] * b2
            return [np.concatenate((ex, eIx_n, ev, b3, ew12), axis=None, dtype=np.float32),
                    np.concatenate((eb1_norm, eIb1_n, eW[2]), axis=None, dtype=np.float32)]
        return [np.concatenate((ex, eIx_n, ev, R.reshape(9, 1, order='F').flatten(), eb1_norm, eIb1_n, eW),
                               axis=None, dtype=np.float32)]

    def step(self, action):
        from numpy.linalg import norm
        self.f = (4 * (self.scale_act * action[0] + self.avrg_act)).clip(4 * self.min_force, 4 * self.max_force)
        y = self.state.copy()
        R = self._ensure_SO3(y[6:15].reshape(3, 3, order='F'))
        y[6:15] = R.reshape(9, 1, order='F').flatten()
        if self.framework == "MODUL":
            tau, W = action[1:4], y[15:18]
            self.M = np.array([R[:, 0] @ tau + self.J[2, 2] * W[2] * W[1],
                               R[:, 1] @ tau - self.J[2, 2] * W[2] * W[0], action[4]], dtype=np.float64)
        else:
            self.M = action[1:4]
        self.nfev = 0
        sol = self._solve_ivp(self._eom, [0, self.dt], y, method='DOP853')
        self.state = sol.y[:, -1]
        obs = self.norm_error_state()
        if self.framework == "MODUL":
            o1, o2 = obs
            r1 = (-self.Cx * (norm(o1[0:3], 2) ** 2) + -self.CIx * (norm(o1[3:6], 2) ** 2)
                  + -self.Cv * (norm(o1[6:9], 2) ** 2) + -self.Cw12 * (norm(o1[12:15], 2) ** 2))
            r2 = -self.Cb1 * abs(o2[0]) + -self.CIb1 * (abs(o2[1]) ** 2) + -self.CW3 * (abs(o2[2]) ** 2)
            rew = [np.interp(r1, [self.reward_min_1, 0.], [0., 1.]), np.interp(r2, [self.reward_min_2, 0.], [0., 1.])]
            done = [bool((abs(o1[0:3]) >= 1.0).any() or (abs(o1[6:9]) >= 1.0).any() or (abs(o1[12:15]) >= 1.0).any()),
                    bool(abs(o2[2]) >= 1.0)]
        else:
            o = obs[0]
            r = (-self.Cx * (norm(o[0:3], 2) ** 2) + -self.CIx * (norm(o[3:6], 2) ** 2) + -self.Cv * (norm(o[6:9], 2) ** 2)
                 + -self.Cb1 * abs(o[18]) + -self.CIb1 * (abs(o[19]) ** 2) + -self.CW * (norm(o[20:23], 2) ** 2))
            rew = [np.interp(r, [self.reward_min, 0.], [0., 1.])]
            done = [bool((abs(o[0:3]) >= 1.0).any() or (abs(o[6:9]) >= 1.0).any() or (abs(o[20:23]) >= 1.0).any())]
        for i, d in enumerate(done):
            if d:
                rew[i] = -1.
        return obs, rew, done, False, {}

    def reset(self, env_type='train', rng=None):
        """quad.py:171-222 restated on a numpy Generator (the reference draws from the global legacy RNG)."""
        rng = rng if rng is not None else np.random.default_rng()
        u = rng.random(20)
        st, integ, par = COracle(self.framework).reset_from_uniforms(
            u[None, :], ENV_TRAIN if env_type == 'train' else ENV_EVAL)
        self.set_params(par[0]); self.state = st[0].copy(); self.integ = integ[0].copy()
        return self.state.astype(np.float32)
