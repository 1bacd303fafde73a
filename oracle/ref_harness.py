"""Harness around the UNMODIFIED reference (fdcl-gwu/gym-rotor at /root/reference).

TEST INFRASTRUCTURE ONLY -- never imported by the product path (gym_rotor_b200/).
It only works in the build container, where /root/reference exists; the GPU box
has no /root/reference, so nothing in `-m gpu` tests, smoke() or bench.py may use
it.  Its job: import the reference's own numpy env to (a) pin the C/numpy
restatement in oracle/quad_oracle.{c,py} and (b) generate the golden vectors
committed under tests/golden/ (see oracle/make_golden.py).

Quirks of the reference this harness has to respect:
  * every constructor parses sys.argv (quad.py:24-25, coupled:18-19, decoupled:19-20,
    trajectory_generator.py:13-14) -> sys.argv is sanitised before construction;
  * `gymnasium`, `matplotlib`, `plum` are not installed -> stand-ins in oracle/ref_shims;
  * all randomness is the global numpy legacy RNG + python `random` (quad.py:6,342).
"""
import os
import sys
import random

import numpy as np

REFERENCE_ROOT = os.environ.get("GYM_ROTOR_REFERENCE", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_shims")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "gym_rotor"))


def _prepare_path():
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    for p in (REFERENCE_ROOT, _SHIMS):
        if p in sys.path:
            sys.path.remove(p)
    # shims first only for modules that are really missing
    sys.path.insert(0, REFERENCE_ROOT)
    sys.path.insert(0, _SHIMS)


def make_env(framework):
    """Construct CoupledWrapper ('MONO') or DecoupledWrapper ('MODUL') exactly as main.py:42,52 does."""
    _prepare_path()
    argv = sys.argv
    sys.argv = ["x", "--framework", framework]
    try:
        from gym_rotor.wrappers.coupled_yaw_wrapper import CoupledWrapper
        from gym_rotor.wrappers.decoupled_yaw_wrapper import DecoupledWrapper
        env = CoupledWrapper() if framework == "MONO" else DecoupledWrapper()
    finally:
        sys.argv = argv
    return env


def make_base_env():
    """Base `Quad-v0` env.  QuadEnv.step() raises as shipped (quad.py:153-158 index a scalar
    reward); the list-returning subclass below is the minimal fix SURVEY 8(d) config 1 names."""
    _prepare_path()
    argv = sys.argv
    sys.argv = ["x", "--framework", "MONO"]
    try:
        from gym_rotor.envs.quad import QuadEnv

        class QuadEnvList(QuadEnv):
            def reward_wrapper(self, obs):
                return [QuadEnv.reward_wrapper(self, obs)]

            def done_wrapper(self, obs):
                return [QuadEnv.done_wrapper(self, obs)]

        env = QuadEnvList()
    finally:
        sys.argv = argv
    return env


def make_trajgen(env):
    _prepare_path()
    argv = sys.argv
    sys.argv = ["x", "--framework", env.framework]
    try:
        from utils.trajectory_generator import TrajectoryGenerator
        tg = TrajectoryGenerator(env)
    finally:
        sys.argv = argv
    return tg


def seed_all(seed):
    random.seed(seed)
    np.random.seed(seed)


# ---- full mutable state of one env (SURVEY 8c "State to inject") -------------------------

def get_params(env):
    """(m, d, J1, J3, c_tf, c_tw) as drawn by quad.py:359-404."""
    return np.array([env.m, env.d, env.J[0, 0], env.J[2, 2], env.c_tf, env.c_tw], dtype=np.float64)


def set_params(env, p):
    m, d, J1, J3, c_tf, c_tw = [float(v) for v in p]
    env.m, env.d = m, d
    env.J = np.diag([J1, J1, J3])
    env.c_tf, env.c_tw = c_tf, c_tw
    # derived quantities, same expressions/order as quad.py:388-404
    env.f = env.m * env.g
    env.hover_force = env.m * env.g / 4.0
    env.min_force = 0.5
    env.max_force = env.c_tw * env.hover_force
    env.forces_to_fM = np.array([
        [1.0, 1.0, 1.0, 1.0],
        [0.0, -env.d, 0.0, env.d],
        [env.d, 0.0, -env.d, 0.0],
        [-env.c_tf, env.c_tf, -env.c_tf, env.c_tf]])
    env.fM_to_forces = np.linalg.inv(env.forces_to_fM)
    env.avrg_act = (env.min_force + env.max_force) / 2.0
    env.scale_act = env.max_force - env.avrg_act


def get_integ(env):
    """[eIx.error(3), eIx.integrand(3), eIb1.error, eIb1.integrand]"""
    return np.concatenate([np.asarray(env.eIx.error, float), np.asarray(env.eIx.integrand, float),
                           [float(env.eIb1.error)], [float(env.eIb1.integrand)]])


def set_integ(env, v):
    v = np.asarray(v, dtype=np.float64)
    env.eIx.error = v[0:3].copy()
    env.eIx.integrand = v[3:6].copy()
    env.eIb1.error = float(v[6])
    env.eIb1.integrand = float(v[7])


def get_goal(env):
    """[xd(3), vd(3), b1d(3), Wd(3)]"""
    return np.concatenate([env.xd, env.vd, env.b1d, env.Wd]).astype(np.float64)


def set_goal(env, g):
    g = np.asarray(g, dtype=np.float64)
    env.set_goal_state(g[0:3].copy(), g[3:6].copy(), g[6:9].copy(), np.zeros(3), g[9:12].copy())


class RhsCounter:
    """Counts RHS evaluations per env.step() by wrapping the bound EoM (quad.py:321 / decoupled:143)."""

    def __init__(self, env):
        self.env = env
        self.n = 0
        name = "decouple_EoM" if hasattr(env, "decouple_EoM") else "EoM"
        orig = getattr(env, name)

        def counted(t, y):
            self.n += 1
            return orig(t, y)

        setattr(env, name, counted)

    def take(self):
        n, self.n = self.n, 0
        return n
