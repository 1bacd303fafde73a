"""gym_rotor_b200 -- B200-native batched quadrotor simulator: the env.step() hot path of fdcl-gwu/gym-rotor.

Only what that path needs: csrc/ (sm_100a kernels + the C ABI of include/quadrotor_b200.h), the ctypes
binding, and the host-side mirror of the reference's env interface.  Importing this package never
compiles or falls back to anything: build with `python -m gym_rotor_b200.build`.
"""
from ._native import NativeError, load  # noqa: F401


def __getattr__(name):
    # torch-dependent classes are imported lazily so that `import gym_rotor_b200` stays cheap
    if name in ("BatchedQuadEnv", "CoupledWrapper", "DecoupledWrapper", "QuadEnv", "QuadVectorEnv", "register_envs", "make_spaces"):
        from . import vec_env
        return getattr(vec_env, name)
    raise AttributeError(name)
