"""N1: the vector env's spaces and its gymnasium conformance, without a GPU.  gymnasium is not installed in the build
image, so the conditional code path (subclassing gymnasium.vector.VectorEnv, gymnasium.spaces, registration) is exercised
against a minimal stand-in package injected into sys.modules."""
import importlib
import sys
import types

import numpy as np
import pytest

pytest.importorskip("torch")


def _fake_gymnasium():
    gym = types.ModuleType("gymnasium")
    spaces = types.ModuleType("gymnasium.spaces")
    vector = types.ModuleType("gymnasium.vector")
    envs = types.ModuleType("gymnasium.envs")
    registration = types.ModuleType("gymnasium.envs.registration")

    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            shape = np.shape(low) if shape is None else tuple(shape)
            self.low = np.broadcast_to(np.asarray(low, dtype), shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype), shape).copy()
            self.shape, self.dtype = shape, np.dtype(dtype)

    class Tuple:
        def __init__(self, spaces_):
            self.spaces = tuple(spaces_)

    class VectorEnv:
        pass
    registration.registry = {}

    def register(id, **kw):
        registration.registry[id] = kw
    registration.register = register
    spaces.Box, spaces.Tuple, vector.VectorEnv = Box, Tuple, VectorEnv
    envs.registration = registration
    gym.spaces, gym.vector, gym.envs = spaces, vector, envs
    return {"gymnasium": gym, "gymnasium.spaces": spaces, "gymnasium.vector": vector, "gymnasium.envs": envs,
            "gymnasium.envs.registration": registration}


def _cfg():
    from gym_rotor_b200 import _native
    c = _native.QrConfig()
    c.x_lim, c.v_lim, c.W_lim = 1.0, 4.0, 2 * np.pi
    return c


def test_spaces_without_gymnasium():
    from gym_rotor_b200 import vec_env
    if vec_env._gymnasium() is not None:
        pytest.skip("gymnasium is installed here")
    so, sa, bo, ba = vec_env.make_spaces("QUAD", 8, _cfg())
    # quad.py:104-132: +-x_lim (3), +-v_lim (3), +-1 (9), +-W_lim (3); actions in [-1, 1]^4
    assert so.shape == (18,) and so.dtype == np.float32 and sa.shape == (4,) and bo.shape == (8, 18) and ba.shape == (8, 4)
    assert np.array_equal(so.high, np.concatenate([np.ones(3), 4 * np.ones(3), np.ones(9), 2 * np.pi * np.ones(3)]).astype(np.float32))
    assert np.array_equal(so.low, -so.high) and (sa.low == -1).all() and (sa.high == 1).all()
    so, sa, bo, ba = vec_env.make_spaces("MONO", 8, _cfg())
    assert so.shape == (23,) and bo.shape == (8, 23) and sa.shape == (4,) and so.contains(np.zeros(23, np.float32))
    so, sa, bo, ba = vec_env.make_spaces("MODUL", 8, _cfg())
    assert [s.shape for s in so.spaces] == [(15,), (3,)] and [s.shape for s in bo.spaces] == [(8, 15), (8, 3)] and sa.shape == (5,)
    assert len(so.sample()) == 2 and vec_env.register_envs() == []
    assert vec_env.QuadVectorEnv.__mro__[1] is object


def test_vector_env_subclasses_gymnasium_when_importable(monkeypatch):
    mods = _fake_gymnasium()
    for k, v in mods.items():
        monkeypatch.setitem(sys.modules, k, v)
    import gym_rotor_b200.vec_env as ve
    ve = importlib.reload(ve)
    try:
        assert issubclass(ve.QuadVectorEnv, mods["gymnasium.vector"].VectorEnv)
        so, sa, bo, ba = ve.make_spaces("MONO", 4, _cfg())
        assert isinstance(so, mods["gymnasium.spaces"].Box) and bo.shape == (4, 23)
        so, *_ = ve.make_spaces("MODUL", 4, _cfg())
        assert isinstance(so, mods["gymnasium.spaces"].Tuple)
        ids = ve.register_envs()
        reg = mods["gymnasium.envs.registration"].registry
        assert ids == ["QuadB200-v0", "CoupledWrapperB200-v0", "DecoupledWrapperB200-v0"] and set(ids) <= set(reg)
        assert reg["QuadB200-v0"]["max_episode_steps"] == 10000          # gym_rotor/__init__.py:3-7
        assert callable(reg["CoupledWrapperB200-v0"]["vector_entry_point"])
    finally:
        for k in mods:
            monkeypatch.delitem(sys.modules, k, raising=False)
        importlib.reload(ve)
