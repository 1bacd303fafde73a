"""Stand-in for the `gymnasium` package (not installed in this image, no network).

TEST INFRASTRUCTURE ONLY.  It exists so that the unmodified reference under
/root/reference can be imported by oracle/ref_harness.py to generate golden
vectors.  It provides exactly what gym_rotor/envs/quad.py:12-14,19,120-132 and
gym_rotor/__init__.py:1-7 touch: an `Env` base class, `spaces.Box`, `utils.seeding`
and `envs.registration`.
"""
import numpy as np
from . import spaces  # noqa: F401
from . import utils  # noqa: F401
from . import envs  # noqa: F401


class Env:
    metadata = {}

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self.np_random = np.random.default_rng(seed)

    def close(self):
        pass
