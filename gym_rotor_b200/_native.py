"""ctypes binding of the C ABI in include/quadrotor_b200.h.

There is no CPU fallback and no alternative backend: if the shared library is missing or no CUDA device
is present, everything here raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QR_LIB_PATH") or os.path.join(HERE, "csrc", "libquadrotor_b200.so")  # override: A/B kernel variants

QR_OK = 0
MODE_QUAD, MODE_COUPLED, MODE_DECOUPLED = 0, 1, 2
F32, F64 = 0, 1
INT_DOP853, INT_EULER = 0, 1
ENV_TRAIN, ENV_EVAL = 0, 1
GOAL_EXTERNAL, GOAL_TRAJ_MODE0, GOAL_TRAJ_HOVER, GOAL_TRAJ_CIRCLE, GOAL_TRAJ_EIGHT = 0, 1, 2, 3, 4
GOAL_TRAJ_TAKEOFF, GOAL_TRAJ_LAND, GOAL_TRAJ_STAY = 5, 6, 7
ACT_POLICY = 2   # qr_rollout: actions from the shipped TD3 actor, evaluated inside the kernel
ST_NONFINITE, ST_TOO_SMALL_STEP, ST_SVD = 1, 2, 4
NUM_STATS = 20
ABI_VERSION = 2
STAT_NAMES = ["episodes", "return0", "return1", "length", "crashed", "truncated", "return0_sq", "steps",
              "bad_status", "nfev", "attempts_1", "attempts_2", "attempts_3", "attempts_4p", "reward0",
              "so3_projections", "bench_reward", "solved_at_limit", "reserved18", "reserved19"]

# every symbol include/quadrotor_b200.h declares (checked by tests/test_cabi_symbols.py)
EXPORTS = ["qr_default_config", "qr_create", "qr_destroy", "qr_get_config", "qr_get_buffers", "qr_reset",
           "qr_init_goal", "qr_goal_update", "qr_norm_error_state", "qr_policy_td3", "qr_step", "qr_rollout", "qr_step_host", "qr_set_state_host",
           "qr_get_state_host", "qr_stats", "qr_launch_count", "qr_last_error", "qr_abi_version", "qr_obs_stride"]


class QrConfig(C.Structure):
    _fields_ = [("n_envs", C.c_int64), ("env_id_offset", C.c_int64), ("seed", C.c_uint64),
                ("mode", C.c_int32), ("dtype", C.c_int32), ("integrator", C.c_int32), ("autoreset", C.c_int32),
                ("goal_mode", C.c_int32), ("env_type", C.c_int32), ("max_episode_steps", C.c_int32),
                ("diagnostics", C.c_int32), ("round_returns", C.c_int32), ("reserved1", C.c_int32)] + [
        (n, C.c_double) for n in (
            "dt", "g", "rtol", "atol", "x_lim", "v_lim", "W_lim", "eIx_lim", "eIb1_lim", "sat_sigma", "alpha", "beta",
            "Cx", "CIx", "Cv", "Cb1", "CIb1", "CW", "Cw12", "CW3", "reward_min", "reward_min_1", "reward_min_2",
            "min_force", "euler_lim_deg", "udm_pct")]


class QrBuffers(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "state", "integ", "params", "goal", "obs", "reward", "done", "terminated", "truncated", "final_obs", "nfev",
        "status", "ep_return", "ep_length", "ep_index", "stats")] + [
        ("obs_dim", C.c_int32), ("act_dim", C.c_int32), ("n_agents", C.c_int32), ("elem_size", C.c_int32),
        ("n_envs", C.c_int64), ("traj", C.c_void_p)]


class NativeError(RuntimeError):
    pass


_lib = None


def load():
    """Loads the in-tree CUDA library; raises if it has not been built (python -m gym_rotor_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeError(
            "%s is missing: build it with `python -m gym_rotor_b200.build` (nvcc, sm_100a). "
            "There is no CPU or PyTorch fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, u8p = C.c_void_p, C.c_void_p
    L.qr_default_config.argtypes = [C.POINTER(QrConfig), C.c_int, C.c_int]
    L.qr_create.argtypes = [C.POINTER(QrConfig), C.c_int, C.POINTER(C.c_void_p)]
    L.qr_destroy.argtypes = [vp]
    L.qr_get_config.argtypes = [vp, C.POINTER(QrConfig)]
    L.qr_get_buffers.argtypes = [vp, C.POINTER(QrBuffers)]
    L.qr_reset.argtypes = [vp, u8p, C.c_int, vp]
    L.qr_init_goal.argtypes = [vp, u8p, vp]
    L.qr_goal_update.argtypes = [vp, vp]
    L.qr_norm_error_state.argtypes = [vp, u8p, vp]
    L.qr_policy_td3.argtypes = [vp, vp, vp]
    L.qr_step.argtypes = [vp, vp, C.c_int, vp]
    L.qr_rollout.argtypes = [vp, C.c_int, vp, C.c_int, vp, vp, vp, vp]
    L.qr_step_host.argtypes = [vp, vp, C.c_int, vp, vp, vp, vp]
    L.qr_set_state_host.argtypes = [vp, vp, vp, vp, vp]
    L.qr_get_state_host.argtypes = [vp, vp, vp, vp, vp]
    L.qr_stats.argtypes = [vp, C.POINTER(C.c_double), C.c_int, vp]
    L.qr_launch_count.restype = C.c_int64
    L.qr_last_error.restype = C.c_char_p
    L.qr_abi_version.restype = C.c_int
    L.qr_obs_stride.argtypes = [vp]
    L.qr_obs_stride.restype = C.c_int
    for name in EXPORTS:
        getattr(L, name)
    if L.qr_abi_version() != ABI_VERSION:
        raise NativeError("%s has ABI version %d, this package needs %d: rebuild it (python -m gym_rotor_b200.build --force)"
                          % (LIB_PATH, L.qr_abi_version(), ABI_VERSION))
    _lib = L
    return L


def check(rc):
    if rc != QR_OK:
        raise NativeError("quadrotor_b200 error %d: %s" % (rc, load().qr_last_error().decode()))
