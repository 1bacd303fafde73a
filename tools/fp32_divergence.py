#!/usr/bin/env python
"""float32 mode against float64 mode over a free-running horizon (BASELINE north_star: "the fp32 mode agrees to 1e-5
after one step, with observation and reward divergence reported over the horizon").

Runs the REAL step kernel in both precisions on the CPU emulator (tests/host_twin: the float32 instantiation uses exact
1/x and sqrt where the GPU uses MUFU approximations, <= 1 ulp apart), same initial states, same float32 action
sequence, no resets, and prints the divergence of state, observation and reward among the envs still alive in both.
usage: python tools/fp32_divergence.py [n_envs] [steps]   ->  markdown table on stdout"""
import ctypes as C
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import test_host_twin_kernel as tk   # noqa: E402  (HostEnv + build helpers; test infrastructure)
import quad_oracle as qo             # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 400
    tmp = tempfile.mkdtemp()
    K = tk._build("twin_kernel.cpp", tmp, "libtwink.so")
    K.tw_kstep.argtypes = [C.c_void_p, C.POINTER(tk.TwArrays), C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int]
    rng = np.random.default_rng(7)
    st, ig, par = qo.COracle("MONO").reset_from_uniforms(rng.random((n, 20)), qo.ENV_EVAL)
    goal = np.zeros((n, 12)); goal[:, 6] = 1.0
    e64 = tk.HostEnv(K, tk._config(1, True, n_envs=n), warps=4)
    e32 = tk.HostEnv(K, tk._config(1, False, n_envs=n), warps=12)
    for e in (e64, e32):
        e.set_state(st, ig, par, goal)
    alive = np.ones(n, bool)
    marks = [1, 2, 5, 10, 20, 50, 100, 200, 300, 400, 600, 800, 1000]
    print("| step | envs alive | state max abs | obs max abs | obs median abs | reward max abs | done flags differing |")
    print("|---|---|---|---|---|---|---|")
    for t in range(1, steps + 1):
        act = (rng.uniform(-1, 1, (n, 4)) * np.array([0.15, 0.02, 0.02, 0.02])).astype(np.float32)   # keeps most envs in the air
        e64.launch(act); e32.launch(act)
        d64, d32 = e64.done[:, 0].astype(bool), e32.done[:, 0].astype(bool)
        differing = int((d64 != d32)[alive].sum())
        alive &= ~(d64 | d32)
        if t in marks and alive.any():
            ds = np.abs(e32.state.T.astype(np.float64) - e64.state.T)[alive]
            do = np.abs(e32.obs.astype(np.float64) - e64.obs)[alive]
            dr = np.abs(e32.reward.astype(np.float64) - e64.reward)[alive]
            print("| %d | %d | %.2e | %.2e | %.2e | %.2e | %d |" % (t, int(alive.sum()), ds.max(), do.max(), np.median(do), dr.max(), differing))


if __name__ == "__main__":
    main()
