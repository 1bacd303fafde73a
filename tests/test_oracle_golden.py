"""Pins the CPU oracle (oracle/quad_oracle.c and the scipy port) to the reference's golden vectors.

The fixtures under tests/golden/ were produced by oracle/make_golden.py from the UNMODIFIED reference
(/root/reference) and from the reference's own flight log.  Tolerances:
  * float64 state / integrals: 1e-12 relative per step (observed <= 1e-15);
  * float32 obs: bit-exact; done flags: bit-exact; RHS-evaluation counts: identical;
  * reward: exact except where numpy's powf(x, 2) is not correctly rounded (then 1 f32 ulp of one term).
"""
import os

import numpy as np
import pytest

import quad_oracle as qo

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return np.load(os.path.join(G, name))


def _rel(a, b):
    return np.abs(a - b).max() / max(1.0, np.abs(b).max())


@pytest.mark.parametrize("fw,tag,a32", [("MONO", "mono", False), ("MONO", "mono", True),
                                        ("MODUL", "modul", False), ("MODUL", "modul", True)])
def test_c_oracle_step_matches_reference(fw, tag, a32):
    g = _load("step_%s_%s.npz" % (tag, "a32" if a32 else "a64"))
    orc = qo.COracle(fw, act_f32=a32)
    st = g["state_in"].copy(); ig = g["integ_in"].copy()
    obs, rew, done, nfev, status = orc.step(st, ig, g["params"], g["goal"], g["action"])
    assert (status == 0).all()
    assert np.abs(st - g["state_out"]).max() <= 1e-12 * max(1.0, np.abs(g["state_out"]).max())
    assert np.abs(ig - g["integ_out"]).max() <= 1e-12
    assert (obs.view(np.uint32) == g["obs"].view(np.uint32)).all(), "float32 observations must be bit-exact"
    assert (done == g["done"]).all()
    assert (nfev == g["nfev"]).all(), "RHS evaluation counts (2 + 12*attempts) must be identical"
    bad = rew != g["reward"]
    assert bad.mean() <= 5e-3
    assert np.abs(rew - g["reward"]).max() <= 2e-7
    # the adaptive controller is exercised: some steps need more than one attempt
    assert (g["nfev"] > 14).any() and (g["nfev"] == 14).mean() > 0.8


def test_c_oracle_free_running_episode():
    """Feed the oracle its own output over whole episodes (no re-sync) and compare with the reference."""
    g = _load("step_mono_a64.npz")
    orc = qo.COracle("MONO")
    starts = np.flatnonzero(g["episode_start"])
    worst = 0.0
    for s, e in zip(starts, list(starts[1:]) + [len(g["action"])]):
        st = g["state_in"][s:s + 1].copy(); ig = g["integ_in"][s:s + 1].copy()
        for t in range(s, e):
            obs, rew, done, nfev, _ = orc.step(st, ig, g["params"][t:t + 1], g["goal"][t:t + 1], g["action"][t:t + 1])
            worst = max(worst, np.abs(st[0] - g["state_out"][t]).max())
            assert (done[0] == g["done"][t]).all()
        assert (obs[0].view(np.uint32) == g["obs"][e - 1].view(np.uint32)).mean() > 0.9
    assert worst < 1e-9


def test_kat1_reference_flight_log():
    """KAT-1: the authors' own 3 600-row MODUL flight log (printed with %.10f) replays through the oracle."""
    rows = _load("kat1_modul_log.npz")["rows"]
    n = rows.shape[0] - 1
    orc = qo.COracle("MODUL")
    st = rows[:-1, 5:23].copy(); ig = np.zeros((n, 8))
    par = np.tile(np.array([2.15, 0.23, 0.022, 0.035, 0.0135, 2.2]), (n, 1))  # eval reset: nominal parameters
    goal = np.zeros((n, 12)); goal[:, 6] = 1.0
    obs, rew, done, nfev, status = orc.step(st, ig, par, goal, rows[:-1, 0:5].copy())
    err = np.abs(st - rows[1:, 5:23]).max(axis=1)
    assert err.max() < 2e-10, err.max()
    assert (nfev == 14).all()


def test_scipy_port_matches_reference():
    g = _load("step_mono_a64.npz")
    port = qo.ScipyPort("MONO")
    for t in range(0, 200):
        port.set_params(g["params"][t]); port.state = g["state_in"][t].copy()
        port.integ = g["integ_in"][t].copy(); port.goal = g["goal"][t].copy()
        obs, rew, done, _, _ = port.step(g["action"][t].copy())
        assert np.abs(port.state - g["state_out"][t]).max() < 1e-13
        assert (obs[0].view(np.uint32) == g["obs"][t].view(np.uint32)).all()
        assert rew[0] == g["reward"][t, 0] and done[0] == g["done"][t, 0]
        assert port.nfev == g["nfev"][t]
    g = _load("step_modul_a64.npz")
    port = qo.ScipyPort("MODUL")
    for t in range(0, 200):
        port.set_params(g["params"][t]); port.state = g["state_in"][t].copy()
        port.integ = g["integ_in"][t].copy(); port.goal = g["goal"][t].copy()
        obs, rew, done, _, _ = port.step(g["action"][t].copy())
        assert np.abs(port.state - g["state_out"][t]).max() < 1e-13
        assert (np.concatenate(obs).view(np.uint32) == g["obs"][t].view(np.uint32)).all()
        assert list(done) == list(g["done"][t])


@pytest.mark.parametrize("integ", ["solve_ivp", "euler"])
def test_quad_v0_base_env(integ):
    g = _load("quad_v0.npz")
    orc = qo.COracle("QUAD", integrator=qo.INT_EULER if integ == "euler" else qo.INT_DOP853)
    st = g[integ + "_state_in"].copy(); ig = np.zeros((st.shape[0], 8))
    obs, rew, done, nfev, status = orc.step(st, ig, g[integ + "_params"], g[integ + "_goal"], g[integ + "_action"])
    assert np.abs(st - g[integ + "_state_out"]).max() < 1e-12
    assert (done == g[integ + "_done"]).all()
    assert np.abs(rew - g[integ + "_reward"]).max() < 1e-12


def test_so3_projection_matches_numpy_svd():
    rng = np.random.default_rng(0)
    R = np.empty((200, 9))
    for i in range(200):
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        if np.linalg.det(q) < 0:
            q[:, 2] *= -1
        R[i] = (q + rng.normal(scale=10.0 ** rng.uniform(-7, -1), size=(3, 3))).reshape(9, order="F")
    P, k = qo.ensure_so3(R)
    port = qo.ScipyPort("MONO")
    nproj = 0
    for i in range(200):
        ref = port._ensure_SO3(R[i].reshape(3, 3, order="F").copy())
        nproj += not np.shares_memory(ref, R) and not np.array_equal(ref.reshape(9, order="F"), R[i])
        assert np.abs(ref.reshape(9, order="F") - P[i]).max() < 1e-14
    assert k == nproj and 0 < k < 200


def test_reset_distribution_matches_reference_samples():
    """The reset restatement (uniforms -> state/params) has the reference's distributions (quad.py:338-404)."""
    g = _load("reset_samples.npz")
    rng = np.random.default_rng(7)
    n = 8192
    st, ig, par = qo.COracle("MONO").reset_from_uniforms(rng.random((n, 20)), qo.ENV_TRAIN)
    ref_s, ref_p = g["state_train"].astype(np.float64), g["params_train"].astype(np.float64)
    from scipy.stats import ks_2samp
    for j in range(6):
        assert ks_2samp(par[:, j], ref_p[:, j]).pvalue > 1e-3, j
    moving = np.abs(st[:, 0:6]).sum(axis=1) > 0
    ref_moving = np.abs(ref_s[:, 0:6]).sum(axis=1) > 0
    assert abs(moving.mean() - 0.8) < 0.02 and abs(ref_moving.mean() - 0.8) < 0.02
    for j in list(range(0, 6)) + list(range(15, 18)) + [6, 7, 8, 14]:
        assert ks_2samp(st[moving, j], ref_s[ref_moving, j]).pvalue > 1e-3, j
    R = st[:, 6:15].reshape(n, 3, 3).transpose(0, 2, 1)
    assert np.abs(R @ R.transpose(0, 2, 1) - np.eye(3)).max() < 1e-14
    st_e, _, par_e = qo.COracle("MONO").reset_from_uniforms(rng.random((1024, 20)), qo.ENV_EVAL)
    assert np.allclose(par_e, [2.15, 0.23, 0.022, 0.035, 0.0135, 2.2])
    assert np.abs(st_e[:, 0:3]).max() <= 0.4 and np.abs(st_e[:, 3:6]).max() == 0 and np.abs(st_e[:, 15:18]).max() == 0
    assert np.abs(g["state_eval"][:, 0:3]).max() <= 0.4 and np.abs(g["state_eval"][:, 3:6]).max() == 0


def test_philox_known_answer():
    """Random123 known-answer vectors for philox4x32-10."""
    assert qo.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert qo.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert qo.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_trajgen_mode0_matches_reference():
    """Goal generator mode 0: Wd from the pre-step state equals what the reference's TrajectoryGenerator produced
    (trajectory_generator.py:165-172); b1d stays a unit heading vector in the horizontal plane."""
    g = _load("step_mono_a64.npz")
    ok = ~g["episode_start"]
    # every recorded goal was produced by get_desired(env.get_current_state()) right before the step (main.py:145-147)
    wd = qo.traj_wd(g["state_in"], g["goal"][:, 6:9])
    assert np.abs(wd - g["goal"][:, 9:12]).max() < 1e-13
    assert np.abs(g["goal"][:, 0:6]).max() == 0 and np.abs(g["goal"][:, 8]).max() == 0
    # the heading goal was fixed at trajectory start from the FLOAT32 reset state (main.py:226-229):
    # Rz(theta) applied to the current heading, |theta| <= 25 deg
    st32 = g["reset_state32"][~ok].astype(np.float64)
    b1d0 = qo.traj_init_mode0(st32, np.zeros(len(st32)))
    ang = np.arctan2(g["goal"][~ok, 7], g["goal"][~ok, 6]) - np.arctan2(b1d0[:, 1], b1d0[:, 0])
    ang = (ang + np.pi) % (2 * np.pi) - np.pi
    assert np.abs(ang).max() <= np.deg2rad(25) + 1e-12
    b1d_again = qo.traj_init_mode0(st32, ang)
    assert np.abs(b1d_again - g["goal"][~ok, 6:9]).max() < 1e-12


@pytest.mark.parametrize("name,mode", [("hover", 1), ("circle", 5), ("eight", 6), ("circle_manual", 5), ("takeoff", 2), ("land", 3),
                                       ("land_low", 3), ("stay", 4)])
def test_trajgen_modes_match_reference(name, mode):
    """Modes 1 - 6 and the manual fallback, call by call against the reference's TrajectoryGenerator
    (the hover's two random draws are injected from the reference run)."""
    g = _load("traj_modes.npz")
    st, goal_ref, bdd_ref, t_ref = g[name + "_state"], g[name + "_goal"], g[name + "_b1d_dot"], g[name + "_t"]
    t_traj, w, smooth, theta0 = g[name + "_draws"]
    draws = np.array([[(t_traj - 2.0) / 3.0, (w + 0.15 * np.pi) / (0.3 * np.pi)]])
    ts = qo.traj_start(st[0:1])
    if not bool(g[name + "_manual"][-1]):   # manual() overwrites theta_init with the heading at the switch
        assert abs(ts[0, 5] - theta0) < 1e-15
    goal = np.zeros((1, 12)); goal[0, 6] = 1.0
    worst = 0.0
    for i in range(len(t_ref)):
        qo.traj_desired(mode, st[i:i + 1], ts, goal, draws)
        if name == "circle_manual" and int(ts[0, 1]) & 1 and ts[0, 8] > 2.0:
            ts[0, 8] = 1.75            # the golden run shortened the circle (num_circles = 0) to reach manual mode
        worst = max(worst, np.abs(goal[0] - goal_ref[i]).max(), np.abs(ts[0, 9:11] - bdd_ref[i, 0:2]).max())
        assert abs(ts[0, 0] - t_ref[i]) < 1e-12
    assert worst < 1e-11, worst
    assert bool(int(ts[0, 1]) & 2) == bool(g[name + "_manual"][-1])
