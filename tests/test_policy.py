"""Effective-weight actor inference: the extracted maps reproduce the reference actors' outputs,
and on the GPU the policy-in-the-loop episode reproduces the reference's own evaluation run (KAT-2)."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("tag,agents", [("mono", 1), ("modul", 2)])
def test_effective_actor_matches_reference_outputs(tag, agents):
    from effective_actor import EffectiveActor
    z = np.load(os.path.join(G, "policy_td3_%s.npz" % tag))
    for i in range(agents):
        pi = EffectiveActor(z, agent=i, device="cpu")
        act = pi(torch.as_tensor(z["a%d_obs" % i])).numpy()
        ref = z["a%d_act" % i]
        assert act.shape == ref.shape
        # float32 forward in a different operation order; the yaw actor has pre-activations of magnitude ~6
        assert np.abs(act - ref).max() < 1e-4, (tag, i, np.abs(act - ref).max())
        assert np.abs(ref).max() <= 1.0 and ref.std() > 0.05


@pytest.mark.gpu
@pytest.mark.parametrize("fw,tag,A", [("MONO", "mono", [4]), ("MODUL", "modul", [4, 1])])
def test_policy_in_the_loop_reproduces_reference_eval_episode(fw, tag, A):
    """BASELINE config 5 at N=1 scale: fp64 env on the GPU + the extracted actor, 1000 steps, against the episode the
    reference flew with the same checkpoint (return 989.8 MONO; 992.6 / 998.0 MODUL)."""
    from gym_rotor_b200 import vec_env
    from effective_actor import EffectiveActor
    ep = np.load(os.path.join(G, "eval_%s.npz" % tag))
    z = np.load(os.path.join(G, "policy_td3_%s.npz" % tag))
    pis = [EffectiveActor(z, agent=i) for i in range(len(A))]
    n = 4     # four identical copies: also checks that envs do not interact
    env = vec_env.BatchedQuadEnv(n, framework=fw, dtype=torch.float64, goal_mode="traj0")
    env.set_state(np.tile(ep["state0"], (n, 1)), np.tile(ep["integ0"], (n, 1)), np.tile(ep["params"], (n, 1)),
                  np.tile(ep["goal0"], (n, 1)))
    obs = torch.as_tensor(np.tile(ep["obs0"], (n, 1)), device="cuda:0")
    ret = np.zeros(len(A))
    H = len(ep["reward"])
    worst_a = worst_s = 0.0
    for t in range(H):
        parts = [obs[:, :15], obs[:, 15:18]] if fw == "MODUL" else [obs]
        act = torch.cat([pi(o) for pi, o in zip(pis, parts)], dim=1).to(torch.float64)
        worst_a = max(worst_a, float((act[0].cpu() - torch.as_tensor(ep["action"][t])).abs().max()))
        o_n, rew, done, _, _ = env.step(act)
        obs = torch.cat(o_n, dim=1)
        ret += rew[0].cpu().numpy()
        assert not bool(done.any())
        if t % 100 == 99:
            worst_s = max(worst_s, float(np.abs(env.get_state()[0][0] - ep["state"][t]).max()))
    st = env.get_state()[0]
    assert np.abs(st - st[0]).max() == 0.0                     # identical envs stay identical
    assert worst_a < 1e-3 and worst_s < 1e-3, (worst_a, worst_s)
    assert np.abs(ret - ep["reward"].sum(axis=0)).max() < 0.05, (ret, ep["reward"].sum(axis=0))
    assert ret[0] > 985
    env.close()


@pytest.mark.gpu
@pytest.mark.parametrize("fw,tag", [("MONO", "mono"), ("MODUL", "modul")])
def test_compiled_actor_kernel_matches_reference_outputs(fw, tag):
    """qr_policy_td3 (effective weights compiled to straight-line sm_100a code) vs the reference actors' outputs."""
    from gym_rotor_b200 import vec_env
    z = np.load(os.path.join(G, "policy_td3_%s.npz" % tag))
    n = z["a0_obs"].shape[0]
    env = vec_env.BatchedQuadEnv(n, framework=fw, dtype=torch.float32)
    if fw == "MONO":
        obs = z["a0_obs"]; ref = z["a0_act"]
    else:
        obs = np.concatenate([z["a0_obs"], z["a1_obs"]], axis=1); ref = np.concatenate([z["a0_act"], z["a1_act"]], axis=1)
    env.obs.copy_(torch.as_tensor(obs, device="cuda:0"))
    act = env.policy_td3().cpu().numpy()
    assert act.shape == ref.shape and np.abs(act - ref).max() < 1e-4, np.abs(act - ref).max()
    env.close()


@pytest.mark.gpu
def test_compiled_actor_in_the_loop_flies_the_reference_episode():
    """Config 5 in miniature: env.step + compiled actor, both on device, 1000 steps against the reference's run."""
    from gym_rotor_b200 import vec_env
    ep = np.load(os.path.join(G, "eval_mono.npz"))
    n = 32
    env = vec_env.BatchedQuadEnv(n, framework="MONO", dtype=torch.float64, goal_mode="traj0")
    env.set_state(np.tile(ep["state0"], (n, 1)), np.tile(ep["integ0"], (n, 1)), np.tile(ep["params"], (n, 1)),
                  np.tile(ep["goal0"], (n, 1)))
    env.obs.copy_(torch.as_tensor(np.tile(ep["obs0"], (n, 1)), device="cuda:0"))
    ret = 0.0
    for t in range(len(ep["reward"])):
        act = env.policy_td3()
        _, rew, done, _, _ = env.step(act.to(torch.float64))
        ret += float(rew[0, 0])
        assert not bool(done.any())
    st = env.get_state()[0]
    assert np.abs(st[0] - ep["state"][-1]).max() < 1e-3
    assert abs(ret - ep["reward"].sum()) < 0.05 and ret > 985
    env.close()


def test_reference_attribute_surface_without_gymnasium():
    """Box stand-in (set_seed seeds both spaces, utils/utils.py:17-18) and forces_to_fM / fM_to_forces (quad.py:396-402)."""
    from gym_rotor_b200.vec_env import Box, forces_to_fM_matrices
    sp = Box(-1.0, 1.0, shape=(4,), dtype=np.float32)
    sp.seed(7); a = sp.sample(); sp.seed(7); b = sp.sample()
    assert a.dtype == np.float32 and a.shape == (4,) and np.array_equal(a, b) and sp.contains(a) and not sp.contains(a + 3)
    d, c = np.array([0.23, 0.25]), np.array([0.0135, 0.012])
    A, Ainv = forces_to_fM_matrices(d, c)
    for i in range(2):
        ref = np.array([[1.0, 1.0, 1.0, 1.0], [0.0, -d[i], 0.0, d[i]], [d[i], 0.0, -d[i], 0.0], [-c[i], c[i], -c[i], c[i]]])
        assert np.array_equal(A[i].numpy(), ref) and np.abs(Ainv[i].numpy() - np.linalg.inv(ref)).max() < 1e-12


def test_trainer_side_helpers_match_reference_formulas():
    """benchmark_reward (utils/utils.py:42-47 on get_error_state, :21-39) and the time-limit relabel (main.py:169-173)."""
    from gym_rotor_b200.vec_env import benchmark_reward, time_limit_relabel
    rng = np.random.default_rng(0)
    obs = rng.uniform(-0.05, 0.05, (64, 23)).astype(np.float32)
    obs[::4, 0:3] *= 0.1
    o_t = [torch.as_tensor(obs)]
    ex = obs[:, 0:3].astype(np.float64) * 1.0; eb1 = obs[:, 18].astype(np.float64) * np.pi
    ref = np.interp(-np.linalg.norm(ex, axis=1) - np.abs(eb1), [-2., 0.], [0., 1.])
    assert np.abs(benchmark_reward(o_t, "MONO").numpy() - ref).max() < 1e-12
    rew = torch.as_tensor(rng.uniform(0, 1, (64, 1))); rew[5, 0] = -1.0
    done = torch.zeros((64, 1), dtype=torch.bool)
    rel = time_limit_relabel(o_t, rew, done, "MONO").numpy()[:, 0]
    exp = (np.abs(ex) <= 0.03).all(axis=1) & (rew.numpy()[:, 0] != -1.0)
    assert (rel == exp).all() and rel.any() and not rel.all()
    o1 = torch.as_tensor(obs[:, :15]); o2 = torch.as_tensor(obs[:, 15:18])
    refm = np.interp(-np.linalg.norm(ex, axis=1) - np.abs(obs[:, 15].astype(np.float64) * np.pi), [-2., 0.], [0., 1.])
    assert np.abs(benchmark_reward([o1, o2], "MODUL").numpy() - refm).max() < 1e-12


def test_forces_from_fM_inverts_the_mixing_matrix():
    """(f, M) -> T1..T4 diagnostic (draw_plot.py:55-72): inverse of forces_to_fM (quad.py:396-401)."""
    from gym_rotor_b200.vec_env import forces_from_fM
    rng = np.random.default_rng(1)
    T = rng.uniform(0.5, 11.0, (50, 4)); d, c = 0.23, 0.0135
    A = np.array([[1, 1, 1, 1], [0, -d, 0, d], [d, 0, -d, 0], [-c, c, -c, c]], float)
    fM = T @ A.T
    out = forces_from_fM(fM[:, 0], fM[:, 1:4]).numpy()
    assert np.abs(out - T).max() < 1e-10
    clipped = forces_from_fM(fM[:, 0], fM[:, 1:4], min_force=2.0, max_force=8.0).numpy()
    assert clipped.min() >= 2.0 and clipped.max() <= 8.0


@pytest.mark.gpu
def test_flight_log_has_the_reference_layout(tmp_path):
    """FlightLog writes rows that reproduce the reference's own .dat (KAT-1) when fed the same flight."""
    from gym_rotor_b200 import vec_env
    rows = np.load(os.path.join(G, "kat1_modul_log.npz"))["rows"]
    env = vec_env.BatchedQuadEnv(2, framework="MODUL", dtype=torch.float64)
    par = np.tile(np.array([2.15, 0.23, 0.022, 0.035, 0.0135, 2.2]), (2, 1))
    log = vec_env.FlightLog(env, index=1)
    H = 40
    goal = np.tile(rows[0, 28:40], (2, 1))
    integ = np.zeros((2, 8)); integ[:, 0:3] = rows[0, 23:26]; integ[:, 6] = rows[0, 27]
    env.set_state(np.tile(rows[0, 5:23], (2, 1)), integ, par, goal)
    for t in range(H):
        act = torch.as_tensor(np.tile(rows[t, 0:5], (2, 1)), device="cuda:0")
        g = np.tile(rows[t, 28:40], (2, 1))
        env.set_state(None, None, None, g)
        # the reference logs the PRE-step state with the action it is about to apply (main.py:344-352)
        obs_pre = [env.obs[:, :15].clone(), env.obs[:, 15:18].clone()] if t else None
        if t:
            log.record(act, obs_pre)
        env.step(act)
    p = tmp_path / "log.dat"
    log.save(str(p))
    out = np.loadtxt(str(p))
    assert out.shape == (H - 1, 40)
    assert np.abs(out[:, 0:5] - rows[1:H, 0:5]).max() < 1e-9            # actions
    assert np.abs(out[:, 5:23] - rows[1:H, 5:23]).max() < 5e-9          # states follow the logged flight
    assert np.abs(out[:, 28:40] - rows[1:H, 28:40]).max() < 1e-9        # commands
    env.close()
