#!/usr/bin/env python
"""Extract the EFFECTIVE weights of the reference's shipped TD3 actors and record closed-loop golden episodes.

TEST INFRASTRUCTURE; runs only in the build container (needs /root/reference and its models/*.pth).

The shipped checkpoints are equivariant-MLP actors whose bilinear layers use index tables drawn with
torch.randint at construction time and NOT stored in the checkpoint (algos/emlp_torch/reps/representation.py:
374-376): the weights only mean something inside a module constructed exactly like main.py does
(set_seed(1992) at main.py:65, agents at main.py:85, test mode).  This script builds the actors that way, loads
the checkpoints and then treats every layer as a black box:

  * Linear           y = A x + b          -> A, b by probing with the zero vector and the unit vectors;
  * BiLinear         q(x) homogeneous quadratic -> symmetric tensor T with q_i = sum_jk T_ijk x_j x_k by polarisation;
  * GatedNonlinearity out_c = sigmoid(pre[g_c]) * pre[c] -> gate index list g (algos/emlp_torch/nn.py:262-280).

Outputs (tests/golden/):
  policy_td3_mono.npz, policy_td3_modul.npz : per agent `a{i}_n_blocks`, `a{i}_b{k}_A/b/T/gate`, `a{i}_out_A/out_b`,
      plus `a{i}_obs` / `a{i}_act` golden input/output pairs of the reference actor (float32);
  eval_mono.npz, eval_modul.npz : one reference evaluation episode (main.py:270-365 protocol, eval reset, 1000 steps):
      initial state/params, per-step goal, action, state, obs, reward.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def build_agents(framework):
    import torch
    rh._prepare_path()
    argv = sys.argv
    sys.argv = ["x", "--framework", framework, "--test_model", "True"]
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp()
    os.symlink(os.path.join(rh.REFERENCE_ROOT, "models"), os.path.join(tmp, "models"))
    os.chdir(tmp)
    try:
        import args_parse
        from utils.utils import set_seed
        from utils.trajectory_generator import TrajectoryGenerator
        from gym_rotor.wrappers.coupled_yaw_wrapper import CoupledWrapper
        from gym_rotor.wrappers.decoupled_yaw_wrapper import DecoupledWrapper
        from algos.td3.td3 import TD3
        args = args_parse.create_parser().parse_args()
        args.device = torch.device("cpu")
        if framework == "MODUL":
            env = DecoupledWrapper(); args.N = 2; args.obs_dim_n = [15, 3]; args.action_dim_n = [4, 1]
        else:
            env = CoupledWrapper(); args.N = 1; args.obs_dim_n = [23]; args.action_dim_n = [4]
        set_seed(env, args.seed)                       # main.py:65
        TrajectoryGenerator(env)                       # main.py:80
        agents = [TD3(args, i) for i in range(args.N)]  # main.py:85
        if framework == "MODUL":                       # main.py:101-110
            agents[0].load("TD3", "MODUL", 564_000, 0, args.seed)
            agents[1].load("TD3", "MODUL", 850_000, 1, args.seed)
        else:
            agents[0].load("TD3", "MONO", 700_000, 0, args.seed)
    finally:
        os.chdir(cwd)
        sys.argv = argv
    return args, agents


def probe_affine(mod, n_in):
    import torch
    with torch.no_grad():   # float32: the reference's equivariant projectors are float32 operators
        b = mod(torch.zeros(1, n_in))[0]
        A = (mod(torch.eye(n_in)) - b).T            # column j = f(e_j) - b
    return A.numpy().astype(np.float64), b.numpy().astype(np.float64)


def probe_quadratic(mod, n):
    import torch
    with torch.no_grad():
        E = torch.eye(n, dtype=torch.float64)
        mod = mod.double()
        q1 = mod(E)                                  # q(e_j)            [n, n_out]
        T = torch.zeros(q1.shape[1], n, n, dtype=torch.float64)
        for j in range(n):
            T[:, j, j] = q1[j]
            for k in range(j + 1, n):
                s = mod((E[j] + E[k])[None])[0] - q1[j] - q1[k]   # T_ijk + T_ikj
                T[:, j, k] = s / 2
                T[:, k, j] = s / 2
        # homogeneity check: q(2x) = 4 q(x)
        x = torch.randn(5, n, dtype=torch.float64)
        assert torch.allclose(mod(2 * x), 4 * mod(x), atol=1e-9)
        assert torch.allclose(torch.einsum("ijk,nj,nk->ni", T, x, x), mod(x), atol=1e-9)
        mod.float()
    return T.numpy()


def extract_actor(actor):
    from algos.emlp_torch.nn import gate_indices
    out = {}
    blocks = list(actor.network)
    out["n_blocks"] = np.int32(len(blocks) - 1)
    n_in = None
    for k, blk in enumerate(blocks[:-1]):
        n_in = blk.linear.weight.shape[1]
        A, b = probe_affine(blk.linear, n_in)
        T = probe_quadratic(blk.bilinear, A.shape[0])
        g = np.asarray(gate_indices(blk.nonlinearity.rep), dtype=np.int64)
        out["b%d_A" % k], out["b%d_b" % k], out["b%d_T" % k], out["b%d_gate" % k] = A, b, T, g
    fin = blocks[-1]
    out["out_A"], out["out_b"] = probe_affine(fin, fin.weight.shape[1])
    return out


def eval_episode(framework, agents, seed=1992, steps=1000):
    """One evaluation episode driven exactly like main.py:304-365 (eval reset, trajectory mode 0, no exploration noise)."""
    env = rh.make_env(framework)
    tg = rh.make_trajgen(env)
    rh.seed_all(seed)
    state32 = env.reset(env_type="eval", seed=seed)
    tg.mark_traj_start(state32)
    xd, vd, b1d, b1d_dot, Wd = tg.get_desired(state32, 0)
    env.set_goal_state(xd, vd, b1d, b1d_dot, Wd)
    obs_n = env.get_norm_error_state(framework)
    rec = {k: [] for k in ("goal", "action", "state", "obs", "reward", "done")}
    out = {"state0": np.array(env.state, np.float64), "params": rh.get_params(env), "goal0": rh.get_goal(env),
           "obs0": np.concatenate(obs_n).astype(np.float32), "integ0": rh.get_integ(env)}
    for _ in range(steps):
        st = env.get_current_state()
        xd, vd, b1d, b1d_dot, Wd = tg.get_desired(st, 0)
        env.set_goal_state(xd, vd, b1d, b1d_dot, Wd)
        act_n = [ag.choose_action(o, explor_noise_std=0.) for ag, o in zip(agents, obs_n)]
        action = np.concatenate(act_n, axis=None)
        rec["goal"].append(rh.get_goal(env)); rec["action"].append(action.astype(np.float64))
        obs_n, rew, done, _, _ = env.step(action.copy())
        rec["state"].append(np.array(env.state, np.float64)); rec["obs"].append(np.concatenate(obs_n).astype(np.float32))
        rec["reward"].append(np.array(rew, np.float64)); rec["done"].append(np.array(done, bool))
        if any(done):
            break
    out.update({k: np.array(v) for k, v in rec.items()})
    return out


def main():
    import torch
    for framework, tag in (("MONO", "mono"), ("MODUL", "modul")):
        args, agents = build_agents(framework)
        fix = {}
        rng = np.random.default_rng(3)
        for i, ag in enumerate(agents):
            ag.actor.eval()
            for k, v in extract_actor(ag.actor).items():
                fix["a%d_%s" % (i, k)] = v
            n_in = args.obs_dim_n[i]
            obs = rng.uniform(-1, 1, (512, n_in)).astype(np.float32)
            with torch.no_grad():
                act = ag.actor(torch.tensor(obs)).numpy()
            fix["a%d_obs" % i], fix["a%d_act" % i] = obs, act
        path = os.path.join(OUT, "policy_td3_%s.npz" % tag)
        np.savez_compressed(path, **fix)
        print("wrote", path, "%.0f KB" % (os.path.getsize(path) / 1024))
        ep = eval_episode(framework, agents)
        path = os.path.join(OUT, "eval_%s.npz" % tag)
        np.savez_compressed(path, **ep)
        print("wrote", path, "%.0f KB" % (os.path.getsize(path) / 1024), "steps", len(ep["reward"]),
              "return", ep["reward"].sum(axis=0))


if __name__ == "__main__":
    main()
