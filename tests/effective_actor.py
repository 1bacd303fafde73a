"""Fixed-weight inference of the reference's shipped TD3 actors (BASELINE config 5).

The shipped checkpoints (models/TD3_*.pth) are equivariant-MLP actors (algos/td3/td3_emlp.py:14-62,139-245;
algos/emlp_torch/nn.py:13-99) that only work inside a module built under torch.manual_seed(1992).  Their
EFFECTIVE per-layer maps were extracted once from a reference-constructed module by black-box probing
(the fixture script make_policy_fixture.py -> tests/golden/policy_td3_*.npz):

    block:  lin = A x + b ;  pre = lin + q(lin),  q_i = sum_jk T_ijk lin_j lin_k ;  h = sigmoid(pre[gate]) * pre[:C]
    head:   action = tanh(A_out h + b_out)

TEST INFRASTRUCTURE (it lives under tests/ on purpose): a torch restatement of those maps that the tests compare the
compiled actor kernels against (qr_policy_td3 / qr_rollout with QR_ACT_POLICY, csrc/generated/actor_td3.cuh).  The
package itself never evaluates a policy with torch/cuBLAS.
"""
import numpy as np
import torch


class EffectiveActor:
    def __init__(self, npz, agent=0, device="cuda:0", dtype=torch.float32):
        z = np.load(npz) if isinstance(npz, str) else npz
        p = "a%d_" % agent
        self.blocks = []
        for k in range(int(z[p + "n_blocks"])):
            A = torch.as_tensor(z[p + "b%d_A" % k], dtype=dtype, device=device)
            b = torch.as_tensor(z[p + "b%d_b" % k], dtype=dtype, device=device)
            T = torch.as_tensor(z[p + "b%d_T" % k], dtype=dtype, device=device)
            g = torch.as_tensor(z[p + "b%d_gate" % k], dtype=torch.long, device=device)
            # q(lin) as one matmul: (lin (x) lin) [n, C*C] @ T^T [C*C, C]; only the non-zero (j,k) columns are kept
            C = T.shape[1]
            Tm = T.reshape(T.shape[0], C * C)
            nz = (Tm != 0).any(dim=0).nonzero()[:, 0]
            self.blocks.append((A, b, Tm[:, nz].t().contiguous(), nz // C, nz % C, g, int(g.numel())))
        self.A_out = torch.as_tensor(z[p + "out_A"], dtype=dtype, device=device)
        self.b_out = torch.as_tensor(z[p + "out_b"], dtype=dtype, device=device)
        self.dtype = dtype

    @torch.no_grad()
    def __call__(self, obs):
        h = obs.to(self.dtype)
        for A, b, Tq, jj, kk, g, C in self.blocks:
            lin = torch.addmm(b, h, A.t())
            pre = lin + (lin[:, jj] * lin[:, kk]) @ Tq
            h = torch.sigmoid(pre[:, g]) * pre[:, :C]
        return torch.tanh(torch.addmm(self.b_out, h, self.A_out.t()))
