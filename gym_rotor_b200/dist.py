"""Multi-GPU plumbing: contiguous env-index sharding and the one collective of the path.

Envs never interact: GPU g owns global envs [g*N/G, (g+1)*N/G) and the Philox streams are
keyed by the GLOBAL env id, so results do not depend on G.  The only traffic is a sum all-reduce of the
20-double episode-statistics vector once per rollout (160 bytes: latency bound, NCCL over NVLink).
"""
import numpy as np
import torch
import torch.distributed as dist

from ._native import NUM_STATS


def shard_range(n_total, rank, world):
    """Contiguous block of global env ids owned by `rank` (sizes differ by at most one)."""
    if not (0 <= rank < world) or n_total < 0:
        raise ValueError("bad shard request")
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_stats_tensor(stats, group=None):
    """Sum-all-reduces a [NUM_STATS] float64 tensor in place (works on NCCL/CUDA and gloo/CPU tensors)."""
    if stats.numel() != NUM_STATS or stats.dtype != torch.float64:
        raise ValueError("stats must be a float64 tensor with %d elements" % NUM_STATS)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def allreduce_stats(env, device, reset=True, group=None):
    """Global episode statistics: all-reduce the device accumulators, then read them (and zero the local ones)."""
    buf = env.stats_dev.clone()
    if reset:
        env.stats_dev.zero_()
    reduce_stats_tensor(buf, group)
    return buf.cpu().numpy()


def summarize(stats):
    """Human-readable view of the statistics vector (indices: include/quadrotor_b200.h QR_STAT_*)."""
    s = np.asarray(stats, dtype=np.float64)
    ep = max(s[0], 1.0)
    mean_ret = s[1] / ep
    return {"episodes": s[0], "mean_return_agent0": mean_ret, "mean_return_agent1": s[2] / ep,
            "std_return_agent0": float(np.sqrt(max(s[6] / ep - mean_ret ** 2, 0.0))),
            "mean_episode_length": s[3] / ep, "crashed": s[4], "truncated": s[5], "steps": s[7],
            "bad_status": s[8], "mean_rhs_evals": s[9] / max(s[7], 1.0),
            "attempt_hist": (s[10:14] / max(s[7], 1.0)).tolist(), "mean_reward_agent0": s[14] / max(s[7], 1.0),
            "so3_projections_per_step": s[15] / max(s[7], 1.0),
            "mean_benchmark_reward": s[16] / max(s[7], 1.0),      # utils/utils.py:21-47, per env-step
            "solved_at_time_limit": s[17]}                        # main.py:169-173
