"""CPU-only, world_size 2 over gloo: the host-side multi-GPU logic (sharding + the statistics all-reduce)."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from gym_rotor_b200._native import NUM_STATS  # noqa: E402


def test_shard_range_partitions_exactly():
    from gym_rotor_b200.dist import shard_range
    for n in (0, 1, 7, 4096, (1 << 24) + 5):
        for w in (1, 2, 3, 8):
            blocks = [shard_range(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gym_rotor_b200.dist import reduce_stats_tensor, shard_range, summarize
    lo, hi = shard_range(1000, rank, world)
    s = torch.zeros(NUM_STATS, dtype=torch.float64)
    s[0] = hi - lo                 # episodes
    s[1] = float(sum(range(lo, hi)))   # sum of returns
    s[3] = 10.0 * (hi - lo)
    s[7] = 100.0 * (hi - lo)
    reduce_stats_tensor(s)
    summ = summarize(s.numpy())
    out[rank] = (s.numpy().copy(), summ["mean_return_agent0"], summ["mean_episode_length"])
    dist.destroy_process_group()


def test_stats_allreduce_world2():
    world, port = 2, _free_port()
    mgr = mp.Manager(); out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    a, b = out[0], out[1]
    assert np.array_equal(a[0], b[0])
    assert a[0][0] == 1000 and a[0][1] == sum(range(1000)) and a[0][7] == 100000
    assert abs(a[1] - 499.5) < 1e-12 and a[2] == 10.0


def test_reduce_stats_rejects_wrong_shape():
    from gym_rotor_b200.dist import reduce_stats_tensor
    with pytest.raises(ValueError):
        reduce_stats_tensor(torch.zeros(8, dtype=torch.float64))
    s = torch.arange(NUM_STATS, dtype=torch.float64)
    assert torch.equal(reduce_stats_tensor(s.clone()), s)   # no process group: identity
