"""CPU: the N > 1 sharding path with a world_size-2 gloo process group."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_bounds_cover_everything():
    from chromevol_enhanced_b200.distributed import shard_bounds
    for total in (0, 1, 7, 512, 4096, 4099):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from chromevol_enhanced_b200 import distributed as D
    D.init_from_env(backend="gloo")
    rows = torch.arange(total * 9, dtype=torch.float64).reshape(total, 9)

    def fake_eval(r):                       # stands in for LikelihoodEngine.log_likelihood
        return -(r[:, 0] * 2.0 + r[:, 7])

    full = D.sharded_log_likelihood(fake_eval, rows)
    hist = torch.full((4, 8), rank + 1, dtype=torch.int64)
    D.sum_across_ranks(hist)
    mx = D.max_across_ranks(10.0 + rank, torch.device("cpu"))
    np.save(os.path.join(out_dir, f"r{rank}.npy"), full.numpy())
    assert int(hist[0, 0]) == sum(range(1, world + 1)) and mx == 10.0 + world - 1
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [7, 64])
def test_sharded_evaluation_gathers_on_every_rank(tmp_path, total):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), total, str(tmp_path)), nprocs=world, join=True)
    rows = np.arange(total * 9, dtype=np.float64).reshape(total, 9)
    want = -(rows[:, 0] * 2.0 + rows[:, 7])
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"r{r}.npy"), want)
