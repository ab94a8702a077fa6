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
    # pipelined gather: collect() returns the vector of the PREVIOUS submit
    lo, hi = D.shard_bounds(total, rank, world)
    pipe = D.GatherPipeline(total, world)
    assert pipe.collect() is None
    pipe.submit(fake_eval(rows[lo:hi]))
    first = pipe.collect()
    pipe.submit(2.0 * fake_eval(rows[lo:hi]))
    second = pipe.collect()
    assert torch.equal(first, full) and torch.equal(second, 2.0 * full) and pipe.collect() is None
    # fewer vectors than ranks: the empty shard still joins the collective
    tiny = D.sharded_log_likelihood(fake_eval, rows[:1])
    assert tiny.shape == (1,) and tiny[0] == full[0]
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


def _selection_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import json
    from chromevol_enhanced_b200 import distributed as D
    D.init_from_env(backend="gloo")
    configs = D.chromevol_model_configurations()

    def fake_select(cfgs):                  # stands in for calculate_model_selection_criteria
        return {c['name']: {'AIC': 2.0 * len(c['params_to_optimize']), 'rank': rank} for c in cfgs}

    merged = D.sharded_model_selection(fake_select, configs)
    with open(os.path.join(out_dir, f"sel{rank}.json"), "w") as fh:
        json.dump(merged, fh)
    dist.destroy_process_group()


def test_model_selection_candidates_are_sharded_round_robin(tmp_path):
    import json
    from chromevol_enhanced_b200 import distributed as D
    world = 2
    mp.spawn(_selection_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    configs = D.chromevol_model_configurations()
    names = [c['name'] for c in configs]
    assert len(names) == len(set(names)) == 8
    for r in range(world):
        merged = json.load(open(tmp_path / f"sel{r}.json"))
        assert list(merged) == names                                 # candidate order, every rank
        for k, c in enumerate(configs):
            assert merged[c['name']]['rank'] == k % world
            assert merged[c['name']]['AIC'] == 2.0 * len(c['params_to_optimize'])
    # single process: no collective
    solo = D.sharded_model_selection(lambda cfgs: {c['name']: 1 for c in cfgs}, configs)
    assert list(solo) == names
