"""World-size-2 CPU test (gloo) of the N>1 host logic in bench.py: point-range shard parameters, the
all-gather of 144-byte partial results, and that per-shard closed forms add up to the closed form of the
whole range.  The group arithmetic itself runs on the GPU and is covered by tests/test_msm_gpu.py."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, n, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import bench
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sp = bench.shard_params(rank, n)
    s = bench.random_fr_limbs(sp["scalar_seed"], n)
    k_local = bench.closed_form_scalar(s, sp["a"], sp["d"])
    ks = [None] * world
    dist.all_gather_object(ks, k_local)
    fake_partial = np.arange(18, dtype=np.uint64) + np.uint64(1000 * rank) + np.uint64(1 << 63)  # exercises the sign bit
    parts = bench.gather_partials(dist, fake_partial, world, torch.device("cpu"))
    if rank == 0:
        ret["ks"] = ks
        ret["parts"] = parts.tolist()
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_adds_up():
    import bench
    n, world, port = 1 << 12, 2, 29511 + os.getpid() % 500
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, n, ret), nprocs=world, join=True)
        ks, parts = ret["ks"], np.array(ret["parts"], dtype=np.uint64)
    # the whole range [0, 2n) with both scalar streams concatenated
    whole = np.concatenate([bench.random_fr_limbs(bench.shard_params(r, n)["scalar_seed"], n) for r in range(world)])
    assert sum(ks) % bench.R_MOD == bench.closed_form_scalar(whole, bench.A0, bench.D0)
    for r in range(world):
        assert (parts[r] == np.arange(18, dtype=np.uint64) + np.uint64(1000 * r) + np.uint64(1 << 63)).all()
    # shards tile the index range without gaps: a_{r+1} = a_r + n·d
    assert bench.shard_params(1, n)["a"] - bench.shard_params(0, n)["a"] == n * bench.D0
