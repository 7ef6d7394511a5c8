"""World-size-2 CPU test (gloo) of the N>1 host logic in bench.py: point-range shard parameters (strong scaling: every rank count splits the same global problem), the
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


def _worker(rank, world, port, log_total, ret):
    n = (1 << log_total) // world
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import bench
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s = bench.msm_scalar_slice(log_total, rank, world, np.empty((n, 4), np.uint64))   # this rank's slice of the ONE global problem
    k_local = bench.closed_form_scalar(s, bench.msm_shard_base(rank, n), bench.D0)
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
    log_total, world, port = 13, 2, 29511 + os.getpid() % 500
    n = (1 << log_total) // world
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, log_total, ret), nprocs=world, join=True)
        ks, parts = ret["ks"], np.array(ret["parts"], dtype=np.uint64)
    # strong scaling: the ranks' slices are the single-GPU problem — same global scalar vector, same bases
    whole = bench.msm_scalar_slice(log_total, 0, 1, np.empty((1 << log_total, 4), np.uint64))
    assert sum(ks) % bench.R_MOD == bench.closed_form_scalar(whole, bench.A0, bench.D0)
    for r in range(world):
        assert (parts[r] == np.arange(18, dtype=np.uint64) + np.uint64(1000 * r) + np.uint64(1 << 63)).all()
    # shards tile the index range without gaps: a_{r+1} = a_r + n·d
    assert bench.msm_shard_base(1, n) - bench.msm_shard_base(0, n) == n * bench.D0
