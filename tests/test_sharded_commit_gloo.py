"""World-size-2 CPU test (gloo) of the host logic behind point-range-sharded proving (pb200_preprocess_sharded /
pb200_prove on a sharded key): each rank commits its coefficient slice against its slice of the commit key, the
144-byte partial sums travel through `torch_allgather` — the very callback the C ABI invokes — and their sum must be
the commitment of the whole polynomial.  The local MSM is the oracle here (no GPU in this container); on the GPU the
same exchange is driven from csrc/plonk.cu and covered by scripts/dist_prove_check.py."""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def _worker(rank, world, port, n, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import pyoracle as O
    import plonk_prototype_b200 as pb
    dist.init_process_group("gloo", rank=rank, world_size=world)
    gather = pb.torch_allgather(dist)                       # device=None: host tensors over gloo
    pts = O.synthetic_bases(n)                              # stands in for powers_of_g
    coeffs = O.fr_to_mont(O.random_fr(0xC0117, n))
    per = n // world
    lo = rank * per
    batch = 3                                               # a batched commit sends batch × 144 bytes per rank
    partial = np.stack([O.msm_variable_base(pts[lo:lo + per], np.roll(coeffs, j, axis=0)[lo:lo + per]) for j in range(batch)])
    out = gather(partial.tobytes())
    assert len(out) == world * batch * 144
    allp = np.frombuffer(out, dtype=np.uint64).reshape(world, batch, 18)
    assert (allp[rank] == partial).all()                    # rank-major layout
    if rank == 0:
        ret["parts"] = allp.tolist()
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_commit_adds_up():
    import model
    import pyoracle as O
    n, world, port = 1 << 10, 2, 29600 + os.getpid() % 300
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, n, ret), nprocs=world, join=True)
        parts = np.array(ret["parts"], dtype=np.uint64)
    pts = O.synthetic_bases(n)
    coeffs = O.fr_to_mont(O.random_fr(0xC0117, n))
    for j in range(3):
        total = None
        for r in range(world):
            total = model.g1_add(total, O.g1_proj_to_affine_canonical(parts[r, j]))
        want = O.g1_proj_to_affine_canonical(O.msm_variable_base(pts, np.roll(coeffs, j, axis=0), threads=4))
        assert total == want
