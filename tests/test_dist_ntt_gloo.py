"""World-size-2 and -4 CPU tests (gloo) of the sharded four-step NTT orchestration (plonk-prototype_b200/dist_ntt.py):
layouts, the all-to-all, the block transposes and the step order, with the local column/row transforms supplied by
the CPU oracle instead of the GPU kernels.  Result must equal the single-vector EvaluationDomain::fft / ifft."""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


class OracleBackend:
    """Test double for GpuBackend: numpy buffers, oracle transforms, gloo point-to-point exchange."""

    def __init__(self, O, rank, world):
        self.O, self.rank, self.world = O, rank, world
        self.model = __import__("model")

    def columns(self, buf, log_n, log_n1, cols, col_offset, inverse):
        O, M = self.O, self.model
        n1 = 1 << log_n1
        a = buf.reshape(n1, cols, 4)
        w = M.domain(1 << log_n)["group_gen_inv" if inverse else "group_gen"]
        tw = np.empty((n1, cols, 4), np.uint64)
        for c in range(cols):
            base = pow(w, col_offset + c, M.R)
            vals, cur = [], 1
            for k in range(n1):
                vals.append(cur)
                cur = cur * base % M.R
            tw[:, c] = O.fr_to_mont(O.ints_to_limbs(vals, 4))
        if inverse:
            a[:] = O.fr_mul(a.reshape(-1, 4), tw.reshape(-1, 4)).reshape(n1, cols, 4)
        for c in range(cols):
            a[:, c] = O.ntt(np.ascontiguousarray(a[:, c]), inverse, False)     # oracle ifft includes n1^-1
        if not inverse:
            a[:] = O.fr_mul(a.reshape(-1, 4), tw.reshape(-1, 4)).reshape(n1, cols, 4)

    def rows(self, buf, n_rows, log_m, inverse):
        a = buf.reshape(n_rows, 1 << log_m, 4)
        for r in range(n_rows):
            a[r] = self.O.ntt(np.ascontiguousarray(a[r]), inverse, False)

    def block_transpose(self, dst, src, blocks, rows, cols):
        dst.reshape(rows, blocks, cols, 4)[:] = src.reshape(blocks, rows, cols, 4).transpose(1, 0, 2, 3)

    def all_to_all(self, out, inp, world):
        chunk = inp.size // world
        i64_in = torch.from_numpy(inp.view(np.int64).reshape(world, chunk).copy())
        outs = [torch.empty(chunk, dtype=torch.int64) for _ in range(world)]
        reqs = []
        for peer in range(world):
            if peer == self.rank:
                outs[peer].copy_(i64_in[peer])
            else:
                reqs.append(dist.isend(i64_in[peer].contiguous(), peer))
                reqs.append(dist.irecv(outs[peer], peer))
        for r in reqs:
            r.wait()
        out.reshape(-1)[:] = torch.cat(outs).numpy().view(np.uint64)


def _worker(rank, world, port, log_n, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import pyoracle as O
    import plonk_prototype_b200 as pb
    spec = pb.ShardSpec(log_n, world, log_n1=5)
    x = O.fr_to_mont(O.random_fr(0xD157 + log_n, 1 << log_n))
    dom = pb.DistributedDomain(log_n, rank, world, OracleBackend(O, rank, world), log_n1=5)
    buf = spec.scatter(x, rank, "column").copy()
    tmp = np.empty_like(buf)
    dom.fft(buf, tmp)
    ret["fft%d" % rank] = buf.copy()
    dom.ifft(buf, tmp)
    ret["back%d" % rank] = buf.copy()
    dist.barrier()
    dist.destroy_process_group()


def _run(world, log_n):
    import pyoracle as O
    import plonk_prototype_b200 as pb
    port = 29700 + (os.getpid() * 7 + world) % 200
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, log_n, ret), nprocs=world, join=True)
        spec = pb.ShardSpec(log_n, world, log_n1=5)
        x = O.fr_to_mont(O.random_fr(0xD157 + log_n, 1 << log_n))
        got = spec.gather([ret["fft%d" % r] for r in range(world)], "row")
        assert (got == O.ntt(x, 0, 0)).all()
        back = spec.gather([ret["back%d" % r] for r in range(world)], "column")
        assert (back == x).all()


def test_two_ranks():
    _run(2, 10)


def test_four_ranks():
    _run(4, 11)


def test_layout_maps_are_consistent():
    import plonk_prototype_b200 as pb
    spec = pb.ShardSpec(10, 4, log_n1=4)
    v = np.arange(spec.n * 4, dtype=np.uint64).reshape(spec.n, 4)
    for layout, fn in (("column", spec.column_layout), ("row", spec.row_layout)):
        shards = [spec.scatter(v, r, layout) for r in range(4)]
        assert (spec.gather(shards, layout) == v).all()
        for i in (0, 1, 63, 64, 517, spec.n - 1):
            g, off = fn(i)
            assert (shards[g][off] == v[i]).all(), (layout, i)


def test_sharded_round3_index_formulas_match_shard_spec():
    """csrc/plonk.cu's sharded round 3 addresses its shards with closed formulas — column shard element e = j1·cl + c holds
    coefficient j1·m + rank·cl + c (dist_load_kernel), row shard element e = r·m + k' is evaluation point
    (rank·rl + r) + n1·k' (quotient_kernel, dist = 1).  They must be the maps of ShardSpec, which the transforms implement."""
    import random
    import plonk_prototype_b200 as pb
    rng = random.Random(7)
    for log_n, world in ((14, 2), (16, 4), (22, 8), (26, 8)):
        spec = pb.ShardSpec(log_n, world, 8)
        log_m, log_g = log_n - 8, world.bit_length() - 1
        log_cl, log_rl = log_m - log_g, 8 - log_g
        for _ in range(200):
            rank, e = rng.randrange(world), rng.randrange(spec.local)
            j1, c = e >> log_cl, e & ((1 << log_cl) - 1)
            j = (j1 << log_m) + (rank << log_cl) + c
            assert spec.column_layout(j) == (rank, e)
            r, kq = e >> log_m, e & ((1 << log_m) - 1)
            i = ((rank << log_rl) + r) + (kq << 8)
            assert spec.row_layout(i) == (rank, e)
