"""GPU tests of the sharded four-step NTT building blocks (pb200_ntt_columns_dev, pb200_block_transpose_dev,
pb200_ntt_dev per row).  The G ranks are emulated on ONE GPU — every rank's steps run one after the other and the
all-to-all is a block copy — so the kernels are exercised for multi-rank shapes without needing several devices
(the real NCCL exchange is covered by scripts/dist_ntt_check.py under torchrun and by the gloo CPU test)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class EmulatedRanks:
    """Runs DistributedDomain for every rank on one device; all_to_all waits until all ranks posted."""

    def __init__(self, ctx, world):
        import plonk_prototype_b200 as pb
        self.ctx, self.world = ctx, world
        self.be = pb.GpuBackend(ctx, None, torch)

    def run(self, log_n, shards, inverse, log_n1=None):
        import plonk_prototype_b200 as pb
        G = self.world
        spec = pb.ShardSpec(log_n, G, log_n1)
        bufs = [torch.from_numpy(s.view(np.int64).reshape(-1).copy()).cuda() for s in shards]
        tmps = [torch.empty_like(b) for b in bufs]
        be = self.be
        sync = lambda: (self.ctx.sync(), torch.cuda.synchronize())
        chunk = spec.local * 4 // G

        def exchange(dst, src):
            sync()
            for g in range(G):
                for h in range(G):
                    dst[h][g * chunk:(g + 1) * chunk] = src[g][h * chunk:(h + 1) * chunk]
            sync()

        if not inverse:
            for g in range(G):
                be.columns(bufs[g], spec.log_n, spec.log_n1, spec.cl, g * spec.cl, False)
            exchange(tmps, bufs)
            for g in range(G):
                be.block_transpose(bufs[g], tmps[g], G, spec.rl, spec.cl)
                be.rows(bufs[g], spec.rl, spec.log_n - spec.log_n1, False)
        else:
            for g in range(G):
                be.rows(bufs[g], spec.rl, spec.log_n - spec.log_n1, True)
                be.block_transpose(tmps[g], bufs[g], spec.rl, G, spec.cl)
            exchange(bufs, tmps)
            for g in range(G):
                be.columns(bufs[g], spec.log_n, spec.log_n1, spec.cl, g * spec.cl, True)
        sync()
        return spec, [b.cpu().numpy().view(np.uint64).reshape(-1, 4) for b in bufs]


@pytest.mark.parametrize("world,log_n,log_n1", [(1, 12, 8), (2, 14, 8), (4, 16, 8), (8, 18, 9), (8, 20, 11), (2, 13, 3)])
def test_emulated_ranks_match_single_gpu_and_oracle(ctx, oracle, world, log_n, log_n1):
    import plonk_prototype_b200 as pb
    x = oracle.fr_to_mont(oracle.random_fr(0xD157 + log_n, 1 << log_n))
    spec = pb.ShardSpec(log_n, world, log_n1)
    em = EmulatedRanks(ctx, world)
    _, out = em.run(log_n, [spec.scatter(x, g, "column") for g in range(world)], False, log_n1)
    got = spec.gather(out, "row")
    want = x.copy()
    ctx.ntt(want, log_n, False, False)                       # single-GPU transform
    assert (got == want).all()
    if log_n <= 16:
        assert (got == oracle.ntt(x, 0, 0, threads=8)).all()
    _, back = em.run(log_n, [spec.scatter(got, g, "row") for g in range(world)], True, log_n1)
    assert (spec.gather(back, "column") == x).all()          # ifft(fft(x)) = x through the sharded path


@pytest.mark.parametrize("world,log_n,log_n1", [(2, 14, 8), (4, 16, 8), (8, 18, 9)])
def test_fused_column_scatter_writes_row_layout(ctx, oracle, world, log_n, log_n1):
    """pb200_ntt_columns_scatter_dev: the column pass stores straight into every owner's row-layout buffer (here the
    'peers' are other buffers on the same GPU); after the local rows the result equals the single-GPU transform."""
    import plonk_prototype_b200 as pb
    x = oracle.fr_to_mont(oracle.random_fr(0xF05E + log_n, 1 << log_n))
    spec = pb.ShardSpec(log_n, world, log_n1)
    cols = [torch.from_numpy(spec.scatter(x, g, "column").view(np.int64).reshape(-1).copy()).cuda() for g in range(world)]
    rows = [torch.zeros(spec.local * 4, dtype=torch.int64, device="cuda") for _ in range(world)]
    ptrs = [r.data_ptr() for r in rows]
    for g in range(world):
        ctx.ntt_columns_scatter_dev(cols[g].data_ptr(), log_n, log_n1, spec.cl.bit_length() - 1, g * spec.cl, ptrs)
    ctx.sync()
    for g in range(world):
        ctx.ntt_batch_dev(rows[g].data_ptr(), log_n - log_n1, spec.rl, False, False)
    ctx.sync()
    got = spec.gather([r.cpu().numpy().view(np.uint64).reshape(-1, 4) for r in rows], "row")
    want = x.copy()
    ctx.ntt(want, log_n, False, False)
    assert (got == want).all()
