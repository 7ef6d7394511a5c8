"""Sharded four-step NTT over G ranks for large domains: one process per GPU, one all-to-all.

The 2^L-point transform is viewed as n1 × m (n = n1·m).  Two layouts of a distributed vector:

  column layout (coefficient side):  rank g holds  A_g[j1][c] = x[j1·m + g·cl + c]      (n1 × cl,  cl = m / G)
  row layout    (evaluation side):   rank g holds  B_g[r][k'] = X[(g·rl + r) + n1·k']   (rl × m,   rl = n1 / G)

  forward (column → row):  length-n1 NTT down the local columns fused with the twiddle ω_n^{col·k1}
                           → all-to-all of rl × cl blocks → block transpose → length-m NTT along each local row
  inverse (row → column):  the same steps backwards with ω⁻¹ and the n⁻¹ scale.

Elementwise work between transforms stays in either layout, so a prover never needs a third exchange to restore
natural order.  The local steps go through a backend object: `GpuBackend` calls the C ABI (pb200_ntt_columns_dev,
pb200_block_transpose_dev, pb200_ntt_dev) on this rank's GPU and NCCL for the exchange; tests plug in a CPU backend to
check the orchestration and the index maps over gloo.  This module is host plumbing — the arithmetic is in csrc/.
"""
import numpy as np


class ShardSpec:
    def __init__(self, log_n, world, log_n1=None):
        assert world & (world - 1) == 0, "power-of-two rank count"
        self.log_n, self.world = log_n, world
        log_g = world.bit_length() - 1
        if log_n1 is None:
            log_n1 = max(8, log_g)                      # short columns: one CTA pass; rows use the full single-GPU NTT
        assert log_g <= log_n1 <= 11 and log_n - log_n1 >= log_g
        self.log_n1 = log_n1
        self.n, self.n1, self.m = 1 << log_n, 1 << log_n1, 1 << (log_n - log_n1)
        self.cl, self.rl = self.m // world, self.n1 // world
        self.local = self.n // world

    # ---- index maps (global index → (rank, local offset)) -------------------------------------------------
    def column_layout(self, i):
        j1, jp = divmod(i, self.m)
        g, c = divmod(jp, self.cl)
        return g, j1 * self.cl + c

    def row_layout(self, k):
        kp, k1 = divmod(k, self.n1)
        g, r = divmod(k1, self.rl)
        return g, r * self.m + kp

    def scatter(self, vec, rank, layout):
        """This rank's shard of a full (n, 4) vector (host-side helper for tests and small inputs)."""
        v = np.asarray(vec).reshape(self.n, 4)
        if layout == "column":
            return np.ascontiguousarray(v.reshape(self.n1, self.world, self.cl, 4)[:, rank].reshape(self.local, 4))
        return np.ascontiguousarray(
            v.reshape(self.m, self.world, self.rl, 4)[:, rank].transpose(1, 0, 2).reshape(self.local, 4))

    def gather(self, shards, layout):
        """Inverse of scatter given every rank's shard."""
        out = np.empty((self.n, 4), np.uint64)
        if layout == "column":
            o = out.reshape(self.n1, self.world, self.cl, 4)
            for g, s in enumerate(shards):
                o[:, g] = np.asarray(s).reshape(self.n1, self.cl, 4)
        else:
            o = out.reshape(self.m, self.world, self.rl, 4)
            for g, s in enumerate(shards):
                o[:, g] = np.asarray(s).reshape(self.rl, self.m, 4).transpose(1, 0, 2)
        return out


class DistributedDomain:
    """EvaluationDomain for a vector sharded over `world` ranks (fft: column layout → row layout; ifft: back)."""

    def __init__(self, log_n, rank, world, backend, log_n1=None):
        self.spec = ShardSpec(log_n, world, log_n1)
        self.rank, self.backend = rank, backend

    def fft(self, buf, tmp):
        """buf, tmp: backend buffers of spec.local scalars.  Result (row layout) ends in `buf`."""
        s, be = self.spec, self.backend
        be.columns(buf, s.log_n, s.log_n1, s.cl, self.rank * s.cl, inverse=False)
        be.all_to_all(tmp, buf, s.world)                       # block h (rl × cl) → rank h
        be.block_transpose(buf, tmp, s.world, s.rl, s.cl)      # [G][rl][cl] → [rl][G][cl] = rl rows of length m
        be.rows(buf, s.rl, s.log_n - s.log_n1, inverse=False)
        return buf

    def fft_fused(self, buf, peers):
        """Forward transform with the exchange fused into the column kernel: every rank's column pass stores its
        outputs directly into the owners' row-layout buffers (`peers` = PeerBuffers: NVLink peer memory), then one
        barrier, then the local rows.  The result (row layout) is in peers.mine; `buf` is consumed."""
        s, be = self.spec, self.backend
        be.columns_scatter(buf, s.log_n, s.log_n1, s.cl, self.rank * s.cl, peers)
        be.barrier()                                           # every rank has finished writing into every buffer
        be.rows(peers.mine, s.rl, s.log_n - s.log_n1, inverse=False)
        return peers.mine

    def ifft(self, buf, tmp):
        s, be = self.spec, self.backend
        be.rows(buf, s.rl, s.log_n - s.log_n1, inverse=True)   # includes m⁻¹
        be.block_transpose(tmp, buf, s.rl, s.world, s.cl)      # [rl][G][cl] → [G][rl][cl]
        be.all_to_all(buf, tmp, s.world)                       # rows come home: [G·rl = n1][cl]
        be.columns(buf, s.log_n, s.log_n1, s.cl, self.rank * s.cl, inverse=True)   # includes n1⁻¹
        return buf


class GpuBackend:
    """Local steps on this rank's GPU through the C ABI; exchange through torch.distributed (NCCL)."""

    def __init__(self, ctx, dist, torch):
        self.ctx, self.dist, self.torch = ctx, dist, torch
        self.stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", ctx.device))

    def alloc(self, n_scalars):
        return self.torch.empty(n_scalars * 4, dtype=self.torch.int64, device=self.torch.device("cuda", self.ctx.device))

    def columns(self, buf, log_n, log_n1, cols, col_offset, inverse):
        self.ctx.ntt_columns_dev(buf.data_ptr(), log_n, log_n1, cols.bit_length() - 1, col_offset, inverse)

    def rows(self, buf, n_rows, log_m, inverse):
        self.ctx.ntt_batch_dev(buf.data_ptr(), log_m, n_rows, inverse, False)   # one launch per pass for all rows

    def block_transpose(self, dst, src, blocks, rows, cols):
        self.ctx.block_transpose_dev(dst.data_ptr(), src.data_ptr(), blocks, rows, cols)

    def all_to_all(self, out, inp, world):
        with self.torch.cuda.stream(self.stream):               # ordered after the kernels on the library's stream
            self.dist.all_to_all_single(out, inp)

    # ---- fused exchange over peer memory -----------------------------------------------------------------
    def columns_scatter(self, buf, log_n, log_n1, cols, col_offset, peers):
        self.ctx.ntt_columns_scatter_dev(_ptr_of(buf), log_n, log_n1, cols.bit_length() - 1, col_offset, peers.ptrs)

    def barrier(self):
        self.ctx.sync()
        if self.dist is not None:
            self.dist.barrier()


class _RawBuffer:
    """A pb200_malloc'ed device buffer with the data_ptr() interface the backend expects (IPC needs cudaMalloc memory)."""

    def __init__(self, ctx, nbytes):
        self.ctx, self.ptr, self.nbytes = ctx, ctx.malloc(nbytes), nbytes

    def data_ptr(self):
        return self.ptr


def _ptr_of(buf):
    return buf.data_ptr()


class PeerBuffers:
    """One row-layout receive buffer per rank, mapped into every other rank with CUDA IPC (NVLink peer memory)."""

    def __init__(self, ctx, dist, rank, world, n_local_scalars):
        self.ctx, self.rank, self.world = ctx, rank, world
        self.mine = _RawBuffer(ctx, n_local_scalars * 32)
        handles = [None] * world
        dist.all_gather_object(handles, ctx.ipc_export(self.mine.ptr))
        self._opened = []
        self.ptrs = []
        for h in range(world):
            if h == rank:
                self.ptrs.append(self.mine.ptr)
            else:
                p = ctx.ipc_open(handles[h])
                self._opened.append(p)
                self.ptrs.append(p)

    def close(self):
        for p in self._opened:
            self.ctx.ipc_close(p)
        self._opened = []
        self.ctx.free(self.mine.ptr)
