"""CPU model of ntt_pass_tma_kernel's data movement (csrc/ntt_tma.cuh): the tile geometry the tensor maps describe, the
SWIZZLE_128B shared-memory layout, the thread → point assignment of every radix-8 round, the compact twiddle image,
the bit-reversed write-back and the tensor store — replayed with Python integers and compared with the definition of
the transform.  The index arithmetic below is the kernel's, line by line (same names), so a layout or index mistake
shows up here, on a machine without a GPU, before any GPU time is spent.  (The arithmetic itself — Montgomery limbs —
is covered by tests/test_host_emulation.py and the GPU parity tests.)"""
import pytest

import model

R = model.R


def brev(x, bits):
    return int(format(x, "0%db" % bits)[::-1], 2) if bits else 0


def tile_off(x, c):
    """byte offset of the low half of scalar (x, c); the high half is at offset ^ 16 (csrc/ntt_tma.cuh tile_off)."""
    return (x << 7) + (((c << 1) ^ (x & 7)) << 4)


class Tile:
    """Shared-memory tile as 16-byte chunks addressed by byte offset."""

    def __init__(self, S):
        self.chunks = {}
        self.S = S

    def tma_load_row(self, x, four_scalars):
        """What cp.async.bulk.tensor with SWIZZLE_128B writes for box row x: logical chunk j → physical chunk j ^ (x & 7)."""
        for j in range(8):
            v = four_scalars[j >> 1]
            half = (v >> 128) if (j & 1) else (v & ((1 << 128) - 1))
            self.chunks[(x << 7) + ((j ^ (x & 7)) << 4)] = half

    def tma_store_row(self, x):
        out = []
        for cc in range(4):
            lo = self.chunks[(x << 7) + (((2 * cc) ^ (x & 7)) << 4)]
            hi = self.chunks[(x << 7) + (((2 * cc + 1) ^ (x & 7)) << 4)]
            out.append(lo | (hi << 128))
        return out

    def lds_fr(self, off):
        return self.chunks[off] | (self.chunks[off ^ 16] << 128)

    def sts_fr(self, off, v):
        self.chunks[off] = v & ((1 << 128) - 1)
        self.chunks[off ^ 16] = v >> 128


def tw_image(S, w_sub):
    """ntt_tw_image_kernel: per-round compact twiddles from tw[j] = ω_{2^S}^j."""
    tw = [pow(w_sub, j, R) for j in range(1 << (S - 1))]
    img = []
    b_top = S - 1
    while b_top >= 3:
        b_lo = b_top - 2
        for j in range(7 << b_lo):
            blk, v = j >> b_lo, j & ((1 << b_lo) - 1)
            if blk < 4:
                src = ((blk << b_lo) | v) << (S - 3 - b_lo)
            elif blk < 6:
                src = (((blk - 4) << b_lo) | v) << (S - 2 - b_lo)
            else:
                src = v << (S - 1 - b_lo)
            img.append(tw[src])
        b_top -= 3
    for i in range(3):
        img.append(tw[(i + 1) << (S - 3)])
    return img


def bfly(a, i, j, w):
    s, d = (a[i] + a[j]) % R, (a[i] - a[j]) % R
    a[i], a[j] = s, d * w % R


def run_pass(src, dst, L, p, img):
    """One launch of ntt_pass_tma_kernel over a single vector (batch index b = 0)."""
    S = p["S"]
    T, NTHR = 1 << (S + 2), 1 << (S - 1)
    blocks = 1 << (L - (S + 2))
    for bx in range(blocks):
        tile = Tile(S)
        col0 = row_base = rowrev0 = 0
        if p["type"] == 0:
            bpr_log = p["ncol_log"] - 2
            Rb = bx >> bpr_log
            col0 = (bx & ((1 << bpr_log) - 1)) << 2
            row_base = Rb << S
            for x in range(1 << S):                                  # tensor map: [rows][2^ncol_log] view, box = 4 scalars × 2^S rows
                g = ((row_base + x) << p["ncol_log"]) + col0
                tile.tma_load_row(x, src[g:g + 4])
        else:
            rowrev0 = bx << 2
            n2_log = p["nrows_log"] - p["n1_log"]
            for i in range(T):
                x, cc = i & ((1 << S) - 1), i >> S
                rr = rowrev0 + cc
                row = ((rr & ((1 << p["n1_log"]) - 1)) << n2_log) + (rr >> p["n1_log"])
                tile.sts_fr(tile_off(x, cc), src[(row << S) + x])
        if p["load_mode"] == 2:
            in_base = ((bx >> (p["ncol_log"] - 2)) << (S + p["ncol_log"])) + col0
            for i in range(T):
                x, cc = i >> 2, i & 3
                off = tile_off(x, cc)
                tile.sts_fr(off, tile.lds_fr(off) * p["l_full"][in_base + (x << p["ncol_log"]) + cc] % R)
        b_top, tw_off = S - 1, 0
        while b_top >= 3:
            b_lo = b_top - 2
            for tid in range(NTHR):
                c, tx = tid & 3, tid >> 2
                v, u = tx & ((1 << b_lo) - 1), tx >> b_lo
                xbase = (u << (b_lo + 3)) | v
                a = [tile.lds_fr(tile_off(xbase | (e << b_lo), c)) for e in range(8)]
                for e in range(4):
                    bfly(a, e, e + 4, img[tw_off + (e << b_lo) + v])
                for e in range(2):
                    w = img[tw_off + (4 << b_lo) + (e << b_lo) + v]
                    bfly(a, e, e + 2, w)
                    bfly(a, e + 4, e + 6, w)
                w3 = img[tw_off + (6 << b_lo) + v]
                for e in range(0, 8, 2):
                    bfly(a, e, e + 1, w3)
                for e in range(8):
                    tile.sts_fr(tile_off(xbase | (e << b_lo), c), a[e])
            tw_off += 7 << b_lo
            b_top -= 3
        regs = {}
        for tid in range(NTHR):                                       # every thread reads its rows, then the barrier
            c, tx = tid & 3, tid >> 2
            tx_rev = brev(tx, S - 3)                                   # bit-reversed row assignment of the last round
            regs[tid] = [tile.lds_fr(tile_off((tx_rev << 3) | e, c)) for e in range(8)]
        for tid in range(NTHR):
            c, tx = tid & 3, tid >> 2
            a = regs[tid]
            if b_top >= 2:
                bfly(a, 0, 4, 1)
                for e in range(1, 4):
                    bfly(a, e, e + 4, img[tw_off + e - 1])
            if b_top >= 1:
                w4 = img[tw_off + 1]
                bfly(a, 0, 2, 1)
                bfly(a, 4, 6, 1)
                bfly(a, 1, 3, w4)
                bfly(a, 5, 7, w4)
            for e in range(0, 8, 2):
                bfly(a, e, e + 1, 1)
            k_hi = tx
            for e in range(8):
                e_rev = ((e & 1) << 2) | (e & 2) | ((e >> 2) & 1)
                k = (e_rev << (S - 3)) | k_hi
                val = a[e]
                if p["store_mode"]:
                    idx = ((k << p["ncol_log"]) + col0 + c) if p["store_mode"] == 4 else ((rowrev0 + c) + (k << p["nrows_log"]))
                    val = val * p["s_full"][idx] % R
                tile.sts_fr(tile_off(k, c), val)
        for k in range(1 << S):                                       # tensor store
            row = tile.tma_store_row(k)
            if p["type"] == 0:
                g = ((row_base + k) << p["ncol_log"]) + col0
            else:
                g = (k << p["nrows_log"]) + rowrev0                   # [2^S rows (k)][2^nrows_log] view of the output vector
            dst[g:g + 4] = row


def plan(L, S_list, inverse, coset):
    """ntt_build_plan for the tma shape: passes, single-lookup tables (host ints instead of device tables)."""
    n = 1 << L
    dom = model.domain(n)
    w = dom["group_gen_inv"] if inverse else dom["group_gen"]
    g = pow(7, -1, R) if inverse else 7
    ninv = pow(n, -1, R)
    passes, done = [], 0
    P = len(S_list)
    for i, S in enumerate(S_list):
        p = {"S": S, "load_mode": 0, "store_mode": 0, "ncol_log": 0, "nrows_log": 0, "n1_log": 0, "l_full": None, "s_full": None}
        w_sub = pow(w, 1 << (L - S), R)
        if i < P - 1:
            p["type"] = 0
            p["ncol_log"] = L - done - S
            scale = ninv if (i == 0 and inverse and not coset) else 1      # hi_scaled of pass 0
            p["store_mode"] = 4
            p["s_full"] = [scale * pow(w, (j * k) << done, R) % R for k in range(1 << S) for j in range(1 << p["ncol_log"])]
        else:
            p["type"] = 1
            p["nrows_log"] = L - S
            p["n1_log"] = S_list[0] if P == 3 else p["nrows_log"]
            if inverse and coset:
                p["store_mode"] = 5
                p["s_full"] = [ninv * pow(g, e, R) % R for e in range(n)]
        if i == 0 and coset and not inverse:
            p["load_mode"] = 2
            p["l_full"] = [pow(g, e, R) for e in range(n)]
        passes.append((p, tw_image(S, w_sub)))
        done += S
    return passes


def fast_ntt(a, w):
    """Recursive radix-2 reference, natural order."""
    n = len(a)
    if n == 1:
        return a
    ev, od = fast_ntt(a[0::2], w * w % R), fast_ntt(a[1::2], w * w % R)
    out, t = [0] * n, 1
    for i in range(n // 2):
        x = od[i] * t % R
        out[i], out[i + n // 2] = (ev[i] + x) % R, (ev[i] - x) % R
        t = t * w % R
    return out


def reference(x, inverse, coset):
    n = len(x)
    dom = model.domain(n)
    a = list(x)
    if coset and not inverse:
        a = [v * pow(7, j, R) % R for j, v in enumerate(a)]
    a = fast_ntt(a, dom["group_gen_inv"] if inverse else dom["group_gen"])
    if inverse:
        ninv, gi = pow(n, -1, R), pow(7, -1, R)
        a = [v * ninv % R * (pow(gi, j, R) if coset else 1) % R for j, v in enumerate(a)]
    return a


@pytest.mark.parametrize("L,S_list", [(12, (6, 6)), (13, (7, 6)), (15, (8, 7)), (18, (6, 6, 6))])
def test_tma_pass_model_matches_definition(L, S_list):
    n = 1 << L
    x = model.random_fr(0x7A0 + L, n)
    variants = ((0, 0), (1, 0), (0, 1), (1, 1)) if L <= 13 else ((0, 0), (1, 1))
    for inverse, coset in variants:
        bufs = [list(x), [0] * n]
        passes = plan(L, S_list, inverse, coset)
        data, scratch = bufs
        for i, (p, img) in enumerate(passes):
            src = data if i == 0 else scratch
            dst = data if i == len(passes) - 1 else scratch
            if src is dst:                                                # the middle pass of a 3-pass plan runs in place
                tmp = list(src)
                run_pass(tmp, dst, L, p, img)
            else:
                run_pass(src, dst, L, p, img)
        assert data == reference(x, inverse, coset), (L, inverse, coset)


def test_swizzle_is_conflict_free_where_claimed():
    """Bank groups (16-byte chunks mod 8) of a quarter-warp's accesses: all distinct in the general rounds (b_lo ≥ 1), the
    type-1 staging loop and the coset sweep; 2-way in the last round (documented)."""
    for S in (6, 7, 8, 9):
        b_top = S - 1
        while b_top >= 3:
            b_lo = b_top - 2
            for q in range(0, 1 << (S - 1), 8):
                for e in range(8):
                    groups = set()
                    for tid in range(q, q + 8):
                        c, tx = tid & 3, tid >> 2
                        v, u = tx & ((1 << b_lo) - 1), tx >> b_lo
                        groups.add((tile_off(((u << (b_lo + 3)) | v) | (e << b_lo), c) >> 4) & 7)
                    assert len(groups) == 8, (S, b_lo)
            b_top -= 3
        for q in range(0, 64, 8):                                         # type-1 staging: lanes = consecutive x
            assert len({(tile_off(x, 1) >> 4) & 7 for x in range(q, q + 8)}) == 8
        for q in range(0, 64, 8):                                         # coset sweep / stores: lanes = consecutive (x, c)
            assert len({(tile_off(i >> 2, i & 3) >> 4) & 7 for i in range(q, q + 8)}) == 8
        groups = {(tile_off((brev(tx, S - 3) << 3) | 5, c) >> 4) & 7 for tx in (0, 1) for c in range(4)}
        assert len(groups) == 4                                            # last round's load: 2-way
        for q in range(0, 1 << (S - 1), 8):                                # … its write-back to the output rows: conflict free
            for e in range(8):
                e_rev = ((e & 1) << 2) | (e & 2) | ((e >> 2) & 1)
                assert len({(tile_off((e_rev << (S - 3)) | (tid >> 2), tid & 3) >> 4) & 7 for tid in range(q, q + 8)}) == 8
