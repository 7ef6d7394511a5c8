"""GPU parity tests for the NTT path, through the C ABI (pb200_ntt / pb200_ntt_dev) and the
EvaluationDomain mirror, against (a) the big-int golden vectors, (b) the C oracle on seeded inputs,
(c) size-independent properties at BASELINE sizes.  Bit-exact: integer work, no tolerance."""
import numpy as np
import pytest

import model
from helpers import limbs_to_hex, load_golden, sha_scalars_canonical

pytestmark = pytest.mark.gpu

VARIANTS = (("fft", 0, 0), ("ifft", 1, 0), ("coset_fft", 0, 1), ("coset_ifft", 1, 1))


def test_golden_vectors(ctx, oracle):
    import plonk_prototype_b200 as pb
    for case in load_golden("ntt_kat.json")["cases"]:
        n = 1 << case["log_n"]
        st = case.get("structured")
        if st == "short37_zero_padded":
            x = model.random_fr(case["seed"], 37)          # the mirror zero-pads like upstream
        elif st == "delta0":
            x = [1] + [0] * (n - 1)
        elif st == "all_r_minus_1":
            x = [model.R - 1] * n
        else:
            x = model.random_fr(case["seed"], n)
        xm = oracle.fr_to_mont(oracle.ints_to_limbs(x, 4))
        dom = pb.EvaluationDomain(n, ctx)
        for name, inv, cos in VARIANTS:
            if name + "_sha256" not in case:
                continue
            y = getattr(dom, name)(xm)
            yc = oracle.fr_from_mont(y)
            assert sha_scalars_canonical(yc) == case[name + "_sha256"], (case["log_n"], st, name)
            if name + "_head" in case:
                assert [limbs_to_hex(v, 32) for v in yc[:4]] == case[name + "_head"]


@pytest.mark.parametrize("log_n", list(range(0, 19)))
def test_matches_oracle_every_size(ctx, oracle, log_n):
    n = 1 << log_n
    x = oracle.fr_to_mont(oracle.random_fr(0xF1F00000 + log_n, n))
    for name, inv, cos in VARIANTS:
        got = x.copy()
        ctx.ntt(got, log_n, inv, cos)
        want = oracle.ntt(x, inv, cos, threads=8)
        assert (got == want).all(), (log_n, name, int((got != want).any(axis=1).sum()))


@pytest.mark.parametrize("log_n", [20, 22, 23])
def test_matches_oracle_large(ctx, oracle, log_n):
    n = 1 << log_n
    x = oracle.fr_to_mont(oracle.random_fr(0xF1F00000 + log_n, n))
    for name, inv, cos in (VARIANTS if log_n == 20 else VARIANTS[:1] + VARIANTS[3:]):
        got = x.copy()
        ctx.ntt(got, log_n, inv, cos)
        want = oracle.ntt(x, inv, cos, threads=16)
        assert (got == want).all(), (log_n, name)


@pytest.mark.parametrize("log_n,variants", [(24, (0, 1, 2, 3)), (26, (0,))])
def test_full_vector_matches_oracle_at_baseline_sizes(ctx, oracle, log_n, variants):
    """BASELINE.json configs[2] at its largest sizes, every output element against the CPU restatement (all host threads):
    2^24 in all four variants, 2^26 forward."""
    import os
    n = 1 << log_n
    x = oracle.random_fr(0xF1F00000 + log_n, n)           # raw limbs < r are valid Montgomery representations
    threads = os.cpu_count() or 8
    d = ctx.malloc(x.nbytes)
    try:
        got = np.empty_like(x)
        for v in variants:
            name, inv, cos = VARIANTS[v]
            ctx.h2d(d, x)
            ctx.ntt_dev(d, log_n, inv, cos)
            ctx.d2h(got, d)
            want = oracle.ntt(x, inv, cos, threads=threads)
            assert (got == want).all(), (log_n, name, int((got != want).any(axis=1).sum()))
            del want
    finally:
        ctx.free(d)


def test_edge_vectors(ctx, oracle):
    for log_n in (5, 12, 14):
        n = 1 << log_n
        zero = np.zeros((n, 4), np.uint64)
        for name, inv, cos in VARIANTS:
            got = zero.copy()
            ctx.ntt(got, log_n, inv, cos)
            assert not got.any()
        one = oracle.fr_to_mont(oracle.ints_to_limbs([1], 4))[0]
        delta = zero.copy()
        delta[0] = one
        got = delta.copy()
        ctx.ntt(got, log_n, 0, 0)
        assert (got == one).all()                       # NTT(δ0) = all ones
        ones = np.repeat(one.reshape(1, 4), n, axis=0)
        got = ones.copy()
        ctx.ntt(got, log_n, 1, 0)
        assert (got == delta).all()                     # iNTT(all ones) = δ0


@pytest.mark.parametrize("log_n", [24, 26])
def test_properties_at_baseline_sizes(ctx, oracle, log_n):
    """Device-resident round trips + linearity spot-check at sizes the CPU oracle cannot finish quickly."""
    n = 1 << log_n
    x = oracle.random_fr(0xF1F00000 + log_n, n)  # raw limbs < r are valid Montgomery representations
    nbytes = x.nbytes
    d = ctx.malloc(nbytes)
    try:
        ctx.h2d(d, x)
        back = np.empty_like(x)
        ctx.ntt_dev(d, log_n, 0, 0)
        ctx.ntt_dev(d, log_n, 1, 0)
        ctx.d2h(back, d)
        assert (back == x).all()
        ctx.ntt_dev(d, log_n, 0, 1)
        ctx.ntt_dev(d, log_n, 1, 1)
        ctx.d2h(back, d)
        assert (back == x).all()
        # forward transform: compare a few output points with the definition X_i = Σ_j x_j ω^{ij},
        # evaluated via Horner by the oracle-independent big-int model on a sparse input.
        sparse = np.zeros_like(x)
        idx = [0, 1, 5, n // 2 + 3, n - 1]
        vals = model.random_fr(77, len(idx))
        sm = oracle.fr_to_mont(oracle.ints_to_limbs(vals, 4))
        for k, j in enumerate(idx):
            sparse[j] = sm[k]
        ctx.h2d(d, sparse)
        ctx.ntt_dev(d, log_n, 0, 0)
        ctx.d2h(back, d)
        w = model.domain(n)["group_gen"]
        for i in (0, 1, 2, n // 3, n - 1):
            want = sum(v * pow(w, (i * j) % n, model.R) for v, j in zip(vals, idx)) % model.R
            got = oracle.limbs_to_int(oracle.fr_from_mont(back[i:i + 1])[0])
            assert got == want, (log_n, i)
    finally:
        ctx.free(d)


def test_invalid_domain_is_an_error(ctx):
    import plonk_prototype_b200 as pb
    with pytest.raises(pb.Pb200Error):
        ctx.ntt_dev(ctx.malloc(32), 32, 0, 0)


@pytest.mark.parametrize("log_n,batch", [(1, 3), (5, 7), (11, 4), (13, 5), (16, 3)])
def test_batched_transforms_match_single(ctx, oracle, log_n, batch):
    n = 1 << log_n
    x = oracle.fr_to_mont(oracle.random_fr(0xBA7C + log_n, n * batch))
    d = ctx.malloc(x.nbytes)
    try:
        for name, inv, cos in VARIANTS:
            ctx.h2d(d, x)
            ctx.ntt_batch_dev(d, log_n, batch, inv, cos)
            got = np.empty_like(x)
            ctx.d2h(got, d)
            for b in range(batch):
                assert (got[b * n:(b + 1) * n] == oracle.ntt(x[b * n:(b + 1) * n], inv, cos, threads=4)).all(), (name, b)
    finally:
        ctx.free(d)
