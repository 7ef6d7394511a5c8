"""GPU parity tests for the MSM path, through the C ABI (pb200_msm_g1 / pb200_msm_g1_dev) and the
msm_variable_base / CommitKey mirrors, against (a) the big-int golden vectors, (b) the C oracle's
restatement of msm_variable_base on seeded inputs, (c) closed-form and linearity properties at sizes the
oracle cannot reach.  Results are compared as canonical affine points — bit-exact, no tolerance."""
import numpy as np
import pytest

import model
from helpers import closed_form_msm_scalar, load_golden
from test_oracle_cpu import expected_affine, msm_case_inputs

pytestmark = pytest.mark.gpu
A, D = 0xB2000001, 0x9E3779B1


def aff(oracle, xyz):
    return oracle.g1_proj_to_affine_canonical(xyz)


def test_golden_vectors(ctx, oracle):
    import plonk_prototype_b200 as pb
    for case in load_golden("msm_kat.json")["cases"]:
        pts, s = msm_case_inputs(oracle, case)
        out = pb.msm_variable_base(pts, s, ctx)
        got = aff(oracle, out)
        assert got == expected_affine(case), case["name"]
        assert model.g1_compress(got).hex() == case["compressed"]
        assert pb.g1_to_bytes(out).hex() == case["compressed"]            # Commitment::to_bytes, byte-identical
        # result convention: a valid (un-normalised) projective triple; (0, R, 0) for the identity
        r1 = oracle.fp_consts()["r1"]
        if got is None:
            assert not out[:6].any() and (out[6:12] == r1).all() and not out[12:].any()
        else:
            assert out[12:].any()


def test_empty_input_is_identity(ctx, oracle):
    import plonk_prototype_b200 as pb
    out = pb.msm_variable_base(np.zeros((0, 12), np.uint64), np.zeros((0, 4), np.uint64), ctx)
    assert aff(oracle, out) is None


@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 100, 257, 1000, 4096, 1 << 14, 1 << 16])
def test_random_matches_oracle(ctx, oracle, n):
    pts = oracle.synthetic_bases(n)
    s = oracle.fr_to_mont(oracle.random_fr(0xB2000000 + n, n))
    import plonk_prototype_b200 as pb
    got = aff(oracle, pb.msm_variable_base(pts, s, ctx))
    want = aff(oracle, oracle.msm_variable_base(pts, s, threads=16))
    assert got == want


def test_adversarial_scalar_sets(ctx, oracle):
    """SURVEY.md §8d adversarial sets: bucket skew and the exceptional cases of the group law."""
    n = 1 << 13
    pts = oracle.synthetic_bases(n)
    ck_ctx = ctx
    srs = ck_ctx.srs_upload(pts)
    rnd = model.random_fr(0xADD, n)
    sets = {
        "all_zero": [0] * n,
        "all_one": [1] * n,
        "all_r_minus_1": [model.R - 1] * n,
        "all_two": [2] * n,
        "eight_bit": [v & 0xFF for v in rnd],
        "half_zero": [0 if i & 1 else v for i, v in enumerate(rnd)],
        "one_heavy_bucket": [0x1234 if i % 3 else v for i, v in enumerate(rnd)],
        "top_window_only": [(v >> 240) << 240 for v in rnd],
    }
    try:
        for name, vals in sets.items():
            s = oracle.fr_to_mont(oracle.ints_to_limbs(vals, 4))
            got = aff(oracle, ck_ctx.msm(srs, s))
            want = aff(oracle, oracle.msm_variable_base(pts, s, threads=16))
            assert got == want, name
            k = closed_form_msm_scalar(s, A, D, model.R, model.FR_MONT_R)
            assert got == model.g1_mul(model.G1_GEN, k), name
    finally:
        ck_ctx.srs_free(srs)


def test_adversarial_base_sets(ctx, oracle):
    import plonk_prototype_b200 as pb
    n = 1 << 12
    base = oracle.synthetic_bases(n)
    s = oracle.fr_to_mont(oracle.random_fr(0xBA5E, n))
    same = np.repeat(base[:1], n, axis=0)                              # all-equal bases → P + P everywhere
    assert aff(oracle, pb.msm_variable_base(same, s, ctx)) == aff(oracle, oracle.msm_variable_base(same, s, threads=16))
    half = base[: n // 2]
    neg = half.copy()
    neg[:, 6:] = oracle.fp_sub(np.zeros((n // 2, 6), np.uint64), half[:, 6:])
    pm = np.ascontiguousarray(np.stack([half, neg], axis=1).reshape(n, 12))  # P, −P pairs
    assert aff(oracle, pb.msm_variable_base(pm, s, ctx)) == aff(oracle, oracle.msm_variable_base(pm, s, threads=16))
    s2 = np.repeat(s[: n // 2], 2, axis=0)                              # equal scalars on ±P → identity
    assert aff(oracle, pb.msm_variable_base(pm, np.ascontiguousarray(s2), ctx)) is None


def test_commit_key_prefix_and_offset(ctx, oracle):
    """CommitKey::commit uses powers_of_g[..len]; the ABI's (offset, n) addresses any sub-range."""
    import plonk_prototype_b200 as pb
    n = 3000
    pts = oracle.synthetic_bases(n)
    ck = pb.CommitKey(pts, ctx)
    s = oracle.fr_to_mont(oracle.random_fr(0xC0FFEE, 1234))
    assert aff(oracle, ck.commit(s)) == aff(oracle, oracle.msm_variable_base(pts[:1234], s, threads=8))
    got = aff(oracle, ctx.msm(ck._srs, s[:500], offset=2000))
    assert got == aff(oracle, oracle.msm_variable_base(pts[2000:2500], s[:500], threads=8))
    with pytest.raises(ValueError):
        ck.commit(np.zeros((n + 1, 4), np.uint64))                     # PolynomialDegreeTooLarge
    with pytest.raises(pb.Pb200Error):
        ctx.msm(ck._srs, s, offset=n - 10)                             # range outside the SRS
    ck.close()


def test_synthetic_bases_on_device_match_oracle(ctx, oracle):
    n = 5000
    d = ctx.malloc(n * 96)
    try:
        ctx.synthetic_bases_dev(d, n, A, D)
        got = np.empty((n, 12), np.uint64)
        ctx.d2h(got, d)
        assert (got == oracle.synthetic_bases(n, A, D)).all()
    finally:
        ctx.free(d)


@pytest.mark.parametrize("log_n", [18, 20, 22])
def test_closed_form_and_linearity_large(ctx, oracle, log_n):
    """Device-resident bases + scalars; exact check of the whole MSM against (Σ sᵢ(a+i·d))·G."""
    n = 1 << log_n
    bases = ctx.malloc(n * 96)
    sd = ctx.malloc(n * 32)
    try:
        ctx.synthetic_bases_dev(bases, n, A, D)
        sample = np.empty((64, 12), np.uint64)
        ctx.d2h(sample, bases + (n - 64) * 96)
        assert oracle.g1_on_curve(sample)
        srs = ctx.srs_wrap_dev(bases, n)
        s = oracle.random_fr(0xB2000000 + log_n, n)      # raw limbs < r: valid Montgomery representations
        t = oracle.random_fr(0xB2100000 + log_n, n)
        res = {}
        for name, v in (("s", s), ("t", t), ("s+t", oracle.fr_add(s, t))):
            ctx.h2d(sd, v)
            got = aff(oracle, ctx.msm_dev(srs, sd, n))
            assert got == model.g1_mul(model.G1_GEN, closed_form_msm_scalar(v, A, D, model.R, model.FR_MONT_R)), name
            res[name] = got
        assert model.g1_add(res["s"], res["t"]) == res["s+t"]
        ctx.srs_free(srs)
    finally:
        ctx.free(bases)
        ctx.free(sd)


def test_g1_sum_combines_partial_results(ctx, oracle):
    """Point-range sharding: Σ over shards of MSM(shard) == MSM(whole)."""
    n, parts = 4096, 4
    pts = oracle.synthetic_bases(n)
    s = oracle.fr_to_mont(oracle.random_fr(0x5A4D, n))
    srs = ctx.srs_upload(pts)
    try:
        step = n // parts
        partial = [ctx.msm(srs, s[i * step:(i + 1) * step], offset=i * step) for i in range(parts)]
        partial.append(ctx.msm(srs, np.zeros((8, 4), np.uint64)))        # an identity in the mix
        total = ctx.g1_sum(np.stack(partial))
        assert aff(oracle, total) == aff(oracle, ctx.msm(srs, s))
        assert aff(oracle, total) == aff(oracle, oracle.msm_variable_base(pts, s, threads=8))
        assert aff(oracle, ctx.g1_sum(np.zeros((0, 18), np.uint64))) is None
    finally:
        ctx.srs_free(srs)


@pytest.mark.parametrize("n", [1, 5, 300, 4096, 1 << 15])
def test_precomputed_srs_gives_identical_results(ctx, oracle, n):
    """pb200_srs_precompute (pre-doubled window copies, shared bucket set) must not change any result."""
    pts = oracle.synthetic_bases(n)
    srs = ctx.srs_upload(pts)
    try:
        ctx.srs_precompute(srs)
        rnd = model.random_fr(0x9E + n, n)
        sets = {"random": rnd, "all_one": [1] * n, "all_r_minus_1": [model.R - 1] * n, "zero": [0] * n,
                "eight_bit": [v & 0xFF for v in rnd], "top_window_only": [(v >> 240) << 240 for v in rnd]}
        for name, vals in sets.items():
            s = oracle.fr_to_mont(oracle.ints_to_limbs(vals, 4))
            got = aff(oracle, ctx.msm(srs, s))
            assert got == model.g1_mul(model.G1_GEN, closed_form_msm_scalar(s, A, D, model.R, model.FR_MONT_R)), name
            if n <= 4096:
                assert got == aff(oracle, oracle.msm_variable_base(pts, s, threads=8)), name
        if n >= 300:   # sub-ranges: a large one takes the pre-doubled path, a tiny one the plain path
            s = oracle.fr_to_mont(oracle.random_fr(0x77, n))
            half = n // 2
            assert aff(oracle, ctx.msm(srs, s[:half], offset=n - half)) == aff(
                oracle, oracle.msm_variable_base(pts[n - half:], s[:half], threads=8))
            assert aff(oracle, ctx.msm(srs, s[:7], offset=3)) == aff(oracle, oracle.msm_variable_base(pts[3:10], s[:7]))
    finally:
        ctx.srs_free(srs)


def test_precomputed_equal_bases(ctx, oracle):
    n = 2048
    same = np.repeat(oracle.synthetic_bases(1), n, axis=0)
    s = oracle.fr_to_mont(oracle.random_fr(0xE0, n))
    srs = ctx.srs_upload(same)
    try:
        ctx.srs_precompute(srs)
        assert aff(oracle, ctx.msm(srs, s)) == aff(oracle, oracle.msm_variable_base(same, s, threads=8))
    finally:
        ctx.srs_free(srs)


def test_many_sizes_closed_form(ctx, oracle):
    """Boundary sizes (segment / level / window edges) with and without pre-doubled copies, checked exactly."""
    nmax = 5000
    pts = oracle.synthetic_bases(nmax)
    srs_plain = ctx.srs_upload(pts)
    srs_pre = ctx.srs_upload(pts)
    ctx.srs_precompute(srs_pre)
    s_all = oracle.random_fr(0x51CE, nmax)
    try:
        sizes = list(range(1, 40)) + [63, 64, 65, 127, 128, 129, 255, 256, 257, 511, 513, 1000, 1023, 1025, 2047, 2049,
                                      3001, 4095, 4097, 4999, 5000]
        for n in sizes:
            want = model.g1_mul(model.G1_GEN, closed_form_msm_scalar(s_all[:n], A, D, model.R, model.FR_MONT_R))
            assert aff(oracle, ctx.msm(srs_plain, s_all[:n])) == want, ("plain", n)
            assert aff(oracle, ctx.msm(srs_pre, s_all[:n])) == want, ("pre", n)
    finally:
        ctx.srs_free(srs_plain)
        ctx.srs_free(srs_pre)


def test_two_level_scatter_path(ctx, oracle, monkeypatch):
    """The two-level scatter (shared-memory histograms + CTA-per-bin fine pass) normally only runs above 1 GB of
    entries; force it at small sizes, including skewed digit distributions and the pre-doubled SRS mode."""
    monkeypatch.setenv("PB200_MSM_TWO_LEVEL_MIN_BYTES", "0")
    n = 1 << 15
    pts = oracle.synthetic_bases(n)
    srs = ctx.srs_upload(pts)
    srs_pre = ctx.srs_upload(pts)
    ctx.srs_precompute(srs_pre)
    rnd = model.random_fr(0x2C, n)
    sets = {"random": rnd, "all_one": [1] * n, "eight_bit": [v & 0xFF for v in rnd], "all_r_minus_1": [model.R - 1] * n,
            "one_heavy_bucket": [0x1234 if i % 3 else v for i, v in enumerate(rnd)]}
    try:
        for name, vals in sets.items():
            s = oracle.fr_to_mont(oracle.ints_to_limbs(vals, 4))
            want = model.g1_mul(model.G1_GEN, closed_form_msm_scalar(s, A, D, model.R, model.FR_MONT_R))
            for m in (n, 40000 // 3, 1000):
                w = want if m == n else model.g1_mul(model.G1_GEN, closed_form_msm_scalar(s[:m], A, D, model.R, model.FR_MONT_R))
                assert aff(oracle, ctx.msm(srs, s[:m])) == w, (name, m, "plain")
                assert aff(oracle, ctx.msm(srs_pre, s[:m])) == w, (name, m, "pre")
    finally:
        ctx.srs_free(srs)
        ctx.srs_free(srs_pre)


@pytest.mark.parametrize("n,batch,pre", [(1 << 12, 4, True), (1000, 3, True), (1 << 14, 2, True), (1 << 12, 4, False), (77, 5, True)])
def test_batched_msm_matches_single_calls(ctx, oracle, n, batch, pre):
    """pb200_msm_g1_batch_dev: several scalar vectors over the same bases in one pass (own bucket set each) — equal,
    vector by vector, to separate calls and to the oracle; adversarial vectors (all zero, all one) ride along."""
    pts = oracle.synthetic_bases(n)
    srs = ctx.srs_upload(pts)
    if pre:
        ctx.srs_precompute(srs)
    stride = n + 5
    vecs = []
    for j in range(batch):
        if j == 1:
            vals = [0] * n
        elif j == 2:
            vals = [1] * n
        else:
            vals = model.random_fr(0xBA7C0 + j, n)
        vecs.append(oracle.fr_to_mont(oracle.ints_to_limbs(vals, 4)))
    buf = np.zeros((batch * stride, 4), np.uint64)
    for j, v in enumerate(vecs):
        buf[j * stride:j * stride + n] = v
    dev = ctx.malloc(buf.nbytes)
    ctx.h2d(dev, buf)
    try:
        got = ctx.msm_batch_dev(srs, dev, n, batch, stride)
        for j, v in enumerate(vecs):
            want = aff(oracle, oracle.msm_variable_base(pts, v, threads=8))
            assert aff(oracle, got[j]) == want, j
            assert aff(oracle, ctx.msm(srs, v)) == want, j
    finally:
        ctx.free(dev)
        ctx.srs_free(srs)


@pytest.mark.parametrize("k_log", [2, 3, 5, 8])
def test_bucket_reduction_chunk_sizes(ctx, oracle, monkeypatch, k_log):
    """The bucket reduction's chunk size is chosen by a wave / work cost model; every choice must give the same group
    element (plain and pre-doubled SRS, skewed digits included)."""
    monkeypatch.setenv("PB200_MSM_REDUCE_K_LOG", str(k_log))
    n = 1 << 14
    pts = oracle.synthetic_bases(n)
    srs = ctx.srs_upload(pts)
    srs_pre = ctx.srs_upload(pts)
    ctx.srs_precompute(srs_pre)
    rnd = model.random_fr(0x4B + k_log, n)
    try:
        for name, vals in (("random", rnd), ("eight_bit", [v & 0xFF for v in rnd]), ("all_r_minus_1", [model.R - 1] * n)):
            s = oracle.fr_to_mont(oracle.ints_to_limbs(vals, 4))
            want = model.g1_mul(model.G1_GEN, closed_form_msm_scalar(s, A, D, model.R, model.FR_MONT_R))
            assert aff(oracle, ctx.msm(srs, s)) == want, (name, "plain")
            assert aff(oracle, ctx.msm(srs_pre, s)) == want, (name, "pre")
    finally:
        ctx.srs_free(srs)
        ctx.srs_free(srs_pre)


@pytest.mark.parametrize("n", [1, 7, 8, 9, 100, 1000, 5000])
def test_pippenger_projective_bases(ctx, oracle, n):
    """`multiscalar_mul::pippenger` (iterator form over G1Projective) through pb200_pippenger_g1: bases
    (x·z, y·z, z) with random z per point — and the identity (Z = 0) at a few positions — give the group element the oracle's
    msm_variable_base gives on the affine points (identity terms dropped)."""
    import plonk_prototype_b200 as pb
    pts = oracle.synthetic_bases(n)                                    # (n, 12) affine, Montgomery
    s = oracle.fr_to_mont(oracle.random_fr(0xA13 + n, n))
    g = model.splitmix64_stream(0x2A13 + n)
    zs = []
    while len(zs) < n:
        v = 0
        for i in range(6):
            v |= next(g) << (64 * i)
        v &= (1 << 381) - 1
        if 0 < v < model.P:
            zs.append(v)
    z = oracle.ints_to_limbs(zs, 6)
    xyz = np.concatenate([oracle.fp_mul(np.ascontiguousarray(pts[:, :6]), z), oracle.fp_mul(np.ascontiguousarray(pts[:, 6:]), z), z], axis=1)
    ident = [i for i in (0, 3, n - 1) if i < n and n > 1][: max(0, min(3, n - 1))]
    for i in ident:
        xyz[i, 12:] = 0                                                 # Z = 0: the identity, whatever X and Y hold
    got = aff(oracle, pb.pippenger(xyz, s, ctx))
    s_ref = s.copy()
    s_ref[ident] = 0
    want = aff(oracle, oracle.msm_variable_base(pts, s_ref, threads=8))
    assert got == want


def test_pippenger_empty_and_all_identity(ctx, oracle):
    import plonk_prototype_b200 as pb
    assert aff(oracle, pb.pippenger(np.zeros((0, 18), np.uint64), np.zeros((0, 4), np.uint64), ctx)) is None
    xyz = np.zeros((5, 18), np.uint64)
    xyz[:, 6] = 1
    assert aff(oracle, pb.pippenger(xyz, oracle.fr_to_mont(oracle.random_fr(5, 5)), ctx)) is None
