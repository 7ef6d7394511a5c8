"""Host-side mirror of dusk-bls12_381 0.8 `multiscalar_mul::msm_variable_base` (pinned at
/root/reference/Cargo.toml:20) on top of the C ABI.

`msm_variable_base(points, scalars)`: points an (n, 12) uint64 array of packed affine Montgomery
x‖y, scalars an (n, 4) uint64 array of Montgomery limbs; returns the projective X‖Y‖Z (18 limbs,
not normalised — like upstream; Z = 0 for the identity).  Like upstream it is infallible for well-formed input and treats
an empty input as the identity.  `CommitKey` keeps the bases resident on the GPU the way dusk-plonk's
`CommitKey::powers_of_g` is held across commits.
"""
import numpy as np

from . import _native
from .domain import default_context


def msm_variable_base(points, scalars, ctx=None):
    ctx = ctx or default_context()
    points = np.ascontiguousarray(points, dtype=np.uint64).reshape(-1, 12)
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
    assert points.shape[0] == scalars.shape[0]
    if points.shape[0] == 0:
        out = np.zeros(18, np.uint64)
        return ctx.msm(None, scalars)  # the library returns the identity for n = 0
    srs = ctx.srs_upload(points)
    try:
        return ctx.msm(srs, scalars)
    finally:
        ctx.srs_free(srs)


def pippenger(points, scalars, ctx=None):
    """`multiscalar_mul::pippenger(points, scalars)` — the iterator form over `G1Projective` bases:
    points an (n, 18) uint64 array X‖Y‖Z (Z = 0: the identity), scalars (n, 4); returns the projective sum like
    `msm_variable_base`."""
    return (ctx or default_context()).pippenger(points, scalars)


_P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
_RP_INV = pow(1 << 384, -1, _P)


def g1_to_bytes(xyz):
    """`G1Affine::from(G1Projective).to_bytes()` of dusk-bls12_381 (zcash format) for the
    projective X‖Y‖Z the library returns: 48 bytes big-endian x; bit 7 = compressed, bit 6 = identity,
    bit 5 = y is the lexicographically larger root.  One point, host-side, cold."""
    v = [int(t) for t in np.asarray(xyz, dtype=np.uint64).reshape(18)]
    if not any(v[12:]):
        return bytes([0xC0]) + bytes(47)
    X, Y, Z = (sum(l << (64 * i) for i, l in enumerate(v[k:k + 6])) for k in (0, 6, 12))
    zi = pow(Z, -1, _P)  # `G1Affine::from(G1Projective)`: x = X/Z, y = Y/Z — the Montgomery factors cancel
    x, y = X * zi % _P, Y * zi % _P
    b = bytearray(x.to_bytes(48, "big"))
    b[0] |= 0x80 | (0x20 if y > (_P - 1) // 2 else 0)
    return bytes(b)


class CommitKey:
    """`CommitKey { powers_of_g }` with the powers resident in HBM; `commit` = one MSM over the prefix."""

    def __init__(self, powers_of_g, ctx=None, precompute=True):
        self.ctx = ctx or default_context()
        pts = np.ascontiguousarray(powers_of_g, dtype=np.uint64).reshape(-1, 12)
        self.n = pts.shape[0]
        self._srs = self.ctx.srs_upload(pts)
        if precompute and 1 <= self.n <= (1 << 22):
            self.ctx.srs_precompute(self._srs)  # pre-doubled window copies: ~2x faster prover-size commits

    def max_degree(self):
        return self.n - 1

    def commit(self, coeffs):
        """`CommitKey::commit(&Polynomial)`; also takes the bare coefficient array."""
        coeffs = np.ascontiguousarray(getattr(coeffs, "coeffs", coeffs), dtype=np.uint64).reshape(-1, 4)
        if coeffs.shape[0] > self.n:
            raise ValueError("PolynomialDegreeTooLarge")  # dusk-plonk `check_degree_is_within_bounds`
        return self.ctx.msm(self._srs, coeffs)

    def commit_bytes(self, coeffs):
        """`Commitment::to_bytes()`: the 48-byte compressed commitment — the first byte-comparable artefact."""
        return g1_to_bytes(self.commit(coeffs))

    def close(self):
        if self._srs is not None:
            self.ctx.srs_free(self._srs)
            self._srs = None
