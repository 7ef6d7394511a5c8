"""Byte (de)serialisation of the KZG10 parameters, host side: dusk-plonk 0.8.2 `commitment_scheme::kzg10::{CommitKey,
OpeningKey, PublicParameters}::{to_raw_bytes, from_slice_unchecked, to_var_bytes}` (crate pinned at
/root/reference/Cargo.toml:19; SURVEY.md §2.2 D5, §5 "checkpoint / resume", §8f-1) — the reference never calls them, a
host that caches a generated SRS on disk does.

**Unpinned** like the rest of the upstream surface (UPSTREAM_ASSUMPTIONS.md §Serialisation): the layouts below are restated
from memory of the crates —
  * `G1Affine::to_raw_bytes` (dusk-bls12_381): 97 bytes = x ‖ y as 2 × 6 little-endian u64 limbs in MONTGOMERY form (the
    memory image) ‖ one infinity byte;  `G1Affine::to_bytes`: the 48-byte compressed encoding (`msm.g1_to_bytes`);
  * `CommitKey::to_raw_bytes`: u64 LE point count ‖ the points' raw bytes;  `to_var_bytes`: the points compressed, no prefix;
  * `OpeningKey::to_bytes`: g (48, compressed G1) ‖ h (96, compressed G2) ‖ β·h (96) = 240 bytes;
  * `PublicParameters::to_raw_bytes` / `to_var_bytes`: the opening key, then the commit key.
The point data never passes through Python integers in the raw form (numpy reshapes of the device image); the opening key's
two G2 points do (compression needs canonical coordinates and a square root on the way back) — cold, three points per SRS."""
import numpy as np

from . import _native

P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
_R_INV = pow(1 << 384, -1, P)
G1_RAW_SIZE, G1_SIZE, G2_SIZE, OPENING_KEY_SIZE = 97, 48, 96, 240
# the generators' compressed encodings (zcash format)
G1_GENERATOR_BYTES = bytes.fromhex("97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb")
G2_GENERATOR_BYTES = bytes.fromhex(
    "93e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e"
    "024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8")


# ---- Fp2 = Fp[u]/(u² + 1), pairs (c0, c1) of canonical integers ------------------------------------------------------
def _fp2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def _fp2_pow(a, e):
    r = (1, 0)
    while e:
        if e & 1:
            r = _fp2_mul(r, a)
        a = _fp2_mul(a, a)
        e >>= 1
    return r


def _fp2_sqrt(a):
    """A square root in Fp2 (p ≡ 3 mod 4), or None."""
    if a == (0, 0):
        return (0, 0)
    a1 = _fp2_pow(a, (P - 3) // 4)
    alpha = _fp2_mul(_fp2_mul(a1, a1), a)
    x0 = _fp2_mul(a1, a)
    if alpha == (P - 1, 0):
        x = ((-x0[1]) % P, x0[0])                       # u·x0
    else:
        b = _fp2_pow(((1 + alpha[0]) % P, alpha[1]), (P - 1) // 2)
        x = _fp2_mul(b, x0)
    return x if _fp2_mul(x, x) == a else None


def _limbs_to_int_mont(l6):
    return sum(int(v) << (64 * i) for i, v in enumerate(l6)) * _R_INV % P


def _int_to_limbs_mont(v):
    v = v * (1 << 384) % P
    return [(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(6)]


def g2_to_bytes(xy24):
    """`G2Affine::to_bytes`: x.c1 ‖ x.c0 big-endian; bit 7 compressed, bit 5 = y lexicographically larger (c1 first)."""
    v = np.asarray(xy24, dtype=np.uint64).reshape(4, 6)
    x0, x1, y0, y1 = (_limbs_to_int_mont(v[k]) for k in range(4))
    b = bytearray(x1.to_bytes(48, "big") + x0.to_bytes(48, "big"))
    larger = y1 > (P - 1) // 2 if y1 else y0 > (P - 1) // 2
    b[0] |= 0x80 | (0x20 if larger else 0)
    return bytes(b)


def g2_from_bytes(b):
    """Inverse of g2_to_bytes for a finite compressed point → 24 u64 (x.c0 ‖ x.c1 ‖ y.c0 ‖ y.c1, Montgomery)."""
    assert len(b) == 96 and b[0] & 0x80 and not b[0] & 0x40, "a finite compressed G2 point"
    sign = bool(b[0] & 0x20)
    x1 = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:48], "big")
    x0 = int.from_bytes(b[48:], "big")
    assert x0 < P and x1 < P
    x = (x0, x1)
    rhs = _fp2_mul(_fp2_mul(x, x), x)
    rhs = ((rhs[0] + 4) % P, (rhs[1] + 4) % P)          # y² = x³ + 4(u + 1)
    y = _fp2_sqrt(rhs)
    assert y is not None, "not on the curve"
    larger = y[1] > (P - 1) // 2 if y[1] else y[0] > (P - 1) // 2
    if larger != sign:
        y = ((-y[0]) % P, (-y[1]) % P)
    return np.array(_int_to_limbs_mont(x[0]) + _int_to_limbs_mont(x[1]) + _int_to_limbs_mont(y[0]) + _int_to_limbs_mont(y[1]), dtype=np.uint64)


def opening_key_to_bytes(beta_h):
    """`OpeningKey::to_bytes` for the opening key (G, H, β·H) of a trapdoor-generated SRS (`_native.opening_key_from_tau`)."""
    return G1_GENERATOR_BYTES + G2_GENERATOR_BYTES + g2_to_bytes(beta_h)


def opening_key_from_bytes(b):
    """→ β·H as the 24 u64 `verify` takes.  g and h must be the generators (dusk's `setup` never produces anything else)."""
    assert len(b) == OPENING_KEY_SIZE
    if b[:48] != G1_GENERATOR_BYTES or b[48:144] != G2_GENERATOR_BYTES:
        raise ValueError("opening key over other generators than the BLS12-381 ones")
    return g2_from_bytes(b[144:])


# ---- commit key ---------------------------------------------------------------------------------------------------------
def commit_key_to_raw_bytes(ctx, srs):
    """`CommitKey::to_raw_bytes` of a device-resident SRS: one D2H copy and a numpy reshape."""
    n = _native.lib().pb200_srs_len(srs)
    pts = np.empty((n, 96), np.uint8)
    if n:
        ctx.d2h(pts, ctx.srs_dev_ptr(srs))
    raw = np.zeros((n, G1_RAW_SIZE), np.uint8)           # the infinity byte stays 0: an SRS never holds the identity
    raw[:, :96] = pts
    return int(n).to_bytes(8, "little") + raw.tobytes()


def commit_key_from_slice_unchecked(ctx, b, precompute=True):
    """`CommitKey::from_slice_unchecked` → an SRS handle resident on the GPU (no curve / subgroup checks, like upstream)."""
    n = int.from_bytes(b[:8], "little")
    if len(b) != 8 + n * G1_RAW_SIZE:
        raise ValueError("CommitKey: %d bytes for %d points" % (len(b), n))
    raw = np.frombuffer(b, dtype=np.uint8, offset=8).reshape(n, G1_RAW_SIZE)
    if raw[:, 96].any():
        raise ValueError("CommitKey holds the point at infinity")
    pts = np.ascontiguousarray(raw[:, :96]).view(np.uint64).reshape(n, 12)
    srs = ctx.srs_upload(pts)
    if precompute and 1 <= n <= (1 << 22):
        ctx.srs_precompute(srs)
    return srs


def commit_key_to_var_bytes(ctx, srs):
    """`CommitKey::to_var_bytes`: compressed points (canonical coordinates need one host pass per point — meant for small keys)."""
    from .msm import g1_to_bytes
    n = _native.lib().pb200_srs_len(srs)
    pts = np.empty((n, 12), np.uint64)
    if n:
        ctx.d2h(pts, ctx.srs_dev_ptr(srs))
    one = np.array([0x760900000002FFFD, 0xEBF4000BC40C0002, 0x5F48985753C758BA, 0x77CE585370525745, 0x5C071A97A256EC6D, 0x15F65EC3FA80E493], np.uint64)
    return b"".join(g1_to_bytes(np.concatenate([p, one])) for p in pts)   # (x, y, 1) in Montgomery form
