"""Shared helpers for the parity tests (canonical encodings, hashing, golden loading)."""
import hashlib
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def sha_scalars_canonical(canon_limbs):
    """sha256 over 32-byte little-endian canonical scalars ((n,4) uint64 LE limbs are exactly that)."""
    a = np.ascontiguousarray(canon_limbs, dtype="<u8")
    return hashlib.sha256(a.tobytes()).hexdigest()


def hex_to_limbs(hx, nlimbs):
    v = int(hx, 16)
    return [(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(nlimbs)]


def limbs_to_hex(l, nbytes):
    v = 0
    for i, x in enumerate(np.asarray(l).reshape(-1)):
        v |= int(x) << (64 * i)
    return format(v, "0%dx" % (2 * nbytes))


def closed_form_msm_scalar(scalars_mont, a, d, r, mont_r):
    """For synthetic bases P_i = (a + i·d)·G:  Σ s_i·P_i = k·G with k = Σ s_i·(a + i·d) mod r.
    `scalars_mont` holds Montgomery limbs (value·2^256 mod r), so k = R⁻¹·Σ limbs_i·(a + i·d).
    Exact integer arithmetic with numpy: 16-bit × 16-bit partial dot products (each < 2^64 for n ≤ 2^30)."""
    s = np.ascontiguousarray(scalars_mont, dtype=np.uint64).reshape(-1, 4)
    n = s.shape[0]
    k = np.uint64(a) + np.arange(n, dtype=np.uint64) * np.uint64(d)           # < 2^64 for the sizes used
    assert a + (n - 1) * d < 2**64
    s16 = s.view(np.uint16).reshape(n, 16).astype(np.uint64)                   # little-endian 16-bit pieces
    k16 = k.view(np.uint16).reshape(n, 4).astype(np.uint64)
    total = 0
    for i in range(16):
        col = np.ascontiguousarray(s16[:, i])
        for j in range(4):
            total += int(np.dot(col, k16[:, j])) << (16 * (i + j))
    return total * pow(mont_r, -1, r) % r
