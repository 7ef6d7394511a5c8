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
