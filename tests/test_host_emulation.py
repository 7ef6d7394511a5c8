"""CPU tests of the *device* arithmetic templates (csrc/field.cuh, csrc/g1.cuh): the same C++ the
kernels inline is compiled for the host with the PTX carry-chain primitives emulated
(PB200_HOST_EMU, csrc/carry.cuh) and compared limb-for-limb with the oracle.  This is how kernel
arithmetic is debugged on a box without a GPU; the emulation never ships."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import model

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host_emu", "emu_field.cpp")
OUT = os.path.join(ROOT, "tests", "host_emu", "emu_field.so")


@pytest.fixture(scope="module")
def emu():
    deps = [SRC] + [os.path.join(ROOT, "plonk-prototype_b200", "csrc", f) for f in ("carry.cuh", "field.cuh", "g1.cuh")]
    deps = [d for d in deps if os.path.exists(d)]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", SRC, "-o", OUT])
    return ctypes.CDLL(OUT)


def _run(lib, fn, op, a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    o = np.empty_like(a)
    getattr(lib, fn)(op, a.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p),
                     o.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(a.shape[0]))
    return o


def _rand_fp(oracle, seed, n):
    g = model.splitmix64_stream(seed)
    out = []
    while len(out) < n:
        v = 0
        for i in range(6):
            v |= next(g) << (64 * i)
        v &= (1 << 381) - 1
        if v < model.P:
            out.append(v)
    return oracle.ints_to_limbs(out, 6)


def test_fr_matches_oracle(emu, oracle):
    n = 5000
    a, b = oracle.random_fr(1, n), oracle.random_fr(2, n)
    edge = oracle.ints_to_limbs([0, 1, model.R - 1, model.R - 2, model.FR_MONT_R, (1 << 255) % model.R, 2**32 - 1,
                                 2**64 - 1], 4)
    a[:8], b[:8] = edge, edge[::-1]
    a[8:16], b[8:16] = edge, edge
    for op, f in ((0, oracle.fr_mul), (1, oracle.fr_add), (2, oracle.fr_sub)):
        assert (_run(emu, "emu_fr", op, a, b) == f(a, b)).all(), op
    assert (_run(emu, "emu_fr", 3, a, b) == oracle.fr_from_mont(a)).all()
    assert (_run(emu, "emu_fr", 4, a, b) == oracle.fr_to_mont(a)).all()
    assert (_run(emu, "emu_fr", 5, a[:64], b[:64]) == oracle.fr_inv(a[:64])).all()
    assert (_run(emu, "emu_fr", 7, a, b) == oracle.fr_mul(a, a)).all()   # Fr has 1 spare bit: sqr falls back to mul


def test_fp_matches_oracle(emu, oracle):
    n = 3000
    a, b = _rand_fp(oracle, 3, n), _rand_fp(oracle, 4, n)
    edge = oracle.ints_to_limbs([0, 1, model.P - 1, model.P - 2, model.FP_MONT_R, 1 << 380, 2**32 - 1, 2**64 - 1], 6)
    a[:8], b[:8] = edge, edge[::-1]
    a[8:16], b[8:16] = edge, edge
    for op, f in ((0, oracle.fp_mul), (1, oracle.fp_add), (2, oracle.fp_sub)):
        assert (_run(emu, "emu_fp", op, a, b) == f(a, b)).all(), op
    assert (_run(emu, "emu_fp", 3, a, b) == oracle.fp_from_mont(a)).all()
    assert (_run(emu, "emu_fp", 4, a, b) == oracle.fp_to_mont(a)).all()
    assert (_run(emu, "emu_fp", 5, a[:32], b[:32]) == oracle.fp_inv(a[:32])).all()
    assert (_run(emu, "emu_fp", 14, a, b) == oracle.fp_sub(np.zeros_like(a), a)).all()   # neg_nonzero = p − x (0 stays 0 in the harness)
    # dedicated squaring (triangular products on the pre-doubled operand) incl. the extreme values p−1, p−2, 2^380
    assert (_run(emu, "emu_fp", 7, a, b) == oracle.fp_mul(a, a)).all()
    top = oracle.ints_to_limbs([model.P - 1 - k for k in range(64)] + [(1 << 381) - 1 - 3 * k for k in range(64)
                                                                        if (1 << 381) - 1 - 3 * k < model.P], 6)
    assert (_run(emu, "emu_fp", 7, top, top) == oracle.fp_mul(top, top)).all()


def _g1(lib, op, P, Q, k=0):
    P32 = np.ascontiguousarray(P).view(np.uint32).reshape(-1, 24)
    Q32 = np.ascontiguousarray(Q).view(np.uint32).reshape(-1, 24)
    o = np.zeros((P32.shape[0], 25), np.uint32)
    lib.emu_g1(op, P32.ctypes.data_as(ctypes.c_void_p), Q32.ctypes.data_as(ctypes.c_void_p),
               o.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(P32.shape[0]), ctypes.c_uint64(k))
    return o


def _affine_ints(oracle, o):
    """emu output rows → list of canonical affine (x, y) or None."""
    res = []
    for row in o:
        if row[24]:
            res.append(None)
            continue
        xy = oracle.fp_from_mont(row[:24].copy().view(np.uint64).reshape(2, 6))
        res.append((oracle.limbs_to_int(xy[0]), oracle.limbs_to_int(xy[1])))
    return res


def test_g1_group_law_matches_model(emu, oracle):
    n = 24
    pts = oracle.synthetic_bases(2 * n)
    P, Q = pts[:n].copy(), pts[n:].copy()
    Q[0] = P[0]                                   # P + P  → doubling branch
    zero = np.zeros((1, 6), np.uint64)
    Q[1, :6] = P[1, :6]
    Q[1, 6:] = oracle.fp_sub(zero, P[1:2, 6:])    # P + (−P) → identity branch
    mp = model.synthetic_bases(2 * n)
    MP, MQ = mp[:n], mp[n:]
    MQ[0] = MP[0]
    MQ[1] = model.g1_neg(MP[1])
    want_add = [model.g1_add(a, b) for a, b in zip(MP, MQ)]
    assert _affine_ints(oracle, _g1(emu, 0, P, Q)) == want_add        # mixed add incl. exceptional cases
    assert _affine_ints(oracle, _g1(emu, 1, P, Q)) == want_add        # general add incl. exceptional cases
    want_dbl = [model.g1_add(a, a) for a in MP]
    assert _affine_ints(oracle, _g1(emu, 2, P, Q)) == want_dbl
    assert _affine_ints(oracle, _g1(emu, 3, P, Q)) == [model.g1_add(d, d) for d in want_dbl]
    for k in (0, 1, 2, 3, 0xB2, 2**21 - 1, 2**40 + 12345):
        assert _affine_ints(oracle, _g1(emu, 4, P[:6], Q[:6], k)) == [model.g1_mul(a, k) for a in MP[:6]]
    assert _affine_ints(oracle, _g1(emu, 5, P, Q)) == [model.g1_add(model.g1_add(model.g1_add(a, b), b), model.g1_neg(b))
                                                        for a, b in zip(MP, MQ)]
    assert _affine_ints(oracle, _g1(emu, 6, P, Q)) == MQ              # identity accumulator paths
    # the lazy mixed addition of the accumulation kernel (coordinates in [0, 2p), harness traps when one leaves the range):
    # P + 3Q − Q + P, incl. the pairs Q = P (tangent) and Q = −P (cancellation), and P + P − P − P + Q
    want7 = []
    for a, b in zip(MP, MQ):
        r = a
        for t in (b, b, b, model.g1_neg(b), a):
            r = model.g1_add(r, t)
        want7.append(r)
    assert _affine_ints(oracle, _g1(emu, 7, P, Q)) == want7
    assert _affine_ints(oracle, _g1(emu, 8, P, Q)) == MQ


def test_lazy_representation_matches_oracle(emu, oracle):
    """Field::{mul_lazy, add_lazy, sub_lazy} (values in [0, 2p), the arithmetic of the TMA NTT kernel between its passes):
    operands from both halves of the range — x and x + p — give the same canonical result as the oracle's modular operation,
    and every intermediate stays below 2p (the harness traps otherwise).  Fr has ONE spare bit, so x + y can exceed 2^256:
    the carry-out case is part of the sample (values next to r)."""
    n = 4000
    a, b = oracle.random_fr(21, n), oracle.random_fr(22, n)
    edge = oracle.ints_to_limbs([0, 1, model.R - 1, model.R - 2, model.FR_MONT_R, (1 << 255) % model.R, 2**32 - 1, 2**64 - 1], 4)
    k = len(edge)
    a[:k * k] = np.repeat(edge, k, axis=0)
    b[:k * k] = np.tile(edge, (k, 1))
    assert (_run(emu, "emu_fr", 9, a, b) == oracle.fr_mul(a, b)).all()
    assert (_run(emu, "emu_fr", 10, a, b) == oracle.fr_add(a, b)).all()
    assert (_run(emu, "emu_fr", 11, a, b) == oracle.fr_sub(a, b)).all()
    fa, fb = _rand_fp(oracle, 23, 1000), _rand_fp(oracle, 24, 1000)
    assert (_run(emu, "emu_fp", 9, fa, fb) == oracle.fp_mul(fa, fb)).all()
    assert (_run(emu, "emu_fp", 10, fa, fb) == oracle.fp_add(fa, fb)).all()
    assert (_run(emu, "emu_fp", 11, fa, fb) == oracle.fp_sub(fa, fb)).all()
    # Fp has three spare bits: both operands of a product may be lazy, and so may the operand of the dedicated squaring
    top = oracle.ints_to_limbs([model.P - 1 - j for j in range(64)] + [0, 1, 2], 6)
    fa[:67], fb[:67] = top, top[::-1].copy()
    assert (_run(emu, "emu_fp", 12, fa, fb) == oracle.fp_mul(fa, fb)).all()
    assert (_run(emu, "emu_fp", 13, fa, fb) == oracle.fp_mul(fa, fa)).all()
