"""ctypes loader for oracle/liboracle.so (TEST INFRASTRUCTURE — the checker, never the product).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this.  Arrays are numpy uint64, little-endian limbs: Fr = 4 limbs, Fp = 6, packed affine G1 =
12 (x‖y), projective = 18 (X‖Y‖Z); all Montgomery unless a name says "canonical".
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_U64P = ctypes.POINTER(ctypes.c_uint64)


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = [os.path.join(_HERE, f) for f in ("oracle.c", "plonk_oracle.inc", "mont.h", "mont_asm.h", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        _LIB = ctypes.CDLL(so)
        _LIB.orc_init()
    return _LIB


def _p(a):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_U64P)


def _binop(name, w):
    def f(a, b):
        a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, w)
        b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, w)
        o = np.empty_like(a)
        getattr(lib(), name)(_p(a), _p(b), _p(o), ctypes.c_size_t(a.shape[0]))
        return o
    return f


def _unop(name, w):
    def f(a):
        a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, w)
        o = np.empty_like(a)
        getattr(lib(), name)(_p(a), _p(o), ctypes.c_size_t(a.shape[0]))
        return o
    return f


fr_mul, fr_add, fr_sub = _binop("orc_fr_mul", 4), _binop("orc_fr_add", 4), _binop("orc_fr_sub", 4)
fp_mul, fp_add, fp_sub = _binop("orc_fp_mul", 6), _binop("orc_fp_add", 6), _binop("orc_fp_sub", 6)
fr_to_mont, fr_from_mont, fr_inv = _unop("orc_fr_to_mont", 4), _unop("orc_fr_from_mont", 4), _unop("orc_fr_inv", 4)
fp_to_mont, fp_from_mont, fp_inv = _unop("orc_fp_to_mont", 6), _unop("orc_fp_from_mont", 6), _unop("orc_fp_inv", 6)


def fr_consts():
    m, r1, r2, root, gen = (np.zeros(4, np.uint64) for _ in range(5))
    inv = ctypes.c_uint64()
    lib().orc_fr_consts(_p(m), ctypes.byref(inv), _p(r1), _p(r2), _p(root), _p(gen))
    return {"modulus": m, "inv": inv.value, "r1": r1, "r2": r2, "root_of_unity": root, "generator": gen}


def fp_consts():
    m, r1, r2 = (np.zeros(6, np.uint64) for _ in range(3))
    inv = ctypes.c_uint64()
    lib().orc_fp_consts(_p(m), ctypes.byref(inv), _p(r1), _p(r2))
    return {"modulus": m, "inv": inv.value, "r1": r1, "r2": r2}


def random_fr(seed, n):
    """n uniform values in [0, r) as raw limbs (n, 4) — same stream as model.random_fr."""
    out = np.empty((n, 4), np.uint64)
    lib().orc_random_fr(ctypes.c_uint64(seed), ctypes.c_size_t(n), _p(out))
    return out


def g1_generator():
    out = np.empty(12, np.uint64)
    lib().orc_g1_generator(_p(out))
    return out


def g1_on_curve(xy):
    xy = np.ascontiguousarray(xy, dtype=np.uint64).reshape(-1, 12)
    return bool(lib().orc_g1_on_curve(_p(xy), ctypes.c_size_t(xy.shape[0])))


def g1_proj_to_affine_canonical(xyz):
    """(is_identity, x_int, y_int) from a projective Montgomery X‖Y‖Z."""
    xyz = np.ascontiguousarray(xyz, dtype=np.uint64).reshape(18)
    out = np.empty(12, np.uint64)
    inf = lib().orc_g1_proj_to_affine_canonical(_p(xyz), _p(out))
    if inf:
        return None
    return (limbs_to_int(out[:6]), limbs_to_int(out[6:]))


def synthetic_bases(n, a=0xB2000001, d=0x9E3779B1):
    out = np.empty((n, 12), np.uint64)
    lib().orc_synthetic_bases(ctypes.c_size_t(n), ctypes.c_uint64(a), ctypes.c_uint64(d), _p(out))
    return out


def msm_variable_base(points, scalars_mont, threads=1):
    points = np.ascontiguousarray(points, dtype=np.uint64).reshape(-1, 12)
    scalars_mont = np.ascontiguousarray(scalars_mont, dtype=np.uint64).reshape(-1, 4)
    assert points.shape[0] == scalars_mont.shape[0]
    out = np.empty(18, np.uint64)
    lib().orc_msm_variable_base(_p(points), _p(scalars_mont), ctypes.c_size_t(points.shape[0]), _p(out),
                                ctypes.c_int(threads))
    return out


def msm_naive(points, scalars_mont):
    points = np.ascontiguousarray(points, dtype=np.uint64).reshape(-1, 12)
    scalars_mont = np.ascontiguousarray(scalars_mont, dtype=np.uint64).reshape(-1, 4)
    out = np.empty(18, np.uint64)
    lib().orc_msm_naive(_p(points), _p(scalars_mont), ctypes.c_size_t(points.shape[0]), _p(out))
    return out


def ntt(data, inverse=False, coset=False, threads=1):
    """EvaluationDomain fft / ifft / coset_fft / coset_ifft on a power-of-two (n, 4) Montgomery array."""
    a = np.array(data, dtype=np.uint64).reshape(-1, 4)
    n = a.shape[0]
    log_n = n.bit_length() - 1
    assert 1 << log_n == n
    rc = lib().orc_ntt(_p(a), ctypes.c_uint32(log_n), int(inverse), int(coset), int(threads))
    if rc != 0:
        raise ValueError("InvalidEvalDomainSize")
    return a


def domain(log_n):
    g, gi, si, geni = (np.zeros(4, np.uint64) for _ in range(4))
    rc = lib().orc_domain(ctypes.c_uint32(log_n), _p(g), _p(gi), _p(si), _p(geni))
    if rc != 0:
        raise ValueError("InvalidEvalDomainSize")
    return {"group_gen": g, "group_gen_inv": gi, "size_inv": si, "generator_inv": geni}


def limbs_to_int(l):
    v = 0
    for i, x in enumerate(np.asarray(l).reshape(-1)):
        v |= int(x) << (64 * i)
    return v


def int_to_limbs(v, n):
    return np.array([(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(n)], dtype=np.uint64)


def ints_to_limbs(vals, n):
    return np.array([[(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(n)] for v in vals], dtype=np.uint64).reshape(-1, n)


# ----------------------------------------------------------------------------- PLONK prover restatement
def merlin_selftest(label, msg_label, msg, ch_label, n):
    out = ctypes.create_string_buffer(n)
    lib().orc_merlin_selftest(label, msg_label, msg, ctypes.c_size_t(len(msg)), ch_label, out, ctypes.c_size_t(n))
    return out.raw


def srs_setup(tau_mont, n):
    """powers_of_g[i] = τ^i·G as packed affine Montgomery points (n, 12) — small n (one scalar mul each)."""
    tau = np.ascontiguousarray(tau_mont, dtype=np.uint64).reshape(4)
    out = np.empty((n, 12), np.uint64)
    lib().orc_srs_setup(_p(tau), ctypes.c_size_t(n), _p(out))
    return out


def plonk_prove(selectors, wires, values_mont, pi_pos, pi_mont, srs, label, threads=1, reps=None):
    """Preprocess + prove on the CPU.  selectors: 11 (n_gates, 4) uint64 arrays or None; wires: 4 uint32 arrays;
    srs: (≥ n_pad, 12) packed affine points.  Returns (proof bytes, vk bytes, preprocess seconds, prove seconds);
    with reps = k the prove runs k times on the one preprocessed key and the last element is the list of k timings."""
    n_gates = len(wires[0])
    keep = []
    sel_arr = (ctypes.c_void_p * 11)()
    for k in range(11):
        if selectors[k] is None:
            sel_arr[k] = None
        else:
            a = np.ascontiguousarray(selectors[k], dtype=np.uint64).reshape(-1, 4)
            assert a.shape[0] == n_gates
            keep.append(a)
            sel_arr[k] = a.ctypes.data
    w_arr = (ctypes.c_void_p * 4)()
    for k in range(4):
        a = np.ascontiguousarray(wires[k], dtype=np.uint32)
        keep.append(a)
        w_arr[k] = a.ctypes.data
    vals = np.ascontiguousarray(values_mont, dtype=np.uint64).reshape(-1, 4)
    pos = np.ascontiguousarray(pi_pos, dtype=np.uint32)
    piv = np.ascontiguousarray(pi_mont, dtype=np.uint64).reshape(-1, 4)
    srs = np.ascontiguousarray(srs, dtype=np.uint64).reshape(-1, 12)
    n_pad = 1
    while n_pad < n_gates:
        n_pad *= 2
    assert srs.shape[0] >= n_pad
    vk = ctypes.create_string_buffer(15 * 48)
    proof = ctypes.create_string_buffer(1040)
    k = 1 if reps is None else int(reps)
    tm = (ctypes.c_double * (1 + k))()
    f = lib().orc_plonk_prove_reps
    f.restype = ctypes.c_int
    rc = f(ctypes.c_size_t(n_gates), ctypes.c_size_t(vals.shape[0]), sel_arr, w_arr, ctypes.c_void_p(vals.ctypes.data),
           ctypes.c_void_p(pos.ctypes.data), ctypes.c_void_p(piv.ctypes.data), ctypes.c_size_t(pos.shape[0]),
           ctypes.c_void_p(srs.ctypes.data), bytes(label), ctypes.c_size_t(len(label)), ctypes.c_int(threads), vk, proof,
           ctypes.c_int(k), tm)
    if rc != 0:
        raise RuntimeError("orc_plonk_prove failed: %d" % rc)
    return proof.raw, vk.raw, tm[0], (tm[1] if reps is None else [tm[1 + i] for i in range(k)])


def synthetic_circuit_columns(n_gates, seed=0x5EED, n_pub=2):
    """The synthetic arithmetic circuit of bench.py assembled by the checker itself (orc_synthetic_circuit),
    so the reference arm of bench.py never touches the product library.  Same return shape as
    plonk_prototype_b200.synth.synthetic_circuit_columns: (selectors[11], wires[4], values_mont, pi_pos, pi_vals_mont)."""
    assert n_gates >= 8 + n_pub and n_pub >= 1
    steps = (n_gates - n_pub - 3) // 2
    n_vars = 6 + 2 * steps + n_pub - 1
    sel7 = [np.empty((n_gates, 4), np.uint64) for _ in range(7)]
    wires = [np.empty(n_gates, np.uint32) for _ in range(4)]
    values = np.empty((n_vars, 4), np.uint64)
    pi_pos, pi_vals = np.empty(n_pub, np.uint32), np.empty((n_pub, 4), np.uint64)
    sel_ptrs = (ctypes.c_void_p * 7)(*[a.ctypes.data for a in sel7])
    wire_ptrs = (ctypes.c_void_p * 4)(*[a.ctypes.data for a in wires])
    got = ctypes.c_size_t()
    f = lib().orc_synthetic_circuit
    f.restype = ctypes.c_int
    rc = f(ctypes.c_size_t(n_gates), ctypes.c_uint64(seed), ctypes.c_uint32(n_pub), sel_ptrs, wire_ptrs, ctypes.c_void_p(values.ctypes.data),
           ctypes.c_size_t(n_vars), ctypes.byref(got), ctypes.c_void_p(pi_pos.ctypes.data), ctypes.c_void_p(pi_vals.ctypes.data))
    if rc != 0 or got.value != n_vars:
        raise RuntimeError("orc_synthetic_circuit failed (%d)" % rc)
    sel = [a if a.any() else None for a in sel7] + [None] * 4
    return sel, wires, values, pi_pos, pi_vals
