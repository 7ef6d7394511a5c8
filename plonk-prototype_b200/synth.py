"""The synthetic arithmetic circuit of SURVEY.md §8d as column images for `pb200_preprocess` / `pb200_prove`:
a chain x_{k+1} = x_k·x_k + x_k + c_k built from the `mul` / `add` gate shapes the reference's gadgets use
(/root/reference/src/zk/gadgets.rs:60,70,81), padded with boolean gates, closed by a few public-input rows.
Same rows as driving `StandardComposer` gate by gate, but assembled with numpy so 2^20 gates take seconds."""
import numpy as np

from .prover import R, SELECTORS, scalars_to_mont

_M64 = (1 << 64) - 1


def _random_fr(seed, n):
    """SplitMix64 → uniform scalars by rejection (the generator every test in this repo shares)."""
    out, x = [], seed & _M64

    def nxt():
        nonlocal x
        x = (x + 0x9E3779B97F4A7C15) & _M64
        z = x
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
        return z ^ (z >> 31)

    while len(out) < n:
        l0, l1, l2, l3 = nxt(), nxt(), nxt(), nxt()
        v = l0 | (l1 << 64) | (l2 << 128) | ((l3 & ((1 << 63) - 1)) << 192)
        if v < R:
            out.append(v)
    return out


def synthetic_circuit_columns(n_gates, seed=0x5EED, n_pub=2):
    """Returns (selectors[11], wires[4], values_mont, pi_pos, pi_vals_mont) — assembled by the library's host helper
    `pb200_synthetic_circuit` (C++; identical rows and witness to `synthetic_circuit_columns_py` below)."""
    import ctypes
    from . import _native
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
    rc = _native.lib().pb200_synthetic_circuit(n_gates, seed, n_pub, sel_ptrs, wire_ptrs, values.ctypes.data, n_vars, ctypes.byref(got),
                                               pi_pos.ctypes.data, pi_vals.ctypes.data)
    if rc != 0 or got.value != n_vars:
        raise _native.Pb200Error("pb200_synthetic_circuit failed (%d)" % rc)
    sel = [a if a.any() else None for a in sel7] + [None] * 4
    return sel, wires, values, pi_pos, pi_vals


def synthetic_circuit_columns_py(n_gates, seed=0x5EED, n_pub=2):
    """The same circuit assembled in Python / numpy (the readable definition; used to cross-check the C++ helper)."""
    assert n_gates >= 8 + n_pub
    consts = _random_fr(seed, 4)
    # variables: 0 zero | 1..4 dummy (6, 1, 7, −20) | 5 x_0 | then sq, x' per chain step | extra public-input variables
    values = [0, 6, 1, 7, R - 20, consts[0]]
    steps = (n_gates - n_pub - 3) // 2
    x = consts[0]
    qc_add = []
    for k in range(steps):
        sq = x * x % R
        c = (consts[1] + k) % R
        x = (sq + x + c) % R
        values.append(sq)
        values.append(x)
        qc_add.append(c)
    x_var = 5 + 2 * steps if steps else 5
    n_bool = n_gates - n_pub - 3 - 2 * steps
    w = np.zeros((4, n_gates), dtype=np.uint32)
    q = {k: {} for k in SELECTORS}                     # sparse description: row → small int, filled below
    cols = {k: np.zeros(n_gates, dtype=np.int64) for k in SELECTORS}
    cols["q_arith"][:] = 1
    # row 0: constrained zero; rows 1-2: dummy constraints
    cols["q_l"][0] = 1
    w[:, 1] = (1, 3, 4, 2)
    for k, v in (("q_m", 1), ("q_l", 2), ("q_r", 3), ("q_o", 4), ("q_c", 4), ("q_4", 1)):
        cols[k][1] = v
    w[:, 2] = (4, 1, 3, 0)
    for k, v in (("q_m", 1), ("q_l", 1), ("q_r", 1), ("q_o", 1), ("q_c", 127)):
        cols[k][2] = v
    # chain rows
    r_mul = 3 + 2 * np.arange(steps)
    r_add = r_mul + 1
    xs = 5 + 2 * np.arange(steps)                        # variable holding x_k
    sqs = xs + 1
    xn = xs + 2
    w[0, r_mul], w[1, r_mul], w[2, r_mul] = xs, xs, sqs
    cols["q_m"][r_mul] = 1
    cols["q_o"][r_mul] = -1
    w[0, r_add], w[1, r_add], w[2, r_add] = sqs, xs, xn
    cols["q_l"][r_add] = 1
    cols["q_r"][r_add] = 1
    cols["q_o"][r_add] = -1
    # boolean padding rows on the zero variable
    r0 = 3 + 2 * steps
    cols["q_m"][r0:r0 + n_bool] = 1
    cols["q_o"][r0:r0 + n_bool] = -1
    # public-input rows
    pi_pos, pi_vals = [], []
    for j in range(n_pub):
        row = n_gates - n_pub + j
        if j == 0:
            var, v = x_var, x
        else:
            values.append(consts[2])
            var, v = len(values) - 1, consts[2]
        w[0, row] = w[1, row] = w[2, row] = var
        cols["q_l"][row] = 1
        pi_pos.append(row)
        pi_vals.append((-v) % R)
    # small-integer columns → Montgomery images via a lookup of the few distinct values
    sel = []
    for k in SELECTORS:
        col = cols[k]
        if k == "q_c":
            img = np.zeros((n_gates, 4), dtype=np.uint64)
            img[1] = scalars_to_mont([4])[0]
            img[2] = scalars_to_mont([127])[0]
            if steps:
                img[r_add] = scalars_to_mont(qc_add)
            sel.append(img)
            continue
        if not col.any():
            sel.append(None)
            continue
        img = np.zeros((n_gates, 4), dtype=np.uint64)
        for v in np.unique(col):
            if v:
                img[col == v] = scalars_to_mont([int(v)])[0]
        sel.append(img)
    return (sel, [w[c].copy() for c in range(4)], scalars_to_mont(values), np.asarray(pi_pos, dtype=np.uint32),
            scalars_to_mont(pi_vals))
