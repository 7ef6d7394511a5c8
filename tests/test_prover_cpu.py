"""CPU tests of the protocol layer above the hot path: the pure-Python model (oracle/plonk_model.py) against its
external anchors and committed golden proofs, and the host-side pieces of the product (Merlin transcript in
libpb200.so, StandardComposer mirror).  No GPU compute here."""
import ctypes
import os
import sys

import numpy as np
import pytest

import model
import plonk_model as pm
from helpers import load_golden

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import gen_plonk_golden as gen  # noqa: E402

MERLIN_KAT = "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"


def test_merlin_known_answer_model():
    """merlin's own `equivalence_simple` vector."""
    t = pm.Transcript(b"test protocol")
    t.append_message(b"some label", b"some data")
    assert t.challenge_bytes(b"challenge", 32).hex() == MERLIN_KAT


def test_merlin_known_answer_library():
    """The host transcript compiled into libpb200.so (csrc/merlin.h) — loads the library, needs no GPU."""
    import plonk_prototype_b200 as pb
    from plonk_prototype_b200 import _native
    out = ctypes.create_string_buffer(32)
    rc = _native.lib().pb200_transcript_selftest(b"test protocol", b"some label", b"some data", 9, b"challenge", out, 32)
    assert rc == 0 and out.raw.hex() == MERLIN_KAT
    # a multi-block squeeze and absorb (crosses the 166-byte STROBE rate) against the model
    msg = bytes(range(256)) * 3
    out = ctypes.create_string_buffer(400)
    assert _native.lib().pb200_transcript_selftest(b"pb200", b"blob", msg, len(msg), b"long", out, 400) == 0
    t = pm.Transcript(b"pb200")
    t.append_message(b"blob", msg)
    assert t.challenge_bytes(b"long", 400) == out.raw


def test_g2_generator_and_pairing():
    assert pm.g2_on_curve(pm.G2_GEN) and pm.g2_mul(pm.G2_GEN, model.R) is None
    a, b = 0x1234567, 0x89ABCDEF01
    e_ab = pm.final_exponentiation(pm.miller_loop(model.g1_mul(model.G1_GEN, a), pm.g2_mul(pm.G2_GEN, b)))
    e = pm.final_exponentiation(pm.miller_loop(model.G1_GEN, pm.G2_GEN))
    assert e != pm.F12_ONE and e_ab == pm.f12_pow(e, a * b % model.R)
    assert pm.f12_pow(e, model.R) == pm.F12_ONE
    # e(aG, H)·e(−G, aH) = 1: the shape of the KZG check
    assert pm.pairing_product_is_one([(model.g1_mul(model.G1_GEN, a), pm.G2_GEN),
                                      (model.g1_neg(model.G1_GEN), pm.g2_mul(pm.G2_GEN, a))])


def test_sigma_permutation_is_a_permutation_with_variable_cycles():
    comp = pm.synthetic_circuit(13)
    sig = pm.sigma_positions(comp.w, 16)
    flat = [p for col in sig for p in col]
    assert sorted(flat) == [(c, i) for c in range(4) for i in range(16)]
    for c in range(4):
        for i in range(comp.n):
            c2, i2 = sig[c][i]
            assert comp.w[c2][i2] == comp.w[c][i]      # stays on the same variable
        for i in range(comp.n, 16):
            assert sig[c][i] == (c, i)                  # identity on padding


def test_model_reproduces_golden_proof_and_verifier_rejects_tampering():
    g = load_golden("plonk_kat.json")
    case = g["cases"][0]
    tau, label = int(g["tau"], 16), g["label"].encode()
    comp = pm.synthetic_circuit(case["n_gates"])
    ck = pm.srs_setup(tau, case["n"])
    pk, vk, tr = pm.preprocess(comp, ck, label)
    _, pb = pm.prove(comp, pk, ck, tr)
    assert pb.hex() == case["proof"]
    ok = pm.opening_key(tau)
    assert pm.verify(vk, pb, comp.pi, ok, label)
    bad = bytearray(pb)
    bad[700] ^= 0x01                                      # one evaluation bit
    assert not pm.verify(vk, bytes(bad), comp.pi, ok, label)
    wrong_pi = {k: (v + 1) % model.R for k, v in comp.pi.items()}
    assert not pm.verify(vk, pb, wrong_pi, ok, label)
    assert not pm.verify(vk, pb, comp.pi, ok, b"another transcript")


def test_unsatisfied_circuit_does_not_verify():
    comp = pm.synthetic_circuit(13)
    comp.values[comp.w[2][5]] = (comp.values[comp.w[2][5]] + 1) % model.R   # break one gate
    assert not comp.check()
    tau, label = 0xABCDEF, b"x"
    ck = pm.srs_setup(tau, 16)
    pk, vk, tr = pm.preprocess(comp, ck, label)
    _, pb = pm.prove(comp, pk, ck, tr)
    assert not pm.verify(vk, pb, comp.pi, pm.opening_key(tau), label)


def test_product_composer_matches_model_composer():
    """The shipped StandardComposer mirror and the model composer build identical columns from the gadget calls the
    reference makes (gadgets.rs:49-84 `maybe_equal`, :120-135 `min_bound`, boolean decomposition :210-220)."""
    import plonk_prototype_b200 as pb

    def build(c, pi_none):
        a, b = c.add_input(1234), c.add_input(1200)
        u = c.add((1, a), (-1, b), 0, pi_none)
        zinv = c.add_input(pow(34, -1, model.R))
        y = c.mul(-1, zinv, u, 1, pi_none)
        c.mul_gate(y, u, u, 1, 0, 0, pi_none)
        bits = [c.add_input((34 >> k) & 1) for k in range(8)]
        acc = c.add_witness_to_circuit_description(0)
        for k, bit in enumerate(bits):
            c.boolean_gate(bit)
            acc = c.add((1 << k, bit), (1, acc), 0, pi_none)
        c.constrain_to_constant(acc, 0, -34)
        return c

    m = build(pm.Composer(), 0)
    p = build(pb.StandardComposer(), None)
    assert m.check()
    assert p.n == m.n and p.variables == m.values
    assert [list(x) for x in (p.w_l, p.w_r, p.w_o, p.w_4)] == m.w
    for k in pm.SELECTORS:
        assert p.q[k] == m.q[k], k
    assert {k: v for k, v in p.public_inputs_sparse_store.items() if v} == m.pi
    cols = p.selector_columns()
    assert cols[pm.SELECTORS.index("q_logic")] is None
    want = np.array([model.to_limbs(model.fr_to_mont(v), 4) for v in m.q["q_c"]], dtype=np.uint64)
    assert (cols[pm.SELECTORS.index("q_c")] == want).all()


def test_range_circuit_model():
    comp = gen.range_circuit(0xB2C7, 16)
    assert comp.check()
    assert any(comp.q["q_range"])


# ---------------------------------------------------------------------------------------------- C restatement
def _mont(oracle, vals):
    return oracle.fr_to_mont(oracle.ints_to_limbs([v % model.R for v in vals], 4))


def c_oracle_prove(oracle, comp, tau, label, threads=1, srs=None):
    n = pm.domain(comp.n)["size"]
    if srs is None:
        srs = oracle.srs_setup(_mont(oracle, [tau])[0], n)
    sel = [_mont(oracle, comp.q[k]) if any(comp.q[k]) else None for k in pm.SELECTORS]
    wires = [np.asarray(w, dtype=np.uint32) for w in comp.w]
    pis = sorted(comp.pi.items())
    proof, vk, _, _ = oracle.plonk_prove(sel, wires, _mont(oracle, comp.values), np.asarray([p for p, _ in pis], dtype=np.uint32),
                                        _mont(oracle, [v for _, v in pis]) if pis else np.zeros((0, 4), np.uint64), srs, label,
                                        threads=threads)
    return proof, vk


def test_c_oracle_merlin_known_answer(oracle):
    assert oracle.merlin_selftest(b"test protocol", b"some label", b"some data", b"challenge", 32).hex() == MERLIN_KAT


@pytest.mark.parametrize("threads", [1, 4])
def test_c_oracle_prover_reproduces_golden_proofs(oracle, threads):
    """The C restatement (the CPU baseline bench.py times) is byte-identical with the big-int model."""
    g = load_golden("plonk_kat.json")
    tau, label = int(g["tau"], 16), g["label"].encode()
    circuits = [pm.synthetic_circuit(13), pm.synthetic_circuit(30, seed=0x77, n_pub=3), gen.range_circuit(0xB2C7, 16)]
    for case, comp in zip(g["cases"], circuits):
        proof, vk = c_oracle_prove(oracle, comp, tau, label, threads)
        assert vk.hex() == case["vk"], case["name"]
        assert proof.hex() == case["proof"], case["name"]


def test_c_oracle_proof_verifies_at_2k_gates(oracle):
    tau, label = 0x7A5, b"c-oracle"
    comp = pm.synthetic_circuit(2000, seed=9)
    proof, vkb = c_oracle_prove(oracle, comp, tau, label, threads=8)
    pts = [pm.bytes_to_g1(vkb[48 * i:48 * i + 48]) for i in range(15)]
    vk = {"n": 2048, "q": dict(zip(pm.SELECTORS, pts[:11])), "sigma": pts[11:]}
    assert pm.verify(vk, proof, comp.pi, pm.opening_key(tau), label)


def test_cxx_synthetic_circuit_helper_matches_python_definition_and_model():
    """pb200_synthetic_circuit (host C++ in libpb200.so, no GPU) builds the same columns as the numpy definition and as
    driving the model composer gate by gate."""
    import plonk_prototype_b200 as pb
    from plonk_prototype_b200.synth import synthetic_circuit_columns, synthetic_circuit_columns_py
    for ng, npub in ((13, 2), (30, 3), (257, 1), (1000, 2)):
        a, b = synthetic_circuit_columns(ng, n_pub=npub), synthetic_circuit_columns_py(ng, n_pub=npub)
        for k in range(11):
            assert (a[0][k] is None) == (b[0][k] is None)
            assert a[0][k] is None or (a[0][k] == b[0][k]).all()
        assert all((x == y).all() for x, y in zip(a[1], b[1]))
        assert (a[2] == b[2]).all() and (a[3] == b[3]).all() and (a[4] == b[4]).all()
    comp = pm.synthetic_circuit(30, n_pub=3)
    sel, wires, values, pi_pos, pi_vals = synthetic_circuit_columns(30, n_pub=3)
    assert [list(w) for w in wires] == comp.w
    assert (values == pb.scalars_to_mont(comp.values)).all()
    assert list(pi_pos) == sorted(comp.pi)


# ---------------------------------------------------------------------------------------------- C++ verifier (product, host)
def _pi_arrays(pi):
    import plonk_prototype_b200 as pb
    items = sorted(pi.items())
    return np.asarray([p for p, _ in items], dtype=np.uint32), pb.scalars_to_mont([v for _, v in items])


def test_library_pairing_selftest_and_opening_key():
    """csrc/verify.cu: bilinearity + non-degeneracy of the pairing, and τ·H equal to the model's."""
    import plonk_prototype_b200 as pb
    from plonk_prototype_b200 import _native
    ok = ctypes.c_int(0)
    a, b = pb.scalars_to_mont([0x1234567]), pb.scalars_to_mont([model.R - 5])
    assert _native.lib().pb200_pairing_selftest(a.ctypes.data, b.ctypes.data, ctypes.byref(ok)) == 0 and ok.value == 1
    tau = 0xDEADBEEFCAFE
    bh = pb.opening_key_from_tau(pb.scalars_to_mont([tau]))
    mh = pm.g2_mul(pm.G2_GEN, tau)
    got = [model.fp_from_mont(model.from_limbs(bh[6 * i:6 * i + 6])) for i in range(4)]
    assert got == [mh[0][0], mh[0][1], mh[1][0], mh[1][1]]


def test_library_verifier_agrees_with_model_on_golden_proofs():
    """pb200_verify (C++ tower + Miller loop + transcript) and the Python model verifier are independent implementations:
    both accept the golden proofs and both reject the same tamperings."""
    import plonk_prototype_b200 as pb
    g = load_golden("plonk_kat.json")
    tau, label = int(g["tau"], 16), g["label"].encode()
    bh = pb.opening_key_from_tau(pb.scalars_to_mont([tau]))
    for case in g["cases"]:
        pi = {int(k): int(v, 16) for k, v in case["pi"].items()}
        pos, piv = _pi_arrays(pi)
        proof, vk = bytes.fromhex(case["proof"]), bytes.fromhex(case["vk"])
        assert pb.verify(vk, case["n"], label, proof, pos, piv, bh), case["name"]
        for off in (3, 48 * 4 + 7, 48 * 9 + 40, 528, 700, 1039):       # commitments, witnesses, evaluations
            bad = bytearray(proof)
            bad[off] ^= 0x04
            assert not pb.verify(vk, case["n"], label, bytes(bad), pos, piv, bh), (case["name"], off)
        assert not pb.verify(vk, case["n"], b"another label", proof, pos, piv, bh)
        assert not pb.verify(vk, case["n"], label, proof, pos, pb.scalars_to_mont([v + 1 for _, v in sorted(pi.items())]), bh)
        wrong_key = pb.opening_key_from_tau(pb.scalars_to_mont([tau + 1]))
        assert not pb.verify(vk, case["n"], label, proof, pos, piv, wrong_key)
        non_canonical = bytearray(proof)
        non_canonical[528:560] = (int.from_bytes(proof[528:560], "little") + model.R).to_bytes(32, "little")
        assert not pb.verify(vk, case["n"], label, bytes(non_canonical), pos, piv, bh)


# ---------------------------------------------------------------------------------------------- the reference's gadgets
def test_reference_gadget_semantics_on_the_composer_mirror():
    """gadgets.rs / circuits.rs restated (plonk-prototype_b200/gadgets.py): outputs and row counts of the reference's
    arithmetic gadgets (min_bound with 30 bits ≈ 65 gates)."""
    import plonk_prototype_b200 as pb
    G = pb.gadgets
    assert [G.bits_count(v) for v in (0, 1, 2, 3, 4, 255, 256)] == [1, 1, 2, 2, 3, 8, 9]
    assert G.num_bits_closest_power_of_two(1000) == 11
    bits = G.scalar_to_bits(0b1011)
    assert bits[:5] == [1, 1, 0, 1, 0] and len(bits) == 256

    def run(note, tx, gas):
        cs = pb.StandardComposer()
        before_rows, before_vars = cs.n, len(cs.variables)
        out = G.MockCircuit(G.AllocatedScalar.allocate(cs, note)).valid_balance(cs, tx, gas)
        return cs, cs.variables[out], cs.n - before_rows, len(cs.variables) - before_vars

    cs, out, rows, nvars = run(1000, 700, 50)
    assert out == 1 and rows == 1 + 1 + 2 * 30 + 3            # add, zero constraint, (boolean + add) per bit, maybe_equal
    assert nvars == 1 + 1 + 256 + 1 + 30 + 3                  # note, x−a, 256 bits, accumulator zero, 30 sums, u z y
    assert run(700, 700, 50)[1] == 0                          # 700 − 750 wraps: not a 30-bit quantity
    assert run(750, 700, 50)[1] == 1

    # every row of the built circuit is satisfied (the same check the model composer does)
    v = cs.variables
    for i in range(cs.n):
        a, b, c, d = v[cs.w_l[i]], v[cs.w_r[i]], v[cs.w_o[i]], v[cs.w_4[i]]
        q = {k: cs.q[k][i] for k in cs.q}
        t = q["q_arith"] * (q["q_m"] * a * b + q["q_l"] * a + q["q_r"] * b + q["q_o"] * c + q["q_4"] * d + q["q_c"])
        assert (t + cs.public_inputs_sparse_store.get(i, 0)) % model.R == 0, i

    cs2 = pb.StandardComposer()
    w = G.AllocatedScalar.allocate(cs2, 500)
    assert cs2.variables[G.range_check(cs2, 100, 1000, w)] == 1
    cs3 = pb.StandardComposer()
    assert cs3.variables[G.range_check(cs3, 100, 1000, G.AllocatedScalar.allocate(cs3, 1000))] == 0   # max is exclusive


def test_reference_ecc_gadgets_on_the_composer_mirror():
    """gadgets.rs:28-41 `commitment_gadget` and circuits.rs:63-66 `prove_ownership` through the shipped composer mirror:
    same rows as the protocol model's composer, the committed point is value·G + blinder·G_nums, all rows satisfied."""
    import plonk_prototype_b200 as pb
    G, jj = pb.gadgets, pb.jubjub
    value, blinder = 0xC0FFEE, 0xB200B200B200
    cs = pb.StandardComposer()
    point = G.commitment_gadget(cs, cs.add_input(value), cs.add_input(blinder))
    assert (cs.variables[point[0]], cs.variables[point[1]]) == jj.add(jj.mul(jj.GENERATOR, value), jj.mul(jj.GENERATOR_NUMS, blinder))
    m = pm.Composer()
    mv, mb = m.add_input(value), m.add_input(blinder)
    mp = m.point_addition_gate(m.fixed_base_scalar_mul(mv, pm.JJ_GENERATOR), m.fixed_base_scalar_mul(mb, pm.JJ_GENERATOR_NUMS))
    assert mp == point and cs.n == m.n == 3 + 2 * 261 + 2 and cs.variables == m.values
    assert [list(x) for x in (cs.w_l, cs.w_r, cs.w_o, cs.w_4)] == m.w
    for k in pm.SELECTORS:
        assert cs.q[k] == m.q[k], k
    assert m.check()
    sk = 0x1234567
    c2 = pb.StandardComposer()
    G.MockCircuit(None, private_key=c2.add_input(sk), public_key=jj.mul(jj.GENERATOR, sk)).prove_ownership(c2)
    assert c2.n == 3 + 261 + 2 and sorted(c2.public_inputs_sparse_store) == [c2.n - 2, c2.n - 1]
    c_bad = pb.StandardComposer()
    with pytest.raises(ValueError):                                         # JubJubScalar::from_bytes(..).unwrap() upstream
        c_bad.fixed_base_scalar_mul(c_bad.add_input(jj.JJ_ORDER), jj.GENERATOR)
    # logic gates: same rows as the model, right outputs
    c3, m3 = pb.StandardComposer(), pm.Composer()
    x = c3.xor_gate(c3.add_input(0xB5C3), c3.add_input(0x6F1A), 16)
    y = c3.and_gate(c3.add_input(0xB5C3), c3.add_input(0x6F1A), 16)
    m3.xor_gate(m3.add_input(0xB5C3), m3.add_input(0x6F1A), 16)
    m3.and_gate(m3.add_input(0xB5C3), m3.add_input(0x6F1A), 16)
    assert c3.variables[x] == 0xB5C3 ^ 0x6F1A and c3.variables[y] == 0xB5C3 & 0x6F1A
    assert c3.variables == m3.values and [list(w) for w in (c3.w_l, c3.w_r, c3.w_o, c3.w_4)] == m3.w and m3.check()
    for k in pm.SELECTORS:
        assert c3.q[k] == m3.q[k], k


def test_polynomial_wrapper_host_logic():
    """`Polynomial::from_coefficients_vec` truncates the zero coefficients at the top; `degree()` is 0 for the zero polynomial
    (dusk-plonk 0.8 `fft/polynomial.rs`).  No GPU involved: the arithmetic entry points are covered by the -m gpu tests."""
    import numpy as np
    import plonk_prototype_b200 as pb
    c = np.zeros((6, 4), np.uint64)
    c[0, 0], c[3, 2] = 5, 7
    p = pb.Polynomial.from_coefficients_vec(c)
    assert len(p) == 4 and p.degree() == 3 and not p.is_zero()
    z = pb.Polynomial.from_coefficients_vec(np.zeros((3, 4), np.uint64))
    assert z.is_zero() and z.degree() == 0 and len(z) == 0 and pb.Polynomial.zero().is_zero()


def test_opening_key_bytes_round_trip():
    """`OpeningKey::to_bytes` / from_bytes (plonk-prototype_b200/serial.py): G ‖ H ‖ β·H in the compressed zcash encodings.
    β·H comes from the library's host-side G2 arithmetic (pb200_opening_key_from_tau, no GPU) and must equal the model's
    τ·H; decompression (an Fp2 square root) must return the same limbs."""
    import plonk_prototype_b200 as pb
    from plonk_prototype_b200 import serial
    for tau in (5, 0xB200B200B200, pm.R - 2):
        beta_h = pb.opening_key_from_tau(pb.scalars_to_mont([tau])[0])
        b = serial.opening_key_to_bytes(beta_h)
        assert len(b) == 240 and b[:48] == model.g1_compress(model.G1_GEN)
        want = pm.g2_mul(pm.G2_GEN, tau)
        x1, x0 = int.from_bytes(bytes([b[144] & 0x1F]) + b[145:192], "big"), int.from_bytes(b[192:240], "big")
        assert (x0, x1) == want[0]
        assert (serial.opening_key_from_bytes(b) == beta_h).all()
        gx1, gx0 = int.from_bytes(bytes([b[48] & 0x1F]) + b[49:96], "big"), int.from_bytes(b[96:144], "big")
        assert (gx0, gx1) == pm.G2_GEN[0]
    bad = bytearray(b)
    bad[200] ^= 1
    try:
        serial.opening_key_from_bytes(bytes(bad))
        raised = False
    except AssertionError:
        raised = True
    # a flipped bit gives either a point off the curve (assertion) or another point — never the same key
    assert raised or not (serial.opening_key_from_bytes(bytes(bad)) == beta_h).all()
