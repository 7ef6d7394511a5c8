"""CPU tests of the ECC (fixed-base, variable-base) and logic widgets of the protocol model and of the C restatement:
the JubJub constants are pinned by the curve itself, every widget polynomial vanishes on honest witnesses and not on
tampered ones, the model's proofs verify, and the two CPU implementations (pure-Python model, C restatement) emit
identical bytes.  dusk-plonk's formulas are restated from memory (UPSTREAM_ASSUMPTIONS.md) — self-consistency, not pinned."""
import pytest

import model
import plonk_model as pm

R = model.R


def test_jubjub_constants_are_pinned_by_the_curve():
    def on_curve(p):
        x, y = p
        return (-x * x + y * y) % R == (1 + pm.EDWARDS_D * x * x % R * y * y) % R
    assert (pm.EDWARDS_D * 10241 + 10240) % R == 0
    for g in (pm.JJ_GENERATOR, pm.JJ_GENERATOR_NUMS):
        assert on_curve(g) and pm.jj_mul(g, pm.JJ_ORDER) == (0, 1) and pm.jj_mul(g, 8) != (0, 1)
    assert pm.JJ_GENERATOR[1] == 18
    # the group law used by the widgets: associativity / inverse spot checks
    p, q = pm.jj_mul(pm.JJ_GENERATOR, 12345), pm.jj_mul(pm.JJ_GENERATOR_NUMS, 999)
    assert pm.jj_add(pm.jj_add(p, q), pm.JJ_GENERATOR) == pm.jj_add(p, pm.jj_add(q, pm.JJ_GENERATOR))
    assert pm.jj_add(p, ((-p[0]) % R, p[1])) == (0, 1)


def test_wnaf2_digits():
    for k in (0, 1, 2, 3, 7, 0xDEADBEEF, pm.JJ_ORDER - 1):
        d = pm.wnaf2(k)
        assert len(d) == 256 and set(d) <= {-1, 0, 1}
        assert sum(v << i for i, v in enumerate(d)) == k
        assert all(not (d[i] and d[i + 1]) for i in range(255))          # non-adjacent


def commitment_circuit(value, blinder):
    """gadgets.rs:28-41 `commitment_gadget` + circuits.rs:63-66 `prove_ownership` shape."""
    comp = pm.Composer()
    v, b = comp.add_input(value), comp.add_input(blinder)
    p1 = comp.fixed_base_scalar_mul(v, pm.JJ_GENERATOR)
    p2 = comp.fixed_base_scalar_mul(b, pm.JJ_GENERATOR_NUMS)
    p3 = comp.point_addition_gate(p1, p2)
    want = pm.jj_add(pm.jj_mul(pm.JJ_GENERATOR, value), pm.jj_mul(pm.JJ_GENERATOR_NUMS, blinder))
    comp.assert_equal_public_point(p3, want)
    return comp, p3, want


def logic_circuit(a, b, bits):
    comp = pm.Composer()
    va, vb = comp.add_input(a), comp.add_input(b)
    x = comp.xor_gate(va, vb, bits)
    y = comp.and_gate(va, vb, bits)
    comp.constrain_to_constant(x, 0, -(a ^ b))
    comp.constrain_to_constant(y, a & b, 0)
    return comp, x, y


def test_ecc_gadgets_compute_the_commitment_and_every_row_is_satisfied():
    comp, p3, want = commitment_circuit(0x1234567890ABCDEF, 0xFEDCBA9876543210FEDCBA)
    assert (comp.values[p3[0]], comp.values[p3[1]]) == want
    assert comp.n == 3 + 2 * (3 + 256 + 1 + 1) + 2 + 2
    assert sum(1 for q in comp.q["q_fixed_group_add"] if q) == 512 and sum(1 for q in comp.q["q_variable_group_add"] if q) == 1
    assert comp.check()
    for col, row in ((0, 100), (1, 300), (2, 17), (3, 200)):                 # any touched accumulator breaks a widget row
        bad = pm.Composer.__new__(pm.Composer)
        bad.__dict__ = {k: (list(v) if isinstance(v, list) else v) for k, v in comp.__dict__.items()}
        bad.values = list(comp.values)
        bad.values[comp.w[col][row]] = (bad.values[comp.w[col][row]] + 1) % R
        assert not bad.check(), (col, row)


def test_logic_gadget_computes_xor_and_and():
    comp, x, y = logic_circuit(0xB5C3, 0x6F1A, 16)
    assert comp.values[x] == 0xB5C3 ^ 0x6F1A and comp.values[y] == 0xB5C3 & 0x6F1A
    assert comp.check()
    comp.values[x] ^= 4
    assert not comp.check()


def _fast_commit(oracle):
    """The model's commitments through the C MSM (a 1024-point pure-Python MSM takes minutes); the widget formulas and the
    protocol flow under test stay the model's own."""
    import numpy as np
    cache = {}

    def commit(ck, coeffs):
        key = id(ck)
        if key not in cache:
            cache[key] = np.array([[*model.to_limbs(model.fp_to_mont(p[0]), 6), *model.to_limbs(model.fp_to_mont(p[1]), 6)] for p in ck],
                                  dtype=np.uint64)
        pts = cache[key][:len(coeffs)]
        sc = oracle.fr_to_mont(oracle.ints_to_limbs([c % R for c in coeffs], 4))
        return oracle.g1_proj_to_affine_canonical(oracle.msm_variable_base(pts, sc, threads=4))
    return commit


@pytest.mark.parametrize("which", ["logic", "ecc"])
def test_model_proof_verifies_and_equals_c_restatement(oracle, monkeypatch, which):
    from test_prover_cpu import c_oracle_prove
    if which == "logic":
        comp, _, _ = logic_circuit(0x9E37, 0x79B9, 16)
    else:
        comp, _, _ = commitment_circuit(0xC0FFEE, 0xB200B200B200)
        monkeypatch.setattr(pm, "commit", _fast_commit(oracle))
    assert comp.check()
    tau, label = 0x51D6E7 + len(which), b"widgets-" + which.encode()
    n = pm.domain(comp.n)["size"]
    ck = pm.srs_setup(tau, n) if which == "logic" else _srs_fast(oracle, tau, n)
    pk, vk, tr = pm.preprocess(comp, ck, label)
    _, proof = pm.prove(comp, pk, ck, tr)
    ok = pm.opening_key(tau)
    assert pm.verify(vk, proof, comp.pi, ok, label)
    bad = bytearray(proof)
    bad[528 + 4 * 32 + 3] ^= 1                                                # a_next_eval enters only through the new widgets
    assert not pm.verify(vk, bytes(bad), comp.pi, ok, label)
    c_proof, c_vk = c_oracle_prove(oracle, comp, tau, label, threads=4)
    vk_bytes = b"".join(model.g1_compress(vk["q"][k]) for k in pm.SELECTORS) + b"".join(model.g1_compress(c) for c in vk["sigma"])
    assert c_vk == vk_bytes
    assert c_proof == proof
    # an unsatisfied widget row must not verify
    comp.values[comp.w[0][comp.n // 2]] = (comp.values[comp.w[0][comp.n // 2]] + 1) % R
    assert not comp.check()
    _, bad_proof = pm.prove(comp, pk, ck, tr)
    assert not pm.verify(vk, bad_proof, comp.pi, ok, label)


def _srs_fast(oracle, tau, n):
    """powers_of_g as affine integer pairs via the C oracle's SRS setup (the model's own takes n scalar multiplications)."""
    pts = oracle.srs_setup(oracle.fr_to_mont(oracle.ints_to_limbs([tau], 4))[0], n)
    canon = oracle.fp_from_mont(pts.reshape(-1, 6)).reshape(n, 12)
    return [(oracle.limbs_to_int(r[:6]), oracle.limbs_to_int(r[6:])) for r in canon]


def _as_model_composer(cs):
    """A protocol-model composer holding the rows of a product StandardComposer (for `check()`)."""
    m = pm.Composer.__new__(pm.Composer)
    m.q = {k: list(cs.q[k]) for k in pm.SELECTORS}
    m.w = [list(cs.w_l), list(cs.w_r), list(cs.w_o), list(cs.w_4)]
    m.values = list(cs.variables)
    m.pi = {k: v for k, v in cs.public_inputs_sparse_store.items() if v}
    m.n, m.zero_var = cs.n, cs.zero_var
    return m


def test_poseidon_sponge_gadget_constrains_the_host_hash():
    """plonk-prototype_b200/poseidon.py (restated from memory of dusk-poseidon / dusk-hades — unpinned): the circuit's output
    variable carries `hash(messages)`, every row is satisfied, the end-of-message marker separates lengths, and
    `MockCircuit::check_hash_inputs` (/root/reference/src/zk/circuits.rs:69-72) binds it to a public input."""
    import plonk_prototype_b200 as pb
    P = pb.poseidon
    assert len(P.round_constants()) == 960 and len(set(P.round_constants())) == 960
    m = P.mds_matrix()
    assert all(m[i][j] * (i + 5 + j) % R == 1 for i in range(5) for j in range(5))
    w = P.permutation([1, 2, 3, 4, 5])
    assert len(w) == 5 and w != [1, 2, 3, 4, 5] and P.permutation([1, 2, 3, 4, 5]) == w
    assert len({P.hash([7]), P.hash([7, 0]), P.hash([7, 0, 0, 0]), P.hash([7, 0, 0, 0, 0])}) == 4      # padding is injective in the length
    for msgs in ([3], [3, 1 << 200, 0xDEADBEEF], [1, 2, 3, 4], [R - 1, 5, 6, 7, 8, 9]):
        cs = pb.StandardComposer()
        out = P.gadget(cs, [cs.add_input(v) for v in msgs])
        assert cs.variables[out] == P.hash(msgs)
        assert _as_model_composer(cs).check()
    cs = pb.StandardComposer()
    inputs = [11, 22, 33, 44]
    circuit = pb.gadgets.MockCircuit(None, hash_inputs=[cs.add_input(v) for v in inputs])
    circuit.check_hash_inputs(cs, P.hash(inputs))
    assert _as_model_composer(cs).check() and len(cs.public_inputs_sparse_store) == 1
    bad = pb.StandardComposer()
    pb.gadgets.MockCircuit(None, hash_inputs=[bad.add_input(v) for v in inputs]).check_hash_inputs(bad, P.hash(inputs) + 1)
    assert not _as_model_composer(bad).check()
