"""GPU parity tests of the KZG layer and the prover rounds (csrc/kzg.cu, csrc/plonk.cu) through the C ABI, against
the pure-Python protocol model: byte-identical proofs and verifier keys on the committed golden circuits, and —
at sizes the model cannot prove — acceptance by the model's pairing verifier plus rejection after tampering."""
import os
import sys

import numpy as np
import pytest

import model
import plonk_model as pm
from helpers import load_golden

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import gen_plonk_golden as gen  # noqa: E402

pytestmark = pytest.mark.gpu


def mont(vals):
    import plonk_prototype_b200 as pb
    return pb.scalars_to_mont(vals)


def columns(comp):
    sel = [mont(comp.q[k]) if any(comp.q[k]) else None for k in pm.SELECTORS]
    wires = [np.asarray(w, dtype=np.uint32) for w in comp.w]
    return sel, wires


def gpu_prove(ctx, comp, tau, label):
    """Preprocess + prove the model composer's circuit on the GPU.  Returns (proof bytes, vk bytes)."""
    import plonk_prototype_b200 as pb
    n = pm.domain(comp.n)["size"]
    pp = pb.PublicParameters(n - 1, tau, ctx)
    try:
        sel, wires = columns(comp)
        pk, vk = ctx.preprocess(pp.srs, sel, wires, len(comp.values), label)
        try:
            pis = sorted(comp.pi.items())
            proof = ctx.prove(pp.srs, pk, mont(comp.values), np.asarray([p for p, _ in pis], dtype=np.uint32),
                              mont([v for _, v in pis]) if pis else np.zeros((0, 4), np.uint64))
            again = ctx.prove(pp.srs, pk, mont(comp.values), np.asarray([p for p, _ in pis], dtype=np.uint32),
                              mont([v for _, v in pis]) if pis else np.zeros((0, 4), np.uint64))
            assert again == proof                      # no prover blinding in 0.8.x: proving is deterministic
        finally:
            ctx.prover_key_free(pk)
    finally:
        pp.close()
    return proof, vk


def vk_from_bytes(vkb, n):
    pts = [pm.bytes_to_g1(vkb[48 * i:48 * i + 48]) for i in range(15)]
    return {"n": n, "q": dict(zip(pm.SELECTORS, pts[:11])), "sigma": pts[11:]}


def golden_circuit(case):
    if case["name"] == "synthetic_13":
        return pm.synthetic_circuit(13)
    if case["name"] == "synthetic_30":
        return pm.synthetic_circuit(30, seed=0x77, n_pub=3)
    if case["name"] == "range_16bit":
        return gen.range_circuit(0xB2C7, 16)
    raise KeyError(case["name"])


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_golden_proofs_byte_identical(ctx, idx):
    g = load_golden("plonk_kat.json")
    case = g["cases"][idx]
    comp = golden_circuit(case)
    proof, vk = gpu_prove(ctx, comp, int(g["tau"], 16), g["label"].encode())
    assert vk.hex() == case["vk"]
    assert proof.hex() == case["proof"]


@pytest.mark.parametrize("n_gates", [100, 1000, 5000, (1 << 14) - 3])
def test_proof_accepted_by_pairing_verifier(ctx, n_gates):
    tau, label = 0x5EED0000 + n_gates, b"pb200-verify"
    comp = pm.synthetic_circuit(n_gates, seed=n_gates)
    proof, vkb = gpu_prove(ctx, comp, tau, label)
    n = pm.domain(comp.n)["size"]
    vk = vk_from_bytes(vkb, n)
    ok = pm.opening_key(tau)
    assert pm.verify(vk, proof, comp.pi, ok, label)
    # the library's own verifier (csrc/verify.cu) agrees
    import plonk_prototype_b200 as pb
    items = sorted(comp.pi.items())
    pos, piv = np.asarray([p for p, _ in items], dtype=np.uint32), mont([v for _, v in items])
    bh = pb.opening_key_from_tau(mont([tau]))
    assert pb.verify(vkb, n, label, proof, pos, piv, bh)
    bad = bytearray(proof)
    bad[528 + 5] ^= 0x10                                   # a_eval
    assert not pb.verify(vkb, n, label, bytes(bad), pos, piv, bh)
    assert not pm.verify(vk, bytes(bad), comp.pi, ok, label)
    bad = bytearray(proof)
    bad[48 * 4 + 20] ^= 0x01                               # z_comm (almost surely not a curve point any more)
    assert not pm.verify(vk, bytes(bad), comp.pi, ok, label)
    wrong_pi = dict(comp.pi)
    k = next(iter(wrong_pi))
    wrong_pi[k] = (wrong_pi[k] + 1) % model.R
    assert not pm.verify(vk, proof, wrong_pi, ok, label)


@pytest.mark.parametrize("n_gates", [1 << 10, 5000])
def test_proof_byte_identical_with_c_restatement(ctx, oracle, n_gates):
    """Whole proofs at sizes the Python model cannot reach: the CUDA prover against the C restatement of the upstream
    prover (oracle/plonk_oracle.inc) on the same circuit, witness and SRS."""
    from test_prover_cpu import c_oracle_prove
    tau, label = 0xFEED + n_gates, b"pb200-c-parity"
    comp = pm.synthetic_circuit(n_gates, seed=3 * n_gates)
    proof, vk = gpu_prove(ctx, comp, tau, label)
    want_proof, want_vk = c_oracle_prove(oracle, comp, tau, label, threads=8)
    assert vk == want_vk
    assert proof == want_proof


def test_full_size_proof_2_20_gates_accepted_by_pairing_verifier(ctx):
    """BASELINE.json configs[3] at full size: the verifier's cost does not depend on the circuit size, so the 2^20-gate
    proof is checked by the same pairing verifier (size-independent property)."""
    import plonk_prototype_b200 as pb
    from plonk_prototype_b200.synth import synthetic_circuit_columns
    L, tau, label = 20, 0xB2000014, b"pb200-full"
    n = 1 << L
    sel, wires, values, pi_pos, pi_vals = synthetic_circuit_columns(n)
    pp = pb.PublicParameters(n - 1, tau, ctx)
    pk, vkb = ctx.preprocess(pp.srs, sel, wires, values.shape[0], label)
    proof = ctx.prove(pp.srs, pk, values, pi_pos, pi_vals)
    ctx.prover_key_free(pk)
    pp.close()
    rinv = pow(model.FR_MONT_R, -1, model.R)
    pi = {int(p): model.from_limbs(v) * rinv % model.R for p, v in zip(pi_pos, pi_vals)}
    assert pm.verify(vk_from_bytes(vkb, n), proof, pi, pm.opening_key(tau), label)
    assert pb.verify(vkb, n, label, proof, pi_pos, pi_vals, pb.opening_key_from_tau(pb.scalars_to_mont([tau])))
    bad = bytearray(proof)
    bad[1039] ^= 0x01
    assert not pm.verify(vk_from_bytes(vkb, n), bytes(bad), pi, pm.opening_key(tau), label)


def tiny_circuit(n_gates, n_pub):
    """The composer's 3 fixed rows plus boolean rows on fresh variables and public-input rows: any size ≥ 3."""
    comp = pm.Composer()
    k = 0
    while comp.n < n_gates - n_pub:
        comp.boolean_gate(comp.add_input(k & 1))
        k += 1
    for j in range(n_pub):
        v = comp.add_input(1000 + j)
        comp.constrain_to_constant(v, 0, -(1000 + j))
    assert comp.n == n_gates and comp.check()
    return comp


@pytest.mark.parametrize("n_gates,n_pub", [(3, 0), (4, 1), (5, 0), (7, 2), (8, 0), (9, 1), (15, 3), (16, 0), (16, 2), (17, 1), (31, 0), (32, 4),
                                           (33, 0), (63, 1), (64, 0), (65, 2), (127, 0), (128, 3), (129, 1), (255, 0), (256, 2), (257, 0)])
def test_edge_sizes_byte_identical_with_c_restatement(ctx, oracle, n_gates, n_pub):
    """Ragged and exact power-of-two gate counts, with and without public inputs (domains 4 … 512: every NTT plan from the
    single-pass kernel up, MSMs below the 32-point window switch of upstream)."""
    from test_prover_cpu import c_oracle_prove
    tau, label = 0xED6E + n_gates, b"edge"
    comp = tiny_circuit(n_gates, n_pub)
    proof, vk = gpu_prove(ctx, comp, tau, label)
    want_proof, want_vk = c_oracle_prove(oracle, comp, tau, label, threads=2)
    assert vk == want_vk
    assert proof == want_proof


def test_new_witness_same_key(ctx, oracle):
    """Prover::prove again with another assignment of the same circuit (the key is reused, as upstream's ProverKey is)."""
    import plonk_prototype_b200 as pb
    from test_prover_cpu import c_oracle_prove
    tau, label = 0xAB, b"rewitness"
    comp = tiny_circuit(40, 2)
    pp = pb.PublicParameters(63, tau, ctx)
    sel, wires = columns(comp)
    pk, _ = ctx.preprocess(pp.srs, sel, wires, len(comp.values), label)
    pis = sorted(comp.pi.items())
    pos, piv = np.asarray([p for p, _ in pis], dtype=np.uint32), mont([v for _, v in pis])
    proofs = []
    for flip in (False, True):
        if flip:                                                   # every boolean input inverted: still satisfies the same rows
            for i in range(3, comp.n - 2):
                v = comp.w[0][i]
                comp.values[v] = 1 - comp.values[v]
            assert comp.check()
        got = ctx.prove(pp.srs, pk, mont(comp.values), pos, piv)
        want, _ = c_oracle_prove(oracle, comp, tau, label)
        assert got == want
        proofs.append(got)
    assert proofs[0] != proofs[1]
    ctx.prover_key_free(pk)
    pp.close()


def test_argument_errors_are_reported_not_crashed(ctx):
    import plonk_prototype_b200 as pb
    comp = pm.synthetic_circuit(13)
    sel, wires = columns(comp)
    small = pb.PublicParameters(7, 5, ctx)                              # 8 points for a 16-row circuit
    with pytest.raises(pb.Pb200Error):
        ctx.preprocess(small.srs, sel, wires, len(comp.values), b"x")
    small.close()
    pp = pb.PublicParameters(15, 5, ctx)
    bad_wires = [w.copy() for w in wires]
    bad_wires[2][5] = len(comp.values) + 3                              # unallocated variable
    with pytest.raises(pb.Pb200Error):
        ctx.preprocess(pp.srs, sel, bad_wires, len(comp.values), b"x")
    pk, _ = ctx.preprocess(pp.srs, sel, wires, len(comp.values), b"x")
    with pytest.raises(pb.Pb200Error):                                  # public input outside the domain
        ctx.prove(pp.srs, pk, mont(comp.values), np.asarray([16], dtype=np.uint32), mont([1]))
    ctx.prover_key_free(pk)
    pp.close()


def test_bad_witness_is_rejected_by_verifier(ctx):
    tau, label = 0xBAD, b"pb200-verify"
    comp = pm.synthetic_circuit(200)
    comp.values[comp.w[2][50]] = (comp.values[comp.w[2][50]] + 1) % model.R
    proof, vkb = gpu_prove(ctx, comp, tau, label)
    assert not pm.verify(vk_from_bytes(vkb, 256), proof, comp.pi, pm.opening_key(tau), label)


def test_prover_mirror_api(ctx):
    """The dusk-plonk-shaped Python surface: Prover::new(label) / mut_cs() gadgets / preprocess / prove."""
    import plonk_prototype_b200 as pb
    tau, label = 0xC0FFEE, b"mirror"
    pp = pb.PublicParameters(63, tau, ctx)
    prover = pb.Prover(label, ctx)
    cs = prover.mut_cs()
    a, b = cs.add_input(20), cs.add_input(22)
    s = cs.add((1, a), (1, b), 0, None)
    cs.constrain_to_constant(s, 0, -42)                      # public input: the sum
    for bit in (0, 1, 0, 1):
        cs.boolean_gate(cs.add_input(bit))
    prover.preprocess(pp)
    proof = prover.prove(pp)
    vk = vk_from_bytes(prover.verifier_key_bytes, prover.padded_size)
    pi = {k: v for k, v in cs.public_inputs_sparse_store.items()}
    assert pm.verify(vk, proof, pi, pm.opening_key(tau), label)
    with pytest.raises(RuntimeError):
        prover.preprocess(pp)
    prover.close()
    pp.close()


def test_reference_mock_circuit_valid_balance_proves_and_verifies(ctx):
    """BASELINE.json configs[0], the part this backend can express: MockCircuit::valid_balance
    (/root/reference/src/zk/circuits.rs:51-60 → gadgets.rs:120-145) synthesised on the composer mirror, proved on the
    GPU, accepted by both verifiers."""
    import plonk_prototype_b200 as pb
    G = pb.gadgets
    tau, label = 0x7E57, b"manta-mock-circuit"
    prover = pb.Prover(label, ctx)
    cs = prover.mut_cs()
    note = G.AllocatedScalar.allocate(cs, 1_000_000)
    out = G.MockCircuit(note).valid_balance(cs, 900_000, 21_000)
    assert cs.variables[out] == 1
    cs.constrain_to_constant(out, 0, -1)                     # expose the 0/1 result as a public input
    pp = pb.PublicParameters(cs.circuit_size() + 64, tau, ctx)
    prover.preprocess(pp)
    proof = prover.prove(pp)
    n = prover.padded_size
    pi = dict(cs.public_inputs_sparse_store)
    assert pm.verify(vk_from_bytes(prover.verifier_key_bytes, n), proof, pi, pm.opening_key(tau), label)
    items = sorted(pi.items())
    pos, piv = np.asarray([p for p, _ in items], dtype=np.uint32), mont([v for _, v in items])
    bh = pb.opening_key_from_tau(mont([tau]))
    assert pb.verify(prover.verifier_key_bytes, n, label, proof, pos, piv, bh)
    assert not pb.verify(prover.verifier_key_bytes, n, label, proof, pos, mont([0]), bh)     # claims the check failed
    prover.close()
    pp.close()


def test_cxx_host_layer_example_runs(ctx):
    """examples/mock_circuit.cpp: MockCircuit::valid_balance built with pb200::StandardComposer, proved, verified, tampering
    rejected, EvaluationDomain round trip, KZG commits — all through include/pb200.hpp."""
    import subprocess
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    exe = os.path.join(root, "examples", "mock_circuit")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.join(root, "plonk-prototype_b200", "csrc"), "-s", "example"])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "proof verified" in r.stdout


def test_every_selector_column_is_accepted_and_an_unsatisfied_widget_row_fails_verification(ctx):
    """Round 1 rejected q_logic / q_fixed_group_add / q_variable_group_add; all eleven columns are widgets now.  Switching the
    logic widget on over rows it does not hold for must still prove (the prover does not check satisfiability) but not verify."""
    import plonk_prototype_b200 as pb
    comp = pm.synthetic_circuit(13)
    sel, wires = columns(comp)
    sel[pm.SELECTORS.index("q_logic")] = mont([1] * comp.n)
    tau = 5
    pp = pb.PublicParameters(15, tau, ctx)
    pk, vk = ctx.preprocess(pp.srs, sel, wires, len(comp.values), b"x")
    pis = sorted(comp.pi.items())
    pos, piv = np.asarray([p for p, _ in pis], dtype=np.uint32), mont([v for _, v in pis])
    proof = ctx.prove(pp.srs, pk, mont(comp.values), pos, piv)
    assert not pb.verify(vk, 16, b"x", proof, pos, piv, pb.opening_key_from_tau(mont([tau])))
    ctx.prover_key_free(pk)
    pp.close()


# ---------------------------------------------------------------------------------------------- KZG layer
def test_srs_generate_matches_model(ctx, oracle):
    tau, n = 0x1234567890ABCDEF1234567, 300
    srs = ctx.srs_generate(mont([tau]), n)
    host = np.zeros((n, 12), np.uint64)
    ctx.d2h(host, ctx.srs_dev_ptr(srs))
    ctx.srs_free(srs)
    want = pm.srs_setup(tau, n)
    for i in (0, 1, 2, 17, 255, 256, 299):
        x = model.fp_from_mont(model.from_limbs(host[i, :6]))
        y = model.fp_from_mont(model.from_limbs(host[i, 6:]))
        assert (x, y) == want[i], i


@pytest.mark.parametrize("n", [1, 2, 5, 1024, 1025, 5000])
def test_kzg_witness_matches_ruffini(ctx, n):
    coeffs = model.random_fr(0x6B7A + n, n)
    z = model.random_fr(0x2222, 1)[0]
    p_dev, q_dev = ctx.malloc(32 * n), ctx.malloc(32 * n)
    for point in (z, 0, 1):
        ctx.h2d(p_dev, mont(coeffs))
        ev = ctx.kzg_witness_dev(p_dev, n, mont([point])[0], q_dev)
        got = np.zeros((n, 4), np.uint64)
        ctx.d2h(got, q_dev)
        want = pm.ruffini(coeffs, point) + [0]
        assert (got == mont(want)).all()
        assert (ev == mont([pm.poly_eval(coeffs, point)])[0]).all()
    ctx.free(p_dev)
    ctx.free(q_dev)


def test_fr_horner_step(ctx):
    n, m = 1000, 700
    acc, p = model.random_fr(1, n), model.random_fr(2, m)
    c = model.random_fr(3, 1)[0]
    a_dev, p_dev = ctx.malloc(32 * n), ctx.malloc(32 * m)
    ctx.h2d(a_dev, mont(acc))
    ctx.h2d(p_dev, mont(p))
    ctx.fr_horner_step_dev(a_dev, n, p_dev, m, mont([c])[0])
    got = np.zeros((n, 4), np.uint64)
    ctx.d2h(got, a_dev)
    want = [(acc[j] * c + (p[j] if j < m else 0)) % model.R for j in range(n)]
    assert (got == mont(want)).all()
    ctx.free(a_dev)
    ctx.free(p_dev)


def test_prove_dev_equals_prove_and_profile_sums(ctx):
    """pb200_prove_dev (witness already in HBM) returns the bytes pb200_prove returns; the accumulated profile sums count
    every MSM call of the proof (4 batched launches = 11 commitments)."""
    import plonk_prototype_b200 as pb
    from plonk_prototype_b200.synth import synthetic_circuit_columns
    n = 1 << 12
    sel, wires, values, pi_pos, pi_vals = synthetic_circuit_columns(n)
    pp = pb.PublicParameters(n - 1, 0xD3F, ctx)
    pk, _ = ctx.preprocess(pp.srs, sel, wires, values.shape[0], b"dev")
    want = ctx.prove(pp.srs, pk, values, pi_pos, pi_vals)
    d = ctx.malloc(values.nbytes)
    try:
        ctx.h2d(d, values)
        ctx.profile_enable(True)
        ctx.profile_reset()
        got = ctx.prove_dev(pp.srs, pk, d, pi_pos, pi_vals)
        ms, count = ctx.profile_sum_ms("msm.accumulate")
        ctx.profile_enable(False)
        assert got == want
        assert count == 4 and ms > 0.0
        assert ctx.profile_sum_ms("no.such.timer") == (0.0, 0)
        ctx.profile_reset()
        assert ctx.profile_sum_ms("msm.accumulate") == (0.0, 0)
    finally:
        ctx.free(d)
        ctx.prover_key_free(pk)
        pp.close()


def test_duplicate_public_input_positions_are_rejected(ctx):
    """A gate carries one public input: prover (scatter would race) and verifier (would sum) both refuse duplicates."""
    import plonk_prototype_b200 as pb
    from plonk_prototype_b200.synth import synthetic_circuit_columns
    n = 64
    sel, wires, values, pi_pos, pi_vals = synthetic_circuit_columns(n)
    pp = pb.PublicParameters(n - 1, 0x77, ctx)
    pk, vk = ctx.preprocess(pp.srs, sel, wires, values.shape[0], b"dup")
    try:
        proof = ctx.prove(pp.srs, pk, values, pi_pos, pi_vals)
        dup_pos = np.array([pi_pos[0], pi_pos[0]], dtype=np.uint32)
        with pytest.raises(pb.Pb200Error):
            ctx.prove(pp.srs, pk, values, dup_pos, pi_vals)
        bh = pb.opening_key_from_tau(pb.scalars_to_mont([0x77]))
        assert pb.verify(vk, n, b"dup", proof, pi_pos, pi_vals, bh)
        with pytest.raises(pb.Pb200Error):
            pb.verify(vk, n, b"dup", proof, dup_pos, pi_vals, bh)
    finally:
        ctx.prover_key_free(pk)
        pp.close()


def test_proof_2_20_gates_byte_identical_with_c_restatement(ctx, oracle):
    """BASELINE.json configs[3] at full size, whole proof: the C restatement proves the same 2^20-gate circuit on the
    GPU's own commit key (copied back to the host) with every host thread; the 1040 bytes and the verifier key match."""
    import os
    import plonk_prototype_b200 as pb
    from plonk_prototype_b200.synth import synthetic_circuit_columns
    L, tau, label = 20, 0xB2000014, b"pb200-full-bytes"
    n = 1 << L
    sel, wires, values, pi_pos, pi_vals = synthetic_circuit_columns(n)
    pp = pb.PublicParameters(n - 1, tau, ctx)
    pk, vkb = ctx.preprocess(pp.srs, sel, wires, values.shape[0], label)
    proof = ctx.prove(pp.srs, pk, values, pi_pos, pi_vals)
    srs_host = np.zeros((n, 12), np.uint64)
    ctx.d2h(srs_host, ctx.srs_dev_ptr(pp.srs))
    ctx.prover_key_free(pk)
    pp.close()
    want_proof, want_vk, _, _ = oracle.plonk_prove(sel, wires, values, pi_pos, pi_vals, srs_host, label, threads=os.cpu_count() or 8)
    assert vkb == want_vk
    assert proof == want_proof


# ---------------------------------------------------------------------------------------------- ECC / logic widgets
def prove_composer(ctx, cs, tau, label):
    """Preprocess + prove a circuit built on the shipped StandardComposer mirror.  → (proof, vk, n, pi_pos, pi_vals, columns)."""
    import plonk_prototype_b200 as pb
    n = 1
    while n < cs.n:
        n *= 2
    pp = pb.PublicParameters(n - 1, tau, ctx)
    sel, wires = cs.selector_columns(), cs.wire_columns()
    pk, vk = ctx.preprocess(pp.srs, sel, wires, len(cs.variables), label)
    pis = sorted(cs.public_inputs_sparse_store.items())
    pos = np.asarray([p for p, _ in pis], dtype=np.uint32)
    piv = mont([v for _, v in pis]) if pis else np.zeros((0, 4), np.uint64)
    vals = mont(cs.variables)
    proof = ctx.prove(pp.srs, pk, vals, pos, piv)
    srs_host = np.zeros((n, 12), np.uint64)
    ctx.d2h(srs_host, ctx.srs_dev_ptr(pp.srs))
    ctx.prover_key_free(pk)
    pp.close()
    return proof, vk, n, pos, piv, (sel, wires, vals, srs_host)


@pytest.mark.parametrize("which", ["commitment_gadget", "prove_ownership", "logic"])
def test_reference_ecc_circuits_prove_verify_and_match_c_restatement(ctx, oracle, which):
    """The reference's remaining entry points (/root/reference/src/zk/gadgets.rs:28-41, circuits.rs:63-66) and the logic
    widget: the GPU proof is accepted by the library's pairing verifier, rejected for a wrong public key / tampered
    evaluation, and byte-identical with the C restatement of the upstream prover on the same SRS."""
    import plonk_prototype_b200 as pb
    G, jj = pb.gadgets, pb.jubjub
    cs = pb.StandardComposer()
    if which == "commitment_gadget":
        value, blinder = 0x1234567890ABCDEF, 0xFEDCBA9876543210FEDCBA
        point = G.commitment_gadget(cs, cs.add_input(value), cs.add_input(blinder))
        cs.assert_equal_public_point(point, jj.add(jj.mul(jj.GENERATOR, value), jj.mul(jj.GENERATOR_NUMS, blinder)))
    elif which == "prove_ownership":
        sk = 0x0A11CE5EC2E7
        G.MockCircuit(None, private_key=cs.add_input(sk), public_key=jj.mul(jj.GENERATOR, sk)).prove_ownership(cs)
    else:
        a, b = 0x9E3779B97F4A7C15, 0xBF58476D1CE4E5B9
        x = cs.xor_gate(cs.add_input(a), cs.add_input(b), 64)
        y = cs.and_gate(cs.add_input(a), cs.add_input(b), 64)
        cs.constrain_to_constant(x, 0, -(a ^ b))
        cs.constrain_to_constant(y, a & b, None)
    tau, label = 0xECC0 + len(which), b"pb200-" + which.encode()
    proof, vk, n, pos, piv, (sel, wires, vals, srs_host) = prove_composer(ctx, cs, tau, label)
    bh = pb.opening_key_from_tau(mont([tau]))
    assert pb.verify(vk, n, label, proof, pos, piv, bh)
    if len(pos):
        wrong = piv.copy()
        wrong[0] = mont([12345])[0]
        assert not pb.verify(vk, n, label, proof, pos, wrong, bh)             # another public key / output
    bad = bytearray(proof)
    bad[528 + 4 * 32 + 2] ^= 0x20                                             # a_next_eval: only the new widgets read it
    assert not pb.verify(vk, n, label, bytes(bad), pos, piv, bh)
    want_proof, want_vk, _, _ = oracle.plonk_prove(sel, wires, vals, pos, piv, srs_host, label, threads=8)
    assert vk == want_vk
    assert proof == want_proof


def test_unsatisfied_ecc_witness_does_not_verify(ctx):
    """A wrong private key for the claimed public key: the circuit still proves (the prover does not check), the verifier rejects."""
    import plonk_prototype_b200 as pb
    G, jj = pb.gadgets, pb.jubjub
    cs = pb.StandardComposer()
    G.MockCircuit(None, private_key=cs.add_input(0x1111), public_key=jj.mul(jj.GENERATOR, 0x2222)).prove_ownership(cs)
    proof, vk, n, pos, piv, _ = prove_composer(ctx, cs, 0xBAD, b"bad-key")
    assert not pb.verify(vk, n, b"bad-key", proof, pos, piv, pb.opening_key_from_tau(mont([0xBAD])))


def test_check_hash_inputs_circuit_proves_and_verifies(ctx, oracle):
    """MockCircuit::check_hash_inputs (/root/reference/src/zk/circuits.rs:69-72): the Poseidon sponge circuit (2 permutations,
    2^12 rows) proves on the GPU, verifies against the right public hash only, and matches the C restatement byte for byte."""
    import plonk_prototype_b200 as pb
    cs = pb.StandardComposer()
    inputs = [0xA11CE, 0xB0B, 1 << 250, 42]
    pb.gadgets.MockCircuit(None, hash_inputs=[cs.add_input(v) for v in inputs]).check_hash_inputs(cs, pb.poseidon.hash(inputs))
    tau, label = 0x9051D, b"pb200-poseidon"
    proof, vk, n, pos, piv, (sel, wires, vals, srs_host) = prove_composer(ctx, cs, tau, label)
    assert n == 4096
    bh = pb.opening_key_from_tau(mont([tau]))
    assert pb.verify(vk, n, label, proof, pos, piv, bh)
    assert not pb.verify(vk, n, label, proof, pos, mont([-(pb.poseidon.hash(inputs) + 1)]), bh)
    want_proof, want_vk, _, _ = oracle.plonk_prove(sel, wires, vals, pos, piv, srs_host, label, threads=8)
    assert vk == want_vk and proof == want_proof


def test_polynomial_and_evaluations_wrappers(ctx):
    """`fft::Polynomial` / `fft::Evaluations` mirrors: evaluate and ruffini against the model's Horner /
    Ruffini, interpolate = ifft with the zero top coefficients dropped, commit(&Polynomial) = commit(coefficients)."""
    import plonk_prototype_b200 as pb
    n = 300
    coeffs = model.random_fr(0xA8, n)
    poly = pb.Polynomial.from_coefficients_vec(mont(coeffs + [0, 0, 0]), ctx)
    assert poly.degree() == n - 1 and len(poly) == n and not poly.is_zero()
    z = model.random_fr(0xA9, 1)[0]
    assert (poly.evaluate(mont([z])[0]) == mont([pm.poly_eval(coeffs, z)])[0]).all()
    q = poly.ruffini(mont([z])[0])
    assert q.degree() == n - 2 and (q.coeffs == mont(pm.ruffini(coeffs, z))).all()
    zero = pb.Polynomial.from_coefficients_vec(mont([0, 0]), ctx)
    assert zero.is_zero() and zero.degree() == 0 and not zero.evaluate(mont([z])[0]).any() and zero.ruffini(mont([z])[0]).is_zero()
    # Evaluations: values of a degree-9 polynomial over a 16-point domain interpolate back to its 10 coefficients
    small = model.random_fr(0xAA, 10)
    dom = pb.EvaluationDomain(16, ctx)
    ev = pb.Evaluations.from_vec_and_domain(dom.fft(mont(small)), dom)
    back = ev.interpolate()
    assert back.degree() == 9 and (back.coeffs == mont(small)).all()
    pp = pb.PublicParameters(n - 1, 0xB200, ctx)
    try:
        # (projective triples of two runs may differ — bucket order is not fixed — the group element may not)
        assert pb.g1_to_bytes(ctx.msm(pp.srs, poly.coeffs)) == pb.g1_to_bytes(ctx.msm(pp.srs, mont(coeffs)))
    finally:
        pp.close()
    ck = pb.CommitKey(model_points(8), ctx)
    try:
        small8 = pb.Polynomial.from_coefficients_vec(mont(model.random_fr(0xAB, 8)), ctx)
        assert pb.g1_to_bytes(ck.commit(small8)) == pb.g1_to_bytes(ck.commit(small8.coeffs))
    finally:
        ck.close()


def model_points(n):
    """n packed affine points i·G (Montgomery limbs) from the big-int model."""
    out, cur = [], model.G1_GEN
    for _ in range(n):
        out.append(np.concatenate([mont_fp(cur[0]), mont_fp(cur[1])]))
        cur = model.g1_add(cur, model.G1_GEN)
    return np.stack(out)



def test_public_parameters_raw_bytes_round_trip(ctx):
    """`PublicParameters::to_raw_bytes` → `from_slice_unchecked` (plonk-prototype_b200/serial.py): 240-byte opening key, u64
    count, 97 bytes per point whose first 96 are the Montgomery image of τ^i·G (checked against the model for the first
    points); the reloaded parameters commit to the same bytes and verify the same proof."""
    import plonk_prototype_b200 as pb
    n, tau = 64, 0x7A05
    pp = pb.PublicParameters(n - 1, tau, ctx)
    try:
        raw = pp.to_raw_bytes()
        assert len(raw) == 240 + 8 + 97 * n and int.from_bytes(raw[240:248], "little") == n
        cur = model.G1_GEN
        for i in range(4):
            rec = raw[248 + 97 * i: 248 + 97 * (i + 1)]
            want = np.concatenate([mont_fp(cur[0]), mont_fp(cur[1])])
            assert (np.frombuffer(rec[:96], dtype=np.uint64) == want).all() and rec[96] == 0
            cur = model.g1_mul(cur, tau)
        coeffs = mont(model.random_fr(0x5E, n))
        pp2 = pb.PublicParameters.from_slice_unchecked(raw, ctx)
        try:
            assert pp2.n_points == n and (pp2.beta_h == pp.beta_h).all()
            assert pb.g1_to_bytes(ctx.msm(pp2.srs, coeffs)) == pb.g1_to_bytes(ctx.msm(pp.srs, coeffs))
            var = pb.serial.commit_key_to_var_bytes(ctx, pp2.srs)
            assert len(var) == 48 * n and var[:48] == model.g1_compress(model.G1_GEN) and var[48:96] == model.g1_compress(model.g1_mul(model.G1_GEN, tau))
        finally:
            pp2.close()
        with pytest.raises(ValueError):
            pb.PublicParameters.from_slice_unchecked(raw[:-1], ctx)
    finally:
        pp.close()


def mont_fp(v):
    return np.array([(v * (1 << 384) % model.P >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(6)], dtype=np.uint64)
