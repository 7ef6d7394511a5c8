"""CPU test of the C++ host layer (include/pb200.hpp): its StandardComposer builds, from the same gadget calls, exactly
the column images of the Python mirror (plonk-prototype_b200/prover.py) — which tests/test_prover_cpu.py ties to the model
composer — and BlsScalar's helpers behave like dusk's.  Compiles a small program; no GPU and no ABI call involved."""
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _limbs_hex(row):
    return "".join("%016x" % int(x) for x in row)


def test_cxx_composer_matches_python_mirror(tmp_path):
    import plonk_prototype_b200 as pb
    exe = str(tmp_path / "composer_dump")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-o", exe, os.path.join(ROOT, "tests", "cxx", "composer_dump.cpp")])
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.splitlines()

    R = pb.prover.R
    cs = pb.StandardComposer()
    a, b = cs.add_input(1234), cs.add_input(1200)
    u = cs.add((1, a), (-1, b), 0, None)
    z = cs.add_input(pow(34, -1, R))
    y = cs.mul(-1, z, u, 1, None)
    cs.mul_gate(y, u, u, 1, 0, 0, None)
    acc = cs.add_witness_to_circuit_description(0)
    for k in range(8):
        bit = cs.add_input((34 >> k) & 1)
        cs.boolean_gate(bit)
        acc = cs.add((1 << k, bit), (1, acc), 0, None)
    cs.constrain_to_constant(acc, 0, -34)

    lines = {l.split(" ", 2)[0] + " " + l.split(" ", 2)[1] if l.startswith(("sel", "wire", "pi")) else l.split(" ", 1)[0]: l for l in out}
    assert out[0] == "n_gates %d n_vars %d" % (cs.n, len(cs.variables))
    cols = cs.selector_columns()
    for k in range(11):
        body = lines["sel %d" % k].split(" ")[2:]
        if cols[k] is None:
            assert body == ["null"] or all(int(h, 16) == 0 for h in body), k     # the C++ layer always passes its 7 arithmetic columns
        else:
            assert body == [_limbs_hex(r) for r in cols[k]], k
    for k, w in enumerate(cs.wire_columns()):
        assert [int(x) for x in lines["wire %d" % k].split(" ")[2:]] == list(w), k
    assert lines["vars"].split(" ")[1:] == [_limbs_hex(r) for r in pb.scalars_to_mont(cs.variables)]
    (gate, val), = cs.public_inputs_sparse_store.items()
    assert lines["pi %d" % gate].split(" ")[2] == _limbs_hex(pb.scalars_to_mont([val])[0])
    assert lines["to_bytes"] == "to_bytes 341200"                      # little-endian canonical bytes of 0x1234
    assert lines["reduce"] == "reduce 0 40"                            # 2^70 = limb 1 bit 6
    assert lines["poly"] == "poly 3 2 0 0 1"                           # Polynomial::from_coefficients_vec / degree / is_zero


def test_cxx_ecc_composer_matches_python_mirror(tmp_path):
    """commitment_gadget + assert_equal_public_point through the C++ composer (fixed_base_scalar_mul, point_addition_gate):
    the same 529 rows, 2064 variables, selector images (q_fixed_group_add / q_variable_group_add included) and public inputs
    as the Python mirror — which tests/test_prover_cpu.py ties to the protocol model."""
    import plonk_prototype_b200 as pb
    exe = str(tmp_path / "composer_dump")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-o", exe, os.path.join(ROOT, "tests", "cxx", "composer_dump.cpp")])
    out = subprocess.run([exe, "ecc"], capture_output=True, text=True, check=True).stdout.splitlines()
    G, jj = pb.gadgets, pb.jubjub
    cs = pb.StandardComposer()
    value, blinder = cs.add_input(0xC0FFEE), cs.add_input(0xB200B200B200)
    p3 = G.commitment_gadget(cs, value, blinder)
    cs.assert_equal_public_point(p3, (cs.variables[p3[0]], cs.variables[p3[1]]))
    assert (cs.variables[p3[0]], cs.variables[p3[1]]) == jj.add(jj.mul(jj.GENERATOR, 0xC0FFEE), jj.mul(jj.GENERATOR_NUMS, 0xB200B200B200))
    lines = {l.split(" ", 2)[0] + " " + l.split(" ", 2)[1] if l.startswith(("sel", "wire", "pi")) else l.split(" ", 1)[0]: l for l in out}
    assert out[0] == "n_gates %d n_vars %d" % (cs.n, len(cs.variables))
    cols = cs.selector_columns()
    for k in range(11):
        body = lines["sel %d" % k].split(" ")[2:]
        if cols[k] is None:
            assert body == ["null"] or all(int(h, 16) == 0 for h in body), k
        else:
            assert body == [_limbs_hex(r) for r in cols[k]], k
    assert cols[9] is not None and cols[10] is not None
    for k, w in enumerate(cs.wire_columns()):
        assert [int(x) for x in lines["wire %d" % k].split(" ")[2:]] == list(w), k
    assert lines["vars"].split(" ")[1:] == [_limbs_hex(r) for r in pb.scalars_to_mont(cs.variables)]
    for gate, val in cs.public_inputs_sparse_store.items():
        assert lines["pi %d" % gate].split(" ")[2] == _limbs_hex(pb.scalars_to_mont([val])[0])
