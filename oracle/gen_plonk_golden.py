"""Generates tests/golden/plonk_kat.json from the pure-Python protocol model (oracle/plonk_model.py).

PARITY UNPINNED (see plonk_model.py): these vectors pin the *model* — and through it the C restatement and the
CUDA prover — to one fixed byte string per circuit; they are not outputs of the Rust reference, which cannot be
built here.  Run: python oracle/gen_plonk_golden.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import plonk_model as pm  # noqa: E402
from model import R, g1_compress  # noqa: E402

TAU = 0x0B200B200B200B200B200B200B200B200B200B200B200B200B200B200B2001
LABEL = b"pb200-plonk-kat"


def range_circuit(value, bits):
    """Arithmetic rows plus one block of range-widget rows proving `value` < 2^bits (bits a multiple of 8)."""
    comp = pm.Composer()
    x = comp.add_input(value)
    y = comp.mul(3, x, x, 5, 0)
    comp.add((1, y), (2, x), 7, 0)
    quads = [(value >> (2 * k)) & 3 for k in range(bits // 2)][::-1]  # most significant first
    acc, accs = 0, [0]
    for qd in quads:
        acc = 4 * acc + qd
        accs.append(acc)
    vars_ = [comp.zero_var] + [comp.add_input(a) for a in accs[1:]]
    rows = len(quads) // 4
    for j in range(rows):
        d, c, b, a = vars_[4 * j], vars_[4 * j + 1], vars_[4 * j + 2], vars_[4 * j + 3]
        comp.poly_gate(a, b, c, d, q_arith=0, q_range=1)
    comp.poly_gate(comp.zero_var, comp.zero_var, comp.zero_var, vars_[4 * rows], q_arith=0)  # carries the last accumulator
    comp.constrain_to_constant(vars_[4 * rows], 0, -value)  # acc_final − 0 + PI = 0
    assert accs[-1] == value
    return comp


def case(name, comp):
    d = pm.domain(comp.n)
    ck = pm.srs_setup(TAU, d["size"])
    pk, vk, tr = pm.preprocess(comp, ck, LABEL)
    proof, pb = pm.prove(comp, pk, ck, tr)
    assert pm.verify(vk, pb, comp.pi, pm.opening_key(TAU), LABEL)
    vk_bytes = b"".join(g1_compress(vk["q"][k]) for k in pm.SELECTORS) + b"".join(g1_compress(c) for c in vk["sigma"])
    return {"name": name, "n_gates": comp.n, "n": d["size"], "proof": pb.hex(), "vk": vk_bytes.hex(),
            "pi": {str(k): format(v, "x") for k, v in comp.pi.items()}}


def main():
    cases = [case("synthetic_13", pm.synthetic_circuit(13)),
             case("synthetic_30", pm.synthetic_circuit(30, seed=0x77, n_pub=3)),
             case("range_16bit", range_circuit(0xB2C7, 16))]
    out = {"tau": format(TAU, "x"), "label": LABEL.decode(), "cases": cases}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "plonk_kat.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
