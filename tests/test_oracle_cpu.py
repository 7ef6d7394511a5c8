"""CPU tests: the C oracle against the big-int model's golden vectors, public constants and algebraic
invariants.  The reference has no tests to borrow; PARITY IS UNPINNED and anchored on
two independent implementations agreeing (oracle.c ↔ model.py) plus public BLS12-381 constants."""
import numpy as np
import pytest

import model
from helpers import hex_to_limbs, limbs_to_hex, load_golden, sha_scalars_canonical


def test_public_constants(oracle):
    c = oracle.fr_consts()
    assert oracle.limbs_to_int(c["modulus"]) == model.R
    assert c["inv"] == 0xFFFFFFFEFFFFFFFF
    assert oracle.limbs_to_int(c["r1"]) == 0x1824B159ACC5056F998C4FEFECBC4FF55884B7FA0003480200000001FFFFFFFE
    assert oracle.limbs_to_int(c["r2"]) == 0x0748D9D99F59FF1105D314967254398F2B6CEDCB87925C23C999E990F3F29C6D
    assert [int(x) for x in c["root_of_unity"]] == [0xB9B58D8C5F0E466A, 0x5B1B4C801819D7EC, 0x0AF53AE352A31E64,
                                                    0x5BF3ADDA19E9B27B]
    p = oracle.fp_consts()
    assert oracle.limbs_to_int(p["modulus"]) == model.P
    assert p["inv"] == 0x89F3FFFCFFFCFFFD
    assert oracle.limbs_to_int(p["r2"]) == int(
        "11988fe592cae3aa9a793e85b519952d67eb88a9939d83c08de5476c4c95b6d50a76e6a609d104f1f4df1f341c341746", 16)
    # root of unity has order exactly 2^32
    w = model.ROOT_OF_UNITY
    assert pow(w, 1 << 32, model.R) == 1 and pow(w, 1 << 31, model.R) != 1
    # generator: on curve, order r, and the well-known compressed encoding
    assert model.g1_on_curve(model.G1_GEN) and model.g1_mul(model.G1_GEN, model.R - 1) == model.g1_neg(model.G1_GEN)
    assert model.g1_compress(model.G1_GEN).hex().startswith("97f1d3a73197d7942695638c4fa9ac0f")
    assert load_golden("bases_kat.json")["generator_compressed"] == model.g1_compress(model.G1_GEN).hex()


def test_prng_and_field_ops_match_model(oracle):
    a, b = oracle.random_fr(11, 300), oracle.random_fr(12, 300)
    am, bm = model.random_fr(11, 300), model.random_fr(12, 300)
    assert [oracle.limbs_to_int(x) for x in a] == am
    prod = oracle.fr_from_mont(oracle.fr_mul(oracle.fr_to_mont(a), oracle.fr_to_mont(b)))
    assert [oracle.limbs_to_int(x) for x in prod] == [x * y % model.R for x, y in zip(am, bm)]
    inv = oracle.fr_from_mont(oracle.fr_inv(oracle.fr_to_mont(a[:20])))
    assert [oracle.limbs_to_int(x) for x in inv] == [pow(x, -1, model.R) for x in am[:20]]


def _ntt_input(oracle, case):
    if "seed" in case and "structured" not in case:
        return model.random_fr(case["seed"], 1 << case["log_n"])
    n = 1 << case["log_n"]
    return {"delta0": [1] + [0] * (n - 1), "all_r_minus_1": [model.R - 1] * n}.get(case.get("structured"))


@pytest.mark.parametrize("threads", [1, 4])
def test_ntt_golden(oracle, threads):
    g = load_golden("ntt_kat.json")
    for case in g["cases"]:
        n = 1 << case["log_n"]
        if case.get("structured") == "short37_zero_padded":
            x = model.random_fr(case["seed"], 37) + [0] * (n - 37)
        else:
            x = _ntt_input(oracle, case)
        xm = oracle.fr_to_mont(oracle.ints_to_limbs(x, 4))
        for name, inv, cos in (("fft", 0, 0), ("ifft", 1, 0), ("coset_fft", 0, 1), ("coset_ifft", 1, 1)):
            key = name + "_sha256"
            if key not in case:
                continue
            y = oracle.fr_from_mont(oracle.ntt(xm, inv, cos, threads))
            assert sha_scalars_canonical(y) == case[key], (case, name)
            if name + "_head" in case:
                assert [limbs_to_hex(v, 32) for v in y[:4]] == case[name + "_head"]


def test_ntt_invariants(oracle):
    for log_n in (3, 8, 11):
        n = 1 << log_n
        x = oracle.fr_to_mont(oracle.random_fr(500 + log_n, n))
        assert (oracle.ntt(oracle.ntt(x, 0, 0), 1, 0) == x).all()
        assert (oracle.ntt(oracle.ntt(x, 0, 1), 1, 1) == x).all()
        y = oracle.fr_to_mont(oracle.random_fr(600 + log_n, n))
        assert (oracle.ntt(oracle.fr_add(x, y)) == oracle.fr_add(oracle.ntt(x), oracle.ntt(y))).all()  # linearity
    # EvaluationDomain::new fails for log_n >= 32
    assert oracle.lib().orc_ntt(None, 32, 0, 0, 1) == -1
    with pytest.raises(ValueError):
        model.domain(1 << 32)


def _scalars_for(case, n):
    name = case["name"]
    if name.startswith("random_") or name in ("all_equal_bases", "plus_minus_pairs"):
        return model.random_fr(case["seed"], n)
    if name == "all_zero":
        return [0] * n
    if name == "all_one":
        return [1] * n
    if name == "all_r_minus_1":
        return [model.R - 1] * n
    if name == "eight_bit":
        return [v & 0xFF for v in model.random_fr(case["seed"], n)]
    if name == "half_zero":
        return [0 if i % 2 else v for i, v in enumerate(model.random_fr(case["seed"], n))]
    if name == "plus_minus_pairs_equal_scalars":
        return [v for v in model.random_fr(case["seed"], n // 2) for _ in (0, 1)]
    if name == "powers_of_two":
        return [(1 << k) % model.R for k in range(0, 255, 5)] + [((1 << k) - 1) % model.R for k in range(1, 255, 7)]
    raise KeyError(name)


def msm_case_inputs(oracle, case):
    """(points (n,12) Montgomery, scalars (n,4) Montgomery) for a golden MSM case."""
    n = case["n"]
    pts = oracle.synthetic_bases(n)
    if case["bases"].startswith("synthetic[0]"):
        pts = np.repeat(pts[:1], n, axis=0)
    elif case["bases"].startswith("P0,-P0"):
        half = pts[: n // 2]
        neg = half.copy()
        zero = np.zeros((n // 2, 6), np.uint64)
        neg[:, 6:] = oracle.fp_sub(zero, half[:, 6:])
        pts = np.stack([half, neg], axis=1).reshape(n, 12)
    s = oracle.fr_to_mont(oracle.ints_to_limbs(_scalars_for(case, n), 4))
    return np.ascontiguousarray(pts), s


def expected_affine(case):
    return None if case["x"] is None else (int(case["x"], 16), int(case["y"], 16))


@pytest.mark.parametrize("threads", [1, 3])
def test_msm_golden(oracle, threads):
    for case in load_golden("msm_kat.json")["cases"]:
        pts, s = msm_case_inputs(oracle, case)
        assert oracle.g1_on_curve(pts)
        got = oracle.g1_proj_to_affine_canonical(oracle.msm_variable_base(pts, s, threads))
        assert got == expected_affine(case), case["name"]
        assert model.g1_compress(got).hex() == case["compressed"]


def test_synthetic_bases_golden(oracle):
    g = load_golden("bases_kat.json")
    pts = oracle.fp_from_mont(oracle.synthetic_bases(5, g["a"], g["d"]).reshape(-1, 6)).reshape(5, 12)
    assert [[limbs_to_hex(p[:6], 48), limbs_to_hex(p[6:], 48)] for p in pts] == g["points"]


def test_msm_invariants(oracle):
    n = 200
    pts = oracle.synthetic_bases(n)
    s, t = oracle.fr_to_mont(oracle.random_fr(21, n)), oracle.fr_to_mont(oracle.random_fr(22, n))
    aff = lambda xyz: oracle.g1_proj_to_affine_canonical(xyz)
    a, b, ab = aff(oracle.msm_variable_base(pts, s)), aff(oracle.msm_variable_base(pts, t)), aff(
        oracle.msm_variable_base(pts, oracle.fr_add(s, t)))
    assert model.g1_add(a, b) == ab                                          # linearity in the scalars
    assert aff(oracle.msm_variable_base(pts, s)) == aff(oracle.msm_naive(pts, s))  # bucket method == definition
    g = np.repeat(oracle.g1_generator().reshape(1, 12), n, axis=0)
    ssum = sum(model.random_fr(21, n)) % model.R
    assert aff(oracle.msm_variable_base(g, s)) == model.g1_mul(model.G1_GEN, ssum)  # MSM(s,[G]*n) = (Σs)·G
