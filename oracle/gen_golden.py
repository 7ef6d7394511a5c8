"""Generate tests/golden/*.json from the pure-Python big-int model (oracle/model.py).

The reference ships no fixtures, so these known-answer files are produced by the
independent big-int model — NOT by the C oracle and NOT by the CUDA path — and both of those are
checked against them.  Run from the repo root:  python oracle/gen_golden.py
"""
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import model as M  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def hx(v, nbytes):
    return format(v, "0%dx" % (2 * nbytes))


def digest_scalars(vals):
    h = hashlib.sha256()
    for v in vals:
        h.update(M.scalar_to_bytes(v))
    return h.hexdigest()


def ntt_kats():
    cases = []
    for log_n, seed in [(0, 1), (1, 2), (2, 3), (4, 4), (7, 5), (10, 6), (12, 7), (13, 8)]:
        n = 1 << log_n
        x = M.random_fr(0xF1F00000 + seed, n)
        d = M.domain(n)
        entry = {"log_n": log_n, "seed": 0xF1F00000 + seed, "input_sha256": digest_scalars(x)}
        for name, fn in (("fft", M.fft), ("ifft", M.ifft), ("coset_fft", M.coset_fft), ("coset_ifft", M.coset_ifft)):
            y = fn(x, d)
            entry[name + "_sha256"] = digest_scalars(y)
            entry[name + "_head"] = [hx(v, 32) for v in y[:4]]
        if log_n <= 4:
            entry["input"] = [hx(v, 32) for v in x]
            entry["fft"] = [hx(v, 32) for v in M.fft(x, d)]
        cases.append(entry)
    # structured vectors: delta -> all ones, all (r-1), zero-padding of a short input
    n = 64
    d = M.domain(n)
    delta = [1] + [0] * (n - 1)
    cases.append({"log_n": 6, "structured": "delta0", "fft_sha256": digest_scalars(M.fft(delta, d)),
                  "coset_fft_sha256": digest_scalars(M.coset_fft(delta, d))})
    allm1 = [M.R - 1] * n
    cases.append({"log_n": 6, "structured": "all_r_minus_1", "fft_sha256": digest_scalars(M.fft(allm1, d)),
                  "ifft_sha256": digest_scalars(M.ifft(allm1, d))})
    short = M.random_fr(0xF1F000AA, 37)
    cases.append({"log_n": 6, "structured": "short37_zero_padded", "seed": 0xF1F000AA,
                  "fft_sha256": digest_scalars(M.fft(short)), "coset_ifft_sha256": digest_scalars(M.coset_ifft(short))})
    return {"about": "EvaluationDomain KATs from oracle/model.py (big-int); scalars hex big-endian canonical; "
                     "sha256 over concatenated 32-byte little-endian canonical encodings (Scalar::to_bytes)",
            "domain_constants": {"root_of_unity_2_32": hx(M.ROOT_OF_UNITY, 32), "generator": 7,
                                 "generator_inv": hx(pow(7, -1, M.R), 32)},
            "cases": cases}


def msm_kats():
    cases = []

    def add(name, pts_desc, pts, scalars, extra=None):
        res = M.msm_naive(pts, scalars)
        assert M.g1_on_curve(res)
        e = {"name": name, "n": len(pts), "bases": pts_desc, "compressed": M.g1_compress(res).hex(),
             "x": None if res is None else hx(res[0], 48), "y": None if res is None else hx(res[1], 48)}
        if extra:
            e.update(extra)
        cases.append(e)

    for n, seed in [(1, 1), (2, 2), (7, 3), (31, 4), (32, 5), (100, 6), (300, 7)]:
        pts = M.synthetic_bases(n)
        s = M.random_fr(0xB2000000 + seed, n)
        add("random_%d" % n, "synthetic", pts, s, {"seed": 0xB2000000 + seed})
    n = 64
    pts = M.synthetic_bases(n)
    add("all_zero", "synthetic", pts, [0] * n, {"scalars": "0"})
    add("all_one", "synthetic", pts, [1] * n, {"scalars": "1"})
    add("all_r_minus_1", "synthetic", pts, [M.R - 1] * n, {"scalars": "r-1"})
    s8 = [v & 0xFF for v in M.random_fr(0xB20000F0, n)]
    add("eight_bit", "synthetic", pts, s8, {"seed": 0xB20000F0, "scalars": "low byte of random_fr"})
    half = [0 if i % 2 else v for i, v in enumerate(M.random_fr(0xB20000F1, n))]
    add("half_zero", "synthetic", pts, half, {"seed": 0xB20000F1, "scalars": "odd indices zeroed"})
    same = [pts[0]] * n
    add("all_equal_bases", "synthetic[0] repeated", same, M.random_fr(0xB20000F2, n), {"seed": 0xB20000F2})
    pm = []
    for i in range(n // 2):
        pm += [pts[i], M.g1_neg(pts[i])]
    add("plus_minus_pairs_equal_scalars", "P0,-P0,P1,-P1,..", pm,
        [v for v in M.random_fr(0xB20000F3, n // 2) for _ in (0, 1)], {"seed": 0xB20000F3})
    add("plus_minus_pairs", "P0,-P0,P1,-P1,..", pm, M.random_fr(0xB20000F4, n), {"seed": 0xB20000F4})
    # window-boundary scalars: 2^k and 2^k - 1 for every k
    wb = [(1 << k) % M.R for k in range(0, 255, 5)] + [((1 << k) - 1) % M.R for k in range(1, 255, 7)]
    add("powers_of_two", "synthetic", M.synthetic_bases(len(wb)), wb, {"scalars": "2^k (k=0,5,..) then 2^k-1 (k=1,8,..)"})
    return {"about": "msm_variable_base KATs from oracle/model.py: naive Σ sᵢ·Pᵢ over affine big-int arithmetic; "
                     "bases 'synthetic' = P_i=(a+i·d)·G with a=0xB2000001, d=0x9E3779B1; scalars = model.random_fr(seed,n) "
                     "(canonical values); result as zcash-compressed G1 (G1Affine::to_bytes) and affine x,y hex",
            "cases": cases}


def base_kats():
    pts = M.synthetic_bases(5)
    return {"about": "first synthetic bases, affine canonical hex", "a": 0xB2000001, "d": 0x9E3779B1,
            "points": [[hx(x, 48), hx(y, 48)] for x, y in pts],
            "generator_compressed": M.g1_compress(M.G1_GEN).hex()}


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    for name, fn in (("ntt_kat.json", ntt_kats), ("msm_kat.json", msm_kats), ("bases_kat.json", base_kats)):
        with open(os.path.join(OUT, name), "w") as f:
            json.dump(fn(), f, indent=1)
        print("wrote", name)
