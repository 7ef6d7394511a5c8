"""Pure-Python big-int model of the hot path (TEST INFRASTRUCTURE — never shipped, never timed).

PARITY UNPINNED: /root/reference holds no MSM/NTT code, no tests and no golden vectors
(SURVEY.md §0.1, §8c).  The arithmetic lives in un-vendored crates pinned by
/root/reference/Cargo.toml:19-20 (dusk-plonk 0.8.2, dusk-bls12_381 0.8).  This file restates
the *published mathematical definitions* those crates implement, with plain Python ints and
no Montgomery form, so it is an independent second implementation against which the C oracle
(oracle/*.c, 64-bit-limb Montgomery) and the CUDA path are byte-compared.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.

Definitions restated
  * Fr / Fp moduli, generator 7, 2^32-th root of unity: SURVEY.md Appendix A (public BLS12-381
    constants; dusk-bls12_381 `Scalar`, Cargo.toml:20).
  * EvaluationDomain {fft, ifft, coset_fft, coset_ifft}: SURVEY.md Appendix B.2
    (dusk-plonk 0.8.2 `fft::EvaluationDomain`, Cargo.toml:19).
  * msm_variable_base: SURVEY.md Appendix B.1 (dusk-bls12_381 `multiscalar_mul`).
  * Encodings (`Scalar::to_bytes`, `G1Affine::to_bytes`): SURVEY.md Appendix A.4; the reference
    depends on the little-endian scalar order at /root/reference/src/zk/gadgets.rs:228-236.
"""

# ----------------------------------------------------------------------------- constants
P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
TWO_ADICITY = 32
GENERATOR = 7  # multiplicative generator of Fr and the coset shift
ROOT_OF_UNITY = pow(GENERATOR, (R - 1) >> TWO_ADICITY, R)
GX = 0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB
GY = 0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1
B_COEFF = 4
FR_MONT_R = (1 << 256) % R
FP_MONT_R = (1 << 384) % P
MASK64 = (1 << 64) - 1


# ----------------------------------------------------------------------------- PRNG (shared by C / CUDA / numpy)
def splitmix64_stream(seed):
    """SplitMix64 — the seeded generator every implementation in this repo shares."""
    x = seed & MASK64
    while True:
        x = (x + 0x9E3779B97F4A7C15) & MASK64
        z = x
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
        yield z ^ (z >> 31)


def random_fr(seed, n):
    """n values uniform in [0, r): candidates are 4 SplitMix64 outputs (LE limbs), top limb masked
    to 63 bits (255-bit candidate), rejected when >= r."""
    out = []
    g = splitmix64_stream(seed)
    while len(out) < n:
        l0, l1, l2, l3 = next(g), next(g), next(g), next(g)
        v = l0 | (l1 << 64) | (l2 << 128) | ((l3 & ((1 << 63) - 1)) << 192)
        if v < R:
            out.append(v)
    return out


# ----------------------------------------------------------------------------- limb / byte helpers
def to_limbs(v, n):
    return [(v >> (64 * i)) & MASK64 for i in range(n)]


def from_limbs(limbs):
    v = 0
    for i, l in enumerate(limbs):
        v |= int(l) << (64 * i)
    return v


def fr_to_mont(v):
    return (v * FR_MONT_R) % R


def fr_from_mont(v):
    return (v * pow(FR_MONT_R, -1, R)) % R


def fp_to_mont(v):
    return (v * FP_MONT_R) % P


def fp_from_mont(v):
    return (v * pow(FP_MONT_R, -1, P)) % P


def scalar_to_bytes(v):
    """`BlsScalar::to_bytes`: 32 bytes, little-endian, canonical (non-Montgomery)."""
    return int(v % R).to_bytes(32, "little")


def g1_compress(pt):
    """`G1Affine::to_bytes`: 48 bytes big-endian x, bit7=compressed, bit6=identity, bit5=y larger."""
    if pt is None:
        return bytes([0xC0]) + bytes(47)
    x, y = pt
    b = bytearray(x.to_bytes(48, "big"))
    b[0] |= 0x80
    if y > (P - 1) // 2:
        b[0] |= 0x20
    return bytes(b)


# ----------------------------------------------------------------------------- Fr / NTT
def domain(n_coeffs):
    """EvaluationDomain::new — SURVEY App. B.2."""
    size = 1
    log = 0
    while size < n_coeffs:
        size <<= 1
        log += 1
    if log >= TWO_ADICITY:
        raise ValueError("InvalidEvalDomainSize")
    gen = pow(ROOT_OF_UNITY, 1 << (TWO_ADICITY - log), R)
    return {
        "size": size,
        "log_size": log,
        "group_gen": gen,
        "group_gen_inv": pow(gen, -1, R),
        "size_inv": pow(size, -1, R),
        "generator_inv": pow(GENERATOR, -1, R),
    }


def naive_dft(a, omega):
    n = len(a)
    return [sum(a[j] * pow(omega, i * j, R) for j in range(n)) % R for i in range(n)]


def _bitrev(k, log):
    r = 0
    for _ in range(log):
        r = (r << 1) | (k & 1)
        k >>= 1
    return r


def serial_fft(a, omega, log):
    """Bit-reverse then log radix-2 DIT stages — SURVEY App. B.2 `serial_fft`."""
    n = 1 << log
    a = list(a)
    for k in range(n):
        rk = _bitrev(k, log)
        if k < rk:
            a[k], a[rk] = a[rk], a[k]
    m = 1
    for _ in range(log):
        w_m = pow(omega, n // (2 * m), R)
        for k in range(0, n, 2 * m):
            w = 1
            for j in range(m):
                t = a[k + j + m] * w % R
                a[k + j + m] = (a[k + j] - t) % R
                a[k + j] = (a[k + j] + t) % R
                w = w * w_m % R
        m *= 2
    return a


def _pad(a, size):
    return list(a) + [0] * (size - len(a))


def fft(a, d=None):
    d = d or domain(len(a))
    return serial_fft(_pad(a, d["size"]), d["group_gen"], d["log_size"])


def ifft(a, d=None):
    d = d or domain(len(a))
    out = serial_fft(_pad(a, d["size"]), d["group_gen_inv"], d["log_size"])
    return [x * d["size_inv"] % R for x in out]


def coset_fft(a, d=None):
    d = d or domain(len(a))
    a = _pad(a, d["size"])
    g = 1
    scaled = []
    for x in a:
        scaled.append(x * g % R)
        g = g * GENERATOR % R
    return fft(scaled, d)


def coset_ifft(a, d=None):
    d = d or domain(len(a))
    out = ifft(a, d)
    g = 1
    res = []
    for x in out:
        res.append(x * g % R)
        g = g * d["generator_inv"] % R
    return res


# ----------------------------------------------------------------------------- G1 (affine, None = identity)
def g1_on_curve(pt):
    if pt is None:
        return True
    x, y = pt
    return (y * y - x * x * x - B_COEFF) % P == 0


def g1_neg(pt):
    if pt is None:
        return None
    return (pt[0], (-pt[1]) % P)


def g1_add(a, b):
    if a is None:
        return b
    if b is None:
        return a
    x1, y1 = a
    x2, y2 = b
    if x1 == x2:
        if (y1 + y2) % P == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, P) % P
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, P) % P
    x3 = (lam * lam - x1 - x2) % P
    y3 = (lam * (x1 - x3) - y1) % P
    return (x3, y3)


def g1_mul(pt, k):
    k %= R
    acc = None
    add = pt
    while k:
        if k & 1:
            acc = g1_add(acc, add)
        add = g1_add(add, add)
        k >>= 1
    return acc


G1_GEN = (GX, GY)


def msm_naive(points, scalars):
    """Σ sᵢ·Pᵢ by independent double-and-add — the definition `msm_variable_base` computes."""
    acc = None
    for p, s in zip(points, scalars):
        acc = g1_add(acc, g1_mul(p, s))
    return acc


def _ln_without_floats(a):
    # ⌊log2(a) · 69/100⌋ — SURVEY App. B.1
    return (a.bit_length() - 1) * 69 // 100


def msm_variable_base(points, scalars):
    """Bucket method exactly as SURVEY App. B.1 restates dusk-bls12_381's msm_variable_base
    (unsigned c-bit windows, scalar==1 shortcut, running-sum reduction, Horner combine)."""
    n = len(points)
    c = 3 if n < 32 else _ln_without_floats(n) + 2
    window_sums = []
    for w_start in range(0, 255, c):
        res = None
        buckets = [None] * ((1 << c) - 1)
        for s, pt in zip(scalars, points):
            if s == 0:
                continue
            if s == 1:
                if w_start == 0:
                    res = g1_add(res, pt)
                continue
            d = (s >> w_start) & ((1 << c) - 1)
            if d:
                buckets[d - 1] = g1_add(buckets[d - 1], pt)
        running = None
        for b in reversed(buckets):
            running = g1_add(running, b)
            res = g1_add(res, running)
        window_sums.append(res)
    total = None
    for s in reversed(window_sums[1:]):
        total = g1_add(total, s)
        for _ in range(c):
            total = g1_add(total, total)
    return g1_add(total, window_sums[0])


def synthetic_bases(n, a=0xB2000001, d=0x9E3779B1):
    """P_i = (a + i·d)·G — the reproducible base set of SURVEY §8d (distinct, order-r subgroup)."""
    pts = []
    cur = g1_mul(G1_GEN, a)
    step = g1_mul(G1_GEN, d)
    for _ in range(n):
        pts.append(cur)
        cur = g1_add(cur, step)
    return pts
