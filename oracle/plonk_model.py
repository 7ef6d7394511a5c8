"""Pure-Python big-int model of the PLONK protocol above the hot path (TEST INFRASTRUCTURE — never
shipped, never timed): Merlin transcript, the arithmetic part of `StandardComposer`, preprocessing,
the five prover rounds, the KZG10 opening scheme, the BLS12-381 pairing and the verifier.

PARITY UNPINNED.  None of this code exists under /root/reference: the reference only *builds circuits*
(/root/reference/src/zk/gadgets.rs:28-225, circuits.rs:51-72) against dusk-plonk 0.8.2
(/root/reference/Cargo.toml:19), which is neither vendored nor buildable here.  The
protocol is restated from SURVEY.md Appendix B.3 and §3.2-3.6 [UPSTREAM-MEMORY]; every proof element is a
canonical field element or group element, so any implementation of the same equations, transcript labels
and encodings yields the same 1040 bytes.  External anchors used: the Merlin known-answer vector
(`test protocol`), public BLS12-381 constants, and the algebraic soundness of the whole loop — a proof made
by `prove` is accepted by the pairing-based `verify`, and rejected after any single-byte change.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from model import (GENERATOR, G1_GEN, P, R, coset_fft, coset_ifft, domain, fft, g1_add, g1_compress, g1_mul,
                   g1_neg, ifft, msm_naive, scalar_to_bytes)

K1, K2, K3 = 7, 13, 17  # coset representatives of the four wire columns


# ============================================================================= Merlin (STROBE-128 / Keccak-f[1600])
_RC = [0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000, 0x000000000000808B,
       0x0000000080000001, 0x8000000080008081, 0x8000000000008009, 0x000000000000008A, 0x0000000000000088,
       0x0000000080008009, 0x000000008000000A, 0x000000008000808B, 0x800000000000008B, 0x8000000000008089,
       0x8000000000008003, 0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
       0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008]
_ROT = [[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61], [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]]
_M64 = (1 << 64) - 1


def _rol(v, n):
    n %= 64
    return ((v << n) | (v >> (64 - n))) & _M64 if n else v


def keccak_f1600(state):
    """state: bytearray(200), permuted in place (FIPS 202, 24 rounds)."""
    a = [[int.from_bytes(state[8 * (x + 5 * y):8 * (x + 5 * y) + 8], "little") for y in range(5)] for x in range(5)]
    for rc in _RC:
        c = [a[x][0] ^ a[x][1] ^ a[x][2] ^ a[x][3] ^ a[x][4] for x in range(5)]
        d = [c[(x - 1) % 5] ^ _rol(c[(x + 1) % 5], 1) for x in range(5)]
        a = [[a[x][y] ^ d[x] for y in range(5)] for x in range(5)]
        b = [[0] * 5 for _ in range(5)]
        for x in range(5):
            for y in range(5):
                b[y][(2 * x + 3 * y) % 5] = _rol(a[x][y], _ROT[x][y])
        a = [[b[x][y] ^ ((~b[(x + 1) % 5][y]) & b[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
        a[0][0] ^= rc
    for x in range(5):
        for y in range(5):
            state[8 * (x + 5 * y):8 * (x + 5 * y) + 8] = a[x][y].to_bytes(8, "little")


class Strobe128:
    R_ = 166
    FLAG_I, FLAG_A, FLAG_C, FLAG_T, FLAG_M, FLAG_K = 1, 2, 4, 8, 16, 32

    def __init__(self, protocol_label):
        st = bytearray(200)
        st[0:6] = bytes([1, self.R_ + 2, 1, 0, 1, 96])
        st[6:18] = b"STROBEv1.0.2"
        keccak_f1600(st)
        self.state, self.pos, self.pos_begin, self.cur_flags = st, 0, 0, 0
        self.meta_ad(protocol_label, False)

    def _run_f(self):
        self.state[self.pos] ^= self.pos_begin
        self.state[self.pos + 1] ^= 0x04
        self.state[self.R_ + 1] ^= 0x80
        keccak_f1600(self.state)
        self.pos = self.pos_begin = 0

    def _absorb(self, data):
        for byte in data:
            self.state[self.pos] ^= byte
            self.pos += 1
            if self.pos == self.R_:
                self._run_f()

    def _squeeze(self, n):
        out = bytearray()
        for _ in range(n):
            out.append(self.state[self.pos])
            self.state[self.pos] = 0
            self.pos += 1
            if self.pos == self.R_:
                self._run_f()
        return bytes(out)

    def _begin_op(self, flags, more):
        if more:
            assert self.cur_flags == flags
            return
        assert not flags & self.FLAG_T
        old_begin = self.pos_begin
        self.pos_begin = self.pos + 1
        self.cur_flags = flags
        self._absorb(bytes([old_begin, flags]))
        if flags & (self.FLAG_C | self.FLAG_K) and self.pos != 0:
            self._run_f()

    def meta_ad(self, data, more):
        self._begin_op(self.FLAG_M | self.FLAG_A, more)
        self._absorb(data)

    def ad(self, data, more):
        self._begin_op(self.FLAG_A, more)
        self._absorb(data)

    def prf(self, n, more=False):
        self._begin_op(self.FLAG_I | self.FLAG_A | self.FLAG_C, more)
        return self._squeeze(n)


class Transcript:
    """merlin::Transcript plus dusk-plonk's TranscriptProtocol helpers."""

    def __init__(self, label):
        self.strobe = Strobe128(b"Merlin v1.0")
        self.append_message(b"dom-sep", label)

    def append_message(self, label, message):
        self.strobe.meta_ad(label, False)
        self.strobe.meta_ad(len(message).to_bytes(4, "little"), True)
        self.strobe.ad(message, False)

    def append_u64(self, label, x):
        self.append_message(label, int(x).to_bytes(8, "little"))

    def challenge_bytes(self, label, n):
        self.strobe.meta_ad(label, False)
        self.strobe.meta_ad(n.to_bytes(4, "little"), True)
        return self.strobe.prf(n)

    # dusk-plonk TranscriptProtocol
    def append_commitment(self, label, pt):
        self.append_message(label, g1_compress(pt))

    def append_scalar(self, label, s):
        self.append_message(label, scalar_to_bytes(s))

    def challenge_scalar(self, label):
        return int.from_bytes(self.challenge_bytes(label, 64), "little") % R  # BlsScalar::from_bytes_wide

    def circuit_domain_sep(self, n):
        self.append_message(b"dom-sep", b"circuit_size")
        self.append_u64(b"n", n)

    def clone(self):
        import copy
        return copy.deepcopy(self)


# ============================================================================= composer (arithmetic + range selectors)
SELECTORS = ["q_m", "q_l", "q_r", "q_o", "q_c", "q_4", "q_arith", "q_range", "q_logic", "q_fixed_group_add",
             "q_variable_group_add"]


class Composer:
    """The part of dusk-plonk's StandardComposer the reference's gadgets drive (call sites
    /root/reference/src/zk/gadgets.rs:60,70,81,132,165,206,211,218 and circuits.rs:71).  Gate equation
    q_arith·(q_m·a·b + q_l·a + q_r·b + q_o·c + q_4·d + q_c) + PI = 0."""

    def __init__(self, with_dummy=True):
        self.q = {k: [] for k in SELECTORS}
        self.w = [[], [], [], []]       # variable index per gate: w_l, w_r, w_o, w_4
        self.values = []                # variable → value
        self.pi = {}                    # gate index → public input value
        self.n = 0
        self.zero_var = 0               # Variable(0): the first allocation below is the constrained zero itself
        self.zero_var = self.add_witness_to_circuit_description(0)
        if with_dummy:
            self.add_dummy_constraints()

    def add_input(self, v):
        self.values.append(v % R)
        return len(self.values) - 1

    def poly_gate(self, a, b, c, d, q_m=0, q_l=0, q_r=0, q_o=0, q_c=0, q_4=0, pi=0, q_arith=1, q_range=0, q_logic=0, q_fixed=0,
                  q_var=0):
        for k, v in (("q_m", q_m), ("q_l", q_l), ("q_r", q_r), ("q_o", q_o), ("q_c", q_c), ("q_4", q_4),
                     ("q_arith", q_arith), ("q_range", q_range), ("q_logic", q_logic), ("q_fixed_group_add", q_fixed),
                     ("q_variable_group_add", q_var)):
            self.q[k].append(v % R)
        for col, var in zip(self.w, (a, b, c, d)):
            col.append(var)
        if pi % R:
            self.pi[self.n] = pi % R
        self.n += 1

    def add(self, ql_a, qr_b, q_c, pi):
        (q_l, a), (q_r, b) = ql_a, qr_b
        c = self.add_input(q_l * self.values[a] + q_r * self.values[b] + q_c + pi)
        self.poly_gate(a, b, c, self.zero_var, q_l=q_l, q_r=q_r, q_o=-1, q_c=q_c, pi=pi)
        return c

    def mul(self, q_m, a, b, q_c, pi):
        c = self.add_input(q_m * self.values[a] * self.values[b] + q_c + pi)
        self.poly_gate(a, b, c, self.zero_var, q_m=q_m, q_o=-1, q_c=q_c, pi=pi)
        return c

    def mul_gate(self, a, b, c, q_m, q_o, q_c, pi):
        self.poly_gate(a, b, c, self.zero_var, q_m=q_m, q_o=q_o, q_c=q_c, pi=pi)

    def boolean_gate(self, a):
        self.poly_gate(a, a, a, self.zero_var, q_m=1, q_o=-1)

    def constrain_to_constant(self, a, k, pi):
        self.poly_gate(a, a, a, self.zero_var, q_l=1, q_c=-k, pi=pi)

    def add_witness_to_circuit_description(self, v):
        var = self.add_input(v)
        self.constrain_to_constant(var, v, 0)
        return var

    def assert_equal(self, a, b):
        self.poly_gate(a, b, self.zero_var, self.zero_var, q_l=1, q_r=-1)

    # ---- ECC gadgets on JubJub (dusk-plonk constraint_system::ecc; call sites /root/reference/src/zk/gadgets.rs:34,37,40,
    # circuits.rs:64-65).  A Point is a pair of variables (x, y).
    def fixed_base_scalar_mul(self, scalar_var, generator):
        """256 rows of the fixed-base widget (2-bit windowed NAF ladder from the most significant digit) plus one plain row
        carrying the final accumulators; the scalar accumulator is constrained to equal `scalar_var`."""
        num_bits = 256
        multiples = [generator]
        for _ in range(num_bits - 1):
            multiples.append(jj_add(multiples[-1], multiples[-1]))
        multiples.reverse()                                           # multiples[i] pairs with the i-th digit from the top
        k = self.values[scalar_var]
        assert k < JJ_ORDER, "JubJubScalar::from_bytes(..).unwrap(): the scalar must be below the JubJub group order"
        naf = wnaf2(k)                                               # 256 digits in {−1, 0, 1}, least significant first
        scalar_acc, point_acc, xy_alphas = [0], [(0, 1)], []
        for i, entry in enumerate(reversed(naf)):
            bx, by = multiples[i]
            to_add = (0, 1) if entry == 0 else ((bx, by) if entry == 1 else ((-bx) % R, by))
            scalar_acc.append((2 * scalar_acc[i] + entry) % R)
            point_acc.append(jj_add(point_acc[i], to_add))
            xy_alphas.append(to_add[0] * to_add[1] % R)
        for i in range(num_bits):
            acc_x, acc_y = self.add_input(point_acc[i][0]), self.add_input(point_acc[i][1])
            acc_bit = self.add_input(scalar_acc[i])
            if i == 0:
                self.constrain_to_constant(acc_x, 0, 0)
                self.constrain_to_constant(acc_y, 1, 0)
                self.constrain_to_constant(acc_bit, 0, 0)
            bx, by = multiples[i]
            xy_alpha = self.add_input(xy_alphas[i])
            self.poly_gate(acc_x, acc_y, xy_alpha, acc_bit, q_l=bx, q_r=by, q_c=bx * by, q_arith=0, q_fixed=1)
        acc_x, acc_y = self.add_input(point_acc[num_bits][0]), self.add_input(point_acc[num_bits][1])
        last_bit = self.add_input(scalar_acc[num_bits])
        self.poly_gate(acc_x, acc_y, self.zero_var, last_bit)        # big_add_gate with zero selectors: the "next" row of gate 255
        self.assert_equal(last_bit, scalar_var)
        return (acc_x, acc_y)

    def point_addition_gate(self, p1, p2):
        (x1, y1), (x2, y2) = p1, p2
        v = self.values
        x3v, y3v = jj_add((v[x1], v[y1]), (v[x2], v[y2]))
        x1y2 = self.add_input(v[x1] * v[y2])
        x3, y3 = self.add_input(x3v), self.add_input(y3v)
        self.poly_gate(x1, y1, x2, y2, q_arith=0, q_var=1)
        self.poly_gate(x3, y3, self.zero_var, x1y2, q_arith=0)
        return (x3, y3)

    def assert_equal_public_point(self, point, affine):
        self.constrain_to_constant(point[0], 0, -affine[0])
        self.constrain_to_constant(point[1], 0, -affine[1])

    # ---- logic widget: XOR / AND of two num_bits-bit values, one 2-bit quad per row from the top (constraint_system::logic)
    def logic_gate(self, a, b, num_bits, is_xor):
        assert num_bits % 2 == 0
        av, bv = self.values[a], self.values[b]
        nq = num_bits // 2
        sel = -1 if is_xor else 1
        rows_l, rows_r, rows_4, rows_o = [self.zero_var], [self.zero_var], [self.zero_var], []
        la = ra = oa = 0
        for i in range(nq):
            sh = 2 * (nq - 1 - i)
            lq, rq = (av >> sh) & 3, (bv >> sh) & 3
            oq = (lq ^ rq) if is_xor else (lq & rq)
            la, ra, oa = 4 * la + lq, 4 * ra + rq, 4 * oa + oq
            rows_l.append(self.add_input(la))
            rows_r.append(self.add_input(ra))
            rows_4.append(self.add_input(oa))
            rows_o.append(self.add_input(lq * rq))
        rows_o.append(self.zero_var)
        for i in range(nq + 1):
            last = i == nq
            self.poly_gate(rows_l[i], rows_r[i], rows_o[i], rows_4[i], q_arith=0, q_c=0 if last else sel, q_logic=0 if last else sel)
        self.assert_equal(a, rows_l[nq])
        self.assert_equal(b, rows_r[nq])
        return rows_4[nq]

    def xor_gate(self, a, b, num_bits):
        return self.logic_gate(a, b, num_bits, True)

    def and_gate(self, a, b, num_bits):
        return self.logic_gate(a, b, num_bits, False)

    def add_dummy_constraints(self):
        six, one, seven, m20 = (self.add_input(v) for v in (6, 1, 7, -20))
        self.poly_gate(six, seven, m20, one, q_m=1, q_l=2, q_r=3, q_o=4, q_c=4, q_4=1)
        self.poly_gate(m20, six, seven, self.zero_var, q_m=1, q_l=1, q_r=1, q_o=1, q_c=127)

    def check(self):
        """Every gate satisfied by the current assignment (host-side sanity, like upstream's check_circuit_satisfied)."""
        v = self.values
        for i in range(self.n):
            a, b, c, d = (v[col[i]] for col in self.w)
            q = {k: self.q[k][i] for k in SELECTORS}
            t = q["q_arith"] * (q["q_m"] * a * b + q["q_l"] * a + q["q_r"] * b + q["q_o"] * c + q["q_4"] * d + q["q_c"])
            if (t + self.pi.get(i, 0)) % R:
                return False
            j = (i + 1) % self.n
            an, bn, dn = v[self.w[0][j]], v[self.w[1][j]], v[self.w[3][j]]
            # every separation challenge must make the widget vanish: test with two unrelated values
            for sep in (3, 0x1234567):
                if q["q_range"] and _range_term(a, b, c, d, dn, sep):
                    return False
                if q["q_logic"] and _logic_term(a, an, b, bn, c, d, dn, q["q_c"], sep):
                    return False
                if q["q_fixed_group_add"] and _fixed_base_term(a, an, b, bn, c, d, dn, q["q_l"], q["q_r"], q["q_c"], sep):
                    return False
                if q["q_variable_group_add"] and _var_base_term(a, an, b, bn, c, d, dn, sep):
                    return False
        return True


def sigma_positions(wires, n_pad):
    """compute_sigma_permutations: position (col, i) ↦ next position of the same variable (in order of
    occurrence: gate order, then l, r, o, 4), identity on padding.  Returns 4 lists of (col, idx)."""
    n = len(wires[0])
    occ = {}
    for i in range(n):
        for col in range(4):
            occ.setdefault(wires[col][i], []).append((col, i))
    sig = [[(col, i) for i in range(n_pad)] for col in range(4)]
    for lst in occ.values():
        for k, (col, i) in enumerate(lst):
            sig[col][i] = lst[(k + 1) % len(lst)]
    return sig


# ============================================================================= KZG10
def srs_setup(tau, n_points):
    """PublicParameters::setup with a known trapdoor: powers_of_g[i] = τ^i·G (G2 side: H, τ·H)."""
    pts, t = [], 1
    for _ in range(n_points):
        pts.append(g1_mul(G1_GEN, t))
        t = t * tau % R
    return pts


def commit(ck, coeffs):
    assert len(coeffs) <= len(ck), "PolynomialDegreeTooLarge"
    return msm_naive(ck[:len(coeffs)], coeffs)


def poly_eval(coeffs, z):
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * z + c) % R
    return acc


def ruffini(coeffs, z):
    """(p(X) − p(z)) / (X − z): Polynomial::ruffini."""
    q = [0] * max(len(coeffs) - 1, 0)
    acc = 0
    for i in range(len(coeffs) - 1, 0, -1):
        acc = (coeffs[i] + acc * z) % R
        q[i - 1] = acc
    return q


def compute_aggregate_witness(polys, point, transcript):
    v = transcript.challenge_scalar(b"aggregate_witness")
    n = max(len(p) for p in polys)
    num = [0] * n
    pw = 1
    for p in polys:
        for j, c in enumerate(p):
            num[j] = (num[j] + c * pw) % R
        pw = pw * v % R
    return ruffini(num, point)


# ============================================================================= preprocessing
def _pad(a, n):
    return list(a) + [0] * (n - len(a))


def preprocess(comp, ck, label):
    """StandardComposer::preprocess_prover + VerifierKey::seed_transcript."""
    d = domain(comp.n)
    n = d["size"]
    d4 = domain(4 * n)
    pk = {"n": n, "domain": d, "domain4": d4, "circuit_size": comp.n}
    pk["q_evals"] = {k: _pad(comp.q[k], n) for k in SELECTORS}
    pk["q_poly"] = {k: ifft(pk["q_evals"][k], d) for k in SELECTORS}
    sig = sigma_positions(comp.w, n)
    ks = [1, K1, K2, K3]
    w = d["group_gen"]
    roots = [pow(w, i, R) for i in range(n)]
    pk["sigma_evals"] = [[ks[c] * roots[i] % R for (c, i) in sig[col]] for col in range(4)]
    pk["sigma_poly"] = [ifft(e, d) for e in pk["sigma_evals"]]
    vk = {"n": n}
    vk["q"] = {k: commit(ck, pk["q_poly"][k]) for k in SELECTORS}
    vk["sigma"] = [commit(ck, p) for p in pk["sigma_poly"]]
    pk["q_4n"] = {k: coset_fft(pk["q_poly"][k], d4) for k in SELECTORS}
    pk["sigma_4n"] = [coset_fft(p, d4) for p in pk["sigma_poly"]]
    pk["linear_4n"] = coset_fft([0, 1], d4)
    # X^n − 1 on the coset 7·H_4n
    g4 = d4["group_gen"]
    pk["vh_4n"] = [(pow(GENERATOR * pow(g4, i, R), n, R) - 1) % R for i in range(4 * n)]
    t = Transcript(label)
    seed_transcript(t, vk)
    return pk, vk, t


def seed_transcript(t, vk):
    for lab, key in ((b"q_m", "q_m"), (b"q_l", "q_l"), (b"q_r", "q_r"), (b"q_o", "q_o"), (b"q_c", "q_c"),
                     (b"q_4", "q_4"), (b"q_arith", "q_arith"), (b"q_range", "q_range"), (b"q_logic", "q_logic"),
                     (b"q_variable_group_add", "q_variable_group_add"), (b"q_fixed_group_add", "q_fixed_group_add")):
        t.append_commitment(lab, vk["q"][key])
    for lab, c in zip((b"left_sigma", b"right_sigma", b"out_sigma", b"fourth_sigma"), vk["sigma"]):
        t.append_commitment(lab, c)
    t.circuit_domain_sep(vk["n"])


# ============================================================================= widgets
def _delta(f):
    return f * (f - 1) * (f - 2) * (f - 3) % R


def _range_term(a, b, c, d, d_next, sep):
    kappa = sep * sep % R
    return (_delta(c - 4 * d) + kappa * _delta(b - 4 * c) + kappa * kappa * _delta(a - 4 * b)
            + kappa * kappa * kappa * _delta(d_next - 4 * a)) * sep % R


# JubJub (dusk-jubjub 0.10, /root/reference/Cargo.toml:21): −x² + y² = 1 + d·x²·y² over Fr, d = −10240/10241.  The two
# generators the reference's gadgets use (gadgets.rs:34,37): GENERATOR has y = 18 (the x below is the root of the curve
# equation with that y, prime-order subgroup); GENERATOR_NUMS as published.  Both are checked on-curve and of order
# JJ_ORDER in tests/test_prover_cpu.py — the constants are pinned by the curve, not by memory.
EDWARDS_D = (-10240 * pow(10241, -1, R)) % R
JJ_ORDER = 0x0E7DB4EA6533AFA906673B0101343B00A6682093CCC81082D0970E5ED6F72CB7
JJ_GENERATOR = (0x3FD2814C43AC65A6F1FBF02D0FD6CCE62E3EBB21FD6C54ED4DF7B7FFEC7BEACA, 0x12)
JJ_GENERATOR_NUMS = (0x5E67B8F316F414F7BD9514C773FD4456931E316A39FE4541921710179DF76377,
                     0x43D80EB3B2F3EB1B7B162DBEEB3B34FD9949BA0F82A5507A6705B707162E3EF8)


def jj_add(p, q):
    (x1, y1), (x2, y2) = p, q
    t = EDWARDS_D * x1 % R * x2 % R * y1 % R * y2 % R
    return ((x1 * y2 + y1 * x2) * pow(1 + t, -1, R) % R, (y1 * y2 + x1 * x2) * pow(1 - t, -1, R) % R)


def jj_mul(p, k):
    acc = (0, 1)
    while k:
        if k & 1:
            acc = jj_add(acc, p)
        p = jj_add(p, p)
        k >>= 1
    return acc


def wnaf2(k):
    """Fr::compute_windowed_naf(2): 256 digits in {−1, 0, 1}, least significant first (non-adjacent form)."""
    out = [0] * 256
    i = 0
    while k >= 1:
        if k & 1:
            d = 2 - (k & 3)            # k mod 4 = 1 → 1, = 3 → −1
            out[i] = d
            k -= d
        k >>= 1
        i += 1
    return out


def _logic_term(a, a_next, b, b_next, c, d, d_next, q_c, sep):
    """logic widget: quads a' = a_next − 4a, b' = b_next − 4b, out d' = d_next − 4d, product w = c."""
    kappa = sep * sep % R
    k2, k3, k4 = kappa * kappa % R, pow(kappa, 3, R), pow(kappa, 4, R)
    qa, qb, qd, w = (a_next - 4 * a) % R, (b_next - 4 * b) % R, (d_next - 4 * d) % R, c
    f = w * (w * (4 * w - 18 * (qa + qb) + 81) + 18 * (qa * qa + qb * qb) - 81 * (qa + qb) + 83)
    e = 3 * (qa + qb + qd) - 2 * f
    bb = q_c * (9 * qd - 3 * (qa + qb))
    return (_delta(qa) + _delta(qb) * kappa + _delta(qd) * k2 + (w - qa * qb) * k3 + (bb + e) * k4) % R * sep % R


def _fixed_base_term(a, a_next, b, b_next, c, d, d_next, q_l, q_r, q_c, sep):
    """fixed-base widget: acc (a, b) + bit·(x_β, y_β) = (a_next, b_next) on JubJub, bit = d_next − 2d ∈ {−1, 0, 1}, c = x_α·y_α."""
    kappa = sep * sep % R
    k2, k3 = kappa * kappa % R, pow(kappa, 3, R)
    bit = (d_next - 2 * d) % R
    bit_cons = bit * (bit - 1) * (bit + 1)
    y_alpha = bit * bit * (q_r - 1) + 1
    x_alpha = bit * q_l
    xy_cons = (bit * q_c - c) * kappa
    t = c * a % R * b % R * EDWARDS_D % R
    x_cons = (a_next + a_next * t - (a * y_alpha + b * x_alpha)) * k2
    y_cons = (b_next - b_next * t - (b * y_alpha + a * x_alpha)) * k3
    return (bit_cons + x_cons + y_cons + xy_cons) % R * sep % R


def _var_base_term(a, a_next, b, b_next, c, d, d_next, sep):
    """variable-base widget: (x1, y1) = (a, b), (x2, y2) = (c, d), (x3, y3) = (a_next, b_next), x1·y2 = d_next."""
    kappa = sep * sep % R
    x1, y1, x2, y2, x3, y3, x1y2 = a, b, c, d, a_next, b_next, d_next
    y1x2, y1y2, x1x2 = y1 * x2 % R, y1 * y2 % R, x1 * x2 % R
    xy_cons = x1 * y2 - x1y2
    t = EDWARDS_D * x1y2 % R * y1x2 % R
    x3_cons = (x1y2 + y1x2 - (x3 + x3 * t)) * kappa
    y3_cons = (y1y2 + x1x2 - (y3 - y3 * t)) * kappa * kappa
    return (xy_cons + x3_cons + y3_cons) % R * sep % R


# ============================================================================= prover
def prove(comp, pk, ck, transcript):
    """Prover::prove_with_preprocessed.  Returns (proof dict, 1040 proof bytes)."""
    t = transcript.clone()
    d, d4, n = pk["domain"], pk["domain4"], pk["n"]
    omega = d["group_gen"]
    vals = comp.values
    # round 1
    w_evals = [_pad([vals[v] for v in col], n) for col in comp.w]
    w_poly = [ifft(e, d) for e in w_evals]
    w_comm = [commit(ck, p) for p in w_poly]
    for lab, c in zip((b"w_l", b"w_r", b"w_o", b"w_4"), w_comm):
        t.append_commitment(lab, c)
    # round 2
    beta = t.challenge_scalar(b"beta")
    t.append_scalar(b"beta", beta)
    gamma = t.challenge_scalar(b"gamma")
    ks = [1, K1, K2, K3]
    z_evals, state, root = [1], 1, 1
    for i in range(n - 1):
        num = den = 1
        for col in range(4):
            num = num * (w_evals[col][i] + beta * ks[col] * root + gamma) % R
            den = den * (w_evals[col][i] + beta * pk["sigma_evals"][col][i] + gamma) % R
        state = state * num * pow(den, -1, R) % R
        z_evals.append(state)
        root = root * omega % R
    z_poly = ifft(z_evals, d)
    z_comm = commit(ck, z_poly)
    t.append_commitment(b"z", z_comm)
    # round 3
    alpha = t.challenge_scalar(b"alpha")
    range_sep = t.challenge_scalar(b"range separation challenge")
    logic_sep = t.challenge_scalar(b"logic separation challenge")
    fixed_sep = t.challenge_scalar(b"fixed base separation challenge")
    var_sep = t.challenge_scalar(b"variable base separation challenge")
    pi_dense = [comp.pi.get(i, 0) for i in range(n)]
    pi_poly = ifft(pi_dense, d)
    N4 = 4 * n
    z4 = coset_fft(z_poly, d4)
    w4 = [coset_fft(p, d4) for p in w_poly]
    pi4 = coset_fft(pi_poly, d4)
    alpha2 = alpha * alpha % R
    l1_4 = coset_fft([alpha2 * d["size_inv"] % R] * n, d4)  # α²·L₁ on the coset
    q4, s4, x4 = pk["q_4n"], pk["sigma_4n"], pk["linear_4n"]
    quot = []
    for i in range(N4):
        a, b, c, dd = (w4[col][i] for col in range(4))
        d_next = w4[3][(i + 4) % N4]
        gate = q4["q_arith"][i] * (q4["q_m"][i] * a * b + q4["q_l"][i] * a + q4["q_r"][i] * b + q4["q_o"][i] * c
                                   + q4["q_4"][i] * dd + q4["q_c"][i])
        gate += q4["q_range"][i] * _range_term(a, b, c, dd, d_next, range_sep)
        a_next, b_next = w4[0][(i + 4) % N4], w4[1][(i + 4) % N4]
        if q4["q_logic"][i]:
            gate += q4["q_logic"][i] * _logic_term(a, a_next, b, b_next, c, dd, d_next, q4["q_c"][i], logic_sep)
        if q4["q_fixed_group_add"][i]:
            gate += q4["q_fixed_group_add"][i] * _fixed_base_term(a, a_next, b, b_next, c, dd, d_next, q4["q_l"][i], q4["q_r"][i],
                                                                  q4["q_c"][i], fixed_sep)
        if q4["q_variable_group_add"][i]:
            gate += q4["q_variable_group_add"][i] * _var_base_term(a, a_next, b, b_next, c, dd, d_next, var_sep)
        gate += pi4[i]
        x = x4[i]
        ident = ((a + beta * x + gamma) * (b + beta * K1 * x + gamma) % R) * ((c + beta * K2 * x + gamma)
                                                                               * (dd + beta * K3 * x + gamma) % R) % R * z4[i] * alpha
        copy = ((a + beta * s4[0][i] + gamma) * (b + beta * s4[1][i] + gamma) % R) * ((c + beta * s4[2][i] + gamma)
                                                                                       * (dd + beta * s4[3][i] + gamma) % R) % R * z4[(i + 4) % N4] * alpha
        one = (z4[i] - 1) * l1_4[i]
        quot.append((gate + ident - copy + one) * pow(pk["vh_4n"][i], -1, R) % R)
    t_poly = coset_ifft(quot, d4)
    t_parts = [t_poly[k * n:(k + 1) * n] for k in range(4)]
    t_comm = [commit(ck, p) for p in t_parts]
    for lab, c in zip((b"t_1", b"t_2", b"t_3", b"t_4"), t_comm):
        t.append_commitment(lab, c)
    # round 4
    z = t.challenge_scalar(b"z")
    zw = z * omega % R
    ev = {}
    ev["a_eval"], ev["b_eval"], ev["c_eval"], ev["d_eval"] = (poly_eval(p, z) for p in w_poly)
    ev["a_next_eval"], ev["b_next_eval"], ev["d_next_eval"] = (poly_eval(w_poly[k], zw) for k in (0, 1, 3))
    ev["left_sigma_eval"], ev["right_sigma_eval"], ev["out_sigma_eval"] = (poly_eval(pk["sigma_poly"][k], z) for k in range(3))
    ev["q_arith_eval"] = poly_eval(pk["q_poly"]["q_arith"], z)
    ev["q_c_eval"] = poly_eval(pk["q_poly"]["q_c"], z)
    ev["q_l_eval"] = poly_eval(pk["q_poly"]["q_l"], z)
    ev["q_r_eval"] = poly_eval(pk["q_poly"]["q_r"], z)
    ev["perm_eval"] = poly_eval(z_poly, zw)
    t_eval = poly_eval(t_poly, z)
    lin = linearisation_terms(ev, alpha, beta, gamma, (range_sep, logic_sep, fixed_sep, var_sep), z, d)
    lin_poly = [0] * n
    for name, coef in lin["q"].items():
        for j, c in enumerate(pk["q_poly"][name]):
            lin_poly[j] = (lin_poly[j] + c * coef) % R
    for j in range(n):
        lin_poly[j] = (lin_poly[j] + z_poly[j] * lin["z"] + pk["sigma_poly"][3][j] * lin["sigma4"]) % R
    ev["lin_poly_eval"] = poly_eval(lin_poly, z)
    for lab, key in EVAL_TRANSCRIPT_ORDER:
        t.append_scalar(lab, t_eval if key == "t_eval" else ev[key])
    # round 5
    zn = pow(z, n, R)
    quot_open = [(t_parts[0][j] + zn * t_parts[1][j] + zn * zn % R * t_parts[2][j] + pow(zn, 3, R) * t_parts[3][j]) % R
                 for j in range(n)]
    agg = compute_aggregate_witness([quot_open, lin_poly, w_poly[0], w_poly[1], w_poly[2], w_poly[3],
                                     pk["sigma_poly"][0], pk["sigma_poly"][1], pk["sigma_poly"][2]], z, t)
    w_z_comm = commit(ck, agg)
    shifted = compute_aggregate_witness([z_poly, w_poly[0], w_poly[1], w_poly[3]], zw, t)
    w_zw_comm = commit(ck, shifted)
    proof = {"a_comm": w_comm[0], "b_comm": w_comm[1], "c_comm": w_comm[2], "d_comm": w_comm[3], "z_comm": z_comm,
             "t_comm": t_comm, "w_z_comm": w_z_comm, "w_zw_comm": w_zw_comm, "evals": ev}
    return proof, proof_to_bytes(proof)


EVAL_TRANSCRIPT_ORDER = [(b"a_eval", "a_eval"), (b"b_eval", "b_eval"), (b"c_eval", "c_eval"), (b"d_eval", "d_eval"),
                         (b"a_next_eval", "a_next_eval"), (b"b_next_eval", "b_next_eval"), (b"d_next_eval", "d_next_eval"),
                         (b"left_sig_eval", "left_sigma_eval"), (b"right_sig_eval", "right_sigma_eval"),
                         (b"out_sig_eval", "out_sigma_eval"), (b"q_arith_eval", "q_arith_eval"), (b"q_c_eval", "q_c_eval"),
                         (b"q_l_eval", "q_l_eval"), (b"q_r_eval", "q_r_eval"), (b"perm_eval", "perm_eval"),
                         (b"t_eval", "t_eval"), (b"r_eval", "lin_poly_eval")]
EVAL_BYTES_ORDER = ["a_eval", "b_eval", "c_eval", "d_eval", "a_next_eval", "b_next_eval", "d_next_eval", "q_arith_eval",
                    "q_c_eval", "q_l_eval", "q_r_eval", "left_sigma_eval", "right_sigma_eval", "out_sigma_eval",
                    "lin_poly_eval", "perm_eval"]


def linearisation_terms(ev, alpha, beta, gamma, seps, z, d):
    """Scalar coefficient of every polynomial (prover) / commitment (verifier) in the linearisation
    r(X) — shared by `prove` and `verify` exactly as dusk-plonk's widgets share their formulas.
    seps = (range, logic, fixed base, variable base) separation challenges."""
    range_sep, logic_sep, fixed_sep, var_sep = seps
    a, b, c, dd = ev["a_eval"], ev["b_eval"], ev["c_eval"], ev["d_eval"]
    an, bn, dn = ev["a_next_eval"], ev["b_next_eval"], ev["d_next_eval"]
    qa = ev["q_arith_eval"]
    q = {"q_m": a * b % R * qa % R, "q_l": a * qa % R, "q_r": b * qa % R, "q_o": c * qa % R, "q_4": dd * qa % R,
         "q_c": qa, "q_range": _range_term(a, b, c, dd, dn, range_sep),
         "q_logic": _logic_term(a, an, b, bn, c, dd, dn, ev["q_c_eval"], logic_sep),
         "q_fixed_group_add": _fixed_base_term(a, an, b, bn, c, dd, dn, ev["q_l_eval"], ev["q_r_eval"], ev["q_c_eval"], fixed_sep),
         "q_variable_group_add": _var_base_term(a, an, b, bn, c, dd, dn, var_sep)}
    n = d["size"]
    z_h = (pow(z, n, R) - 1) % R
    l1 = z_h * pow(n * (z - 1) % R, -1, R) % R
    ident = (a + beta * z + gamma) * (b + beta * K1 * z + gamma) % R * (c + beta * K2 * z + gamma) % R * (dd + beta * K3 * z + gamma) % R
    zc = (ident * alpha + l1 * alpha * alpha) % R
    s4 = -((a + beta * ev["left_sigma_eval"] + gamma) * (b + beta * ev["right_sigma_eval"] + gamma) % R
           * (c + beta * ev["out_sigma_eval"] + gamma) % R * beta % R * ev["perm_eval"] % R * alpha) % R
    return {"q": q, "z": zc, "sigma4": s4, "z_h": z_h, "l1": l1}


def proof_to_bytes(pr):
    out = b"".join(g1_compress(pr[k]) for k in ("a_comm", "b_comm", "c_comm", "d_comm", "z_comm"))
    out += b"".join(g1_compress(c) for c in pr["t_comm"])
    out += g1_compress(pr["w_z_comm"]) + g1_compress(pr["w_zw_comm"])
    out += b"".join(scalar_to_bytes(pr["evals"][k]) for k in EVAL_BYTES_ORDER)
    assert len(out) == 1040
    return out


# ============================================================================= Fp2 / Fp6 / Fp12, G2, pairing
def _f2(a, b=0):
    return (a % P, b % P)


def f2_add(x, y): return ((x[0] + y[0]) % P, (x[1] + y[1]) % P)
def f2_sub(x, y): return ((x[0] - y[0]) % P, (x[1] - y[1]) % P)
def f2_neg(x): return ((-x[0]) % P, (-x[1]) % P)
def f2_mul(x, y): return ((x[0] * y[0] - x[1] * y[1]) % P, (x[0] * y[1] + x[1] * y[0]) % P)
def f2_sqr(x): return f2_mul(x, x)
def f2_muls(x, k): return (x[0] * k % P, x[1] * k % P)


def f2_inv(x):
    t = pow(x[0] * x[0] + x[1] * x[1], -1, P)
    return (x[0] * t % P, (-x[1]) * t % P)


F2_ZERO, F2_ONE = (0, 0), (1, 0)
XI = (1, 1)  # the non-residue u + 1: Fp6 = Fp2[v]/(v³ − ξ), Fp12 = Fp6[w]/(w² − v)

# Fp12 as polynomials in w of degree < 6 over Fp2 with w⁶ = ξ (w² = v).
F12_ONE = (F2_ONE,) + (F2_ZERO,) * 5


def f12_mul(x, y):
    acc = [F2_ZERO] * 11
    for i, xi in enumerate(x):
        if xi == F2_ZERO:
            continue
        for j, yj in enumerate(y):
            if yj == F2_ZERO:
                continue
            acc[i + j] = f2_add(acc[i + j], f2_mul(xi, yj))
    for k in range(10, 5, -1):
        acc[k - 6] = f2_add(acc[k - 6], f2_mul(acc[k], XI))
    return tuple(acc[:6])


def f12_pow(x, e):
    r, b = F12_ONE, x
    while e:
        if e & 1:
            r = f12_mul(r, b)
        b = f12_mul(b, b)
        e >>= 1
    return r


def f12_conj(x):  # the p⁶-Frobenius: w ↦ −w
    return tuple(c if i % 2 == 0 else f2_neg(c) for i, c in enumerate(x))


def f12_inv(x):
    """x⁻¹ via the norm to Fp6 … kept simple: x^(p¹²−2) is far too slow, so solve with conjugates:
    x·conj(x) lies in Fp6 (even powers of w only); invert there by a 3×3 system over Fp2."""
    n = f12_mul(x, f12_conj(x))
    a0, a1, a2 = n[0], n[2], n[4]  # a0 + a1 v + a2 v²
    t0 = f2_sub(f2_sqr(a0), f2_mul(XI, f2_mul(a1, a2)))
    t1 = f2_sub(f2_mul(XI, f2_sqr(a2)), f2_mul(a0, a1))
    t2 = f2_sub(f2_sqr(a1), f2_mul(a0, a2))
    den = f2_add(f2_mul(a0, t0), f2_mul(XI, f2_add(f2_mul(a2, t1), f2_mul(a1, t2))))
    di = f2_inv(den)
    ninv = (f2_mul(t0, di), F2_ZERO, f2_mul(t1, di), F2_ZERO, f2_mul(t2, di), F2_ZERO)
    return f12_mul(f12_conj(x), ninv)


G2_GEN = ((0x024AA2B2F08F0A91260805272DC51051C6E47AD4FA403B02B4510B647AE3D1770BAC0326A805BBEFD48056C8C121BDB8,
           0x13E02B6052719F607DACD3A088274F65596BD0D09920B61AB5DA61BBDC7F5049334CF11213945D57E5AC7D055D042B7E),
          (0x0CE5D527727D6E118CC9CDC6DA2E351AADFD9BAA8CBDD3A76D429A695160D12C923AC9CC3BACA289E193548608B82801,
           0x0606C4A02EA734CC32ACD2B02BC28B99CB3E287E85A763AF267492AB572E99AB3F370D275CEC1DA1AAA9075FF05F79BE))
B2 = (4, 4)  # twist: y² = x³ + 4(u + 1)


def g2_on_curve(pt):
    if pt is None:
        return True
    x, y = pt
    return f2_sub(f2_sqr(y), f2_add(f2_mul(f2_sqr(x), x), B2)) == F2_ZERO


def g2_add(a, b):
    if a is None:
        return b
    if b is None:
        return a
    (x1, y1), (x2, y2) = a, b
    if x1 == x2:
        if f2_add(y1, y2) == F2_ZERO:
            return None
        lam = f2_mul(f2_muls(f2_sqr(x1), 3), f2_inv(f2_muls(y1, 2)))
    else:
        lam = f2_mul(f2_sub(y2, y1), f2_inv(f2_sub(x2, x1)))
    x3 = f2_sub(f2_sub(f2_sqr(lam), x1), x2)
    return (x3, f2_sub(f2_mul(lam, f2_sub(x1, x3)), y1))


def g2_mul(pt, k):
    acc, add = None, pt
    while k:
        if k & 1:
            acc = g2_add(acc, add)
        add = g2_add(add, add)
        k >>= 1
    return acc


BLS_X = 0xD201000000010000  # |x|; the curve parameter is −|x|


def miller_loop(p1, q2):
    """Ate Miller loop f_{|x|,Q}(P) for P ∈ G1 (affine ints), Q ∈ G2 on the twist (affine Fp2).  With the M-twist
    untwisting (x, y) ↦ (x/w², y/w³), the line through T with slope λ evaluated at P, scaled by w³ (a factor
    the final exponentiation kills), is  (λ·x_T − y_T) − λ·x_P·w² + y_P·w³.  The sign of x only conjugates the
    result, which does not change whether a product of pairings is one."""
    if p1 is None or q2 is None:
        return F12_ONE
    xp, yp = p1
    f = F12_ONE
    t = q2

    def line(t, lam):
        c0 = f2_sub(f2_mul(lam, t[0]), t[1])
        return (c0, F2_ZERO, f2_muls(f2_neg(lam), xp), (yp, 0), F2_ZERO, F2_ZERO)

    for bit in bin(BLS_X)[3:]:
        lam = f2_mul(f2_muls(f2_sqr(t[0]), 3), f2_inv(f2_muls(t[1], 2)))
        f = f12_mul(f12_mul(f, f), line(t, lam))
        t = g2_add(t, t)
        if bit == "1":
            lam = f2_mul(f2_sub(q2[1], t[1]), f2_inv(f2_sub(q2[0], t[0])))
            f = f12_mul(f, line(t, lam))
            t = g2_add(t, q2)
    return f


def final_exponentiation(f):
    # easy part (p⁶ − 1): conj(f)/f; the rest (p⁶ + 1)/r by plain exponentiation (test infrastructure: speed is irrelevant)
    f = f12_mul(f12_conj(f), f12_inv(f))
    return f12_pow(f, (P ** 6 + 1) // R)


def pairing_product_is_one(pairs):
    f = F12_ONE
    for p1, q2 in pairs:
        f = f12_mul(f, miller_loop(p1, q2))
    return final_exponentiation(f) == F12_ONE


# ============================================================================= verifier
def bytes_to_g1(b):
    """G1Affine::from_bytes (compressed); returns the point or raises ValueError."""
    if len(b) != 48 or not b[0] & 0x80:
        raise ValueError("not a compressed G1 point")
    if b[0] & 0x40:
        if b[0] != 0xC0 or any(b[1:]):
            raise ValueError("bad identity encoding")
        return None
    x = int.from_bytes(bytes([b[0] & 0x1F]) + bytes(b[1:]), "big")
    if x >= P:
        raise ValueError("x out of range")
    y = pow((x * x * x + 4) % P, (P + 1) // 4, P)
    if (y * y - x * x * x - 4) % P:
        raise ValueError("not on curve")
    if (y > (P - 1) // 2) != bool(b[0] & 0x20):
        y = P - y
    return (x, y)


def proof_from_bytes(b):
    assert len(b) == 1040
    pts = [bytes_to_g1(b[48 * i:48 * i + 48]) for i in range(11)]
    sc = []
    for i in range(16):
        v = int.from_bytes(b[528 + 32 * i:528 + 32 * i + 32], "little")
        if v >= R:
            raise ValueError("non-canonical scalar")
        sc.append(v)
    return {"a_comm": pts[0], "b_comm": pts[1], "c_comm": pts[2], "d_comm": pts[3], "z_comm": pts[4], "t_comm": pts[5:9],
            "w_z_comm": pts[9], "w_zw_comm": pts[10], "evals": dict(zip(EVAL_BYTES_ORDER, sc))}


def verify(vk, proof_bytes, pub_inputs, opening_key, label):
    """Proof::verify.  pub_inputs: {gate index: value}; opening_key = (H, τ·H) on G2.
    Returns True / False."""
    try:
        pr = proof_from_bytes(proof_bytes)
    except ValueError:
        return False
    t = Transcript(label)
    seed_transcript(t, vk)
    n = vk["n"]
    d = domain(n)
    ev = pr["evals"]
    for lab, key in ((b"w_l", "a_comm"), (b"w_r", "b_comm"), (b"w_o", "c_comm"), (b"w_4", "d_comm")):
        t.append_commitment(lab, pr[key])
    beta = t.challenge_scalar(b"beta")
    t.append_scalar(b"beta", beta)
    gamma = t.challenge_scalar(b"gamma")
    t.append_commitment(b"z", pr["z_comm"])
    alpha = t.challenge_scalar(b"alpha")
    range_sep = t.challenge_scalar(b"range separation challenge")
    logic_sep = t.challenge_scalar(b"logic separation challenge")
    fixed_sep = t.challenge_scalar(b"fixed base separation challenge")
    var_sep = t.challenge_scalar(b"variable base separation challenge")
    for lab, c in zip((b"t_1", b"t_2", b"t_3", b"t_4"), pr["t_comm"]):
        t.append_commitment(lab, c)
    z = t.challenge_scalar(b"z")
    lin = linearisation_terms(ev, alpha, beta, gamma, (range_sep, logic_sep, fixed_sep, var_sep), z, d)
    z_h, l1 = lin["z_h"], lin["l1"]
    if z_h == 0:
        return False
    # PI(z) by the barycentric formula
    w = d["group_gen"]
    pi_eval = 0
    for i, v in pub_inputs.items():
        wi = pow(w, i, R)
        pi_eval = (pi_eval + v * wi % R * pow((z - wi) % R, -1, R)) % R
    pi_eval = pi_eval * z_h % R * d["size_inv"] % R
    a, b, c, dd = ev["a_eval"], ev["b_eval"], ev["c_eval"], ev["d_eval"]
    bprod = ((a + beta * ev["left_sigma_eval"] + gamma) * (b + beta * ev["right_sigma_eval"] + gamma) % R
             * (c + beta * ev["out_sigma_eval"] + gamma) % R * ((dd + gamma) * ev["perm_eval"] % R * alpha % R)) % R
    t_eval = (ev["lin_poly_eval"] + pi_eval - bprod - l1 * alpha * alpha) % R * pow(z_h, -1, R) % R
    zn = pow(z, n, R)
    t_comm = msm_naive(pr["t_comm"], [1, zn, zn * zn % R, pow(zn, 3, R)])
    for lab, key in EVAL_TRANSCRIPT_ORDER:
        t.append_scalar(lab, t_eval if key == "t_eval" else ev[key])
    names = list(lin["q"].keys())
    r_comm = msm_naive([vk["q"][k] for k in names] + [pr["z_comm"], vk["sigma"][3]],
                       [lin["q"][k] for k in names] + [lin["z"], lin["sigma4"]])

    def flatten(parts):
        v = t.challenge_scalar(b"aggregate_witness")
        pw, cm, evl = 1, None, 0
        for e, c in parts:
            cm = g1_add(cm, g1_mul(c, pw))
            evl = (evl + e * pw) % R
            pw = pw * v % R
        return cm, evl

    ca, ea = flatten([(t_eval, t_comm), (ev["lin_poly_eval"], r_comm), (a, pr["a_comm"]), (b, pr["b_comm"]),
                      (c, pr["c_comm"]), (dd, pr["d_comm"]), (ev["left_sigma_eval"], vk["sigma"][0]),
                      (ev["right_sigma_eval"], vk["sigma"][1]), (ev["out_sigma_eval"], vk["sigma"][2])])
    cb, eb = flatten([(ev["perm_eval"], pr["z_comm"]), (ev["a_next_eval"], pr["a_comm"]),
                      (ev["b_next_eval"], pr["b_comm"]), (ev["d_next_eval"], pr["d_comm"])])
    t.append_commitment(b"w_z", pr["w_z_comm"])
    t.append_commitment(b"w_z_w", pr["w_zw_comm"])
    # OpeningKey::batch_check
    u = t.challenge_scalar(b"batch")
    total_c, total_w, g_mult = None, None, 0
    for (cm, evl, wit, point), pw in zip(((ca, ea, pr["w_z_comm"], z), (cb, eb, pr["w_zw_comm"], z * w % R)), (1, u)):
        cc = g1_add(cm, g1_mul(wit, point))
        g_mult = (g_mult + pw * evl) % R
        total_c = g1_add(total_c, g1_mul(cc, pw))
        total_w = g1_add(total_w, g1_mul(wit, pw))
    total_c = g1_add(total_c, g1_neg(g1_mul(G1_GEN, g_mult)))
    h, beta_h = opening_key
    return pairing_product_is_one([(g1_neg(total_w), beta_h), (total_c, h)])


def opening_key(tau):
    return (G2_GEN, g2_mul(G2_GEN, tau % R))


# ============================================================================= synthetic circuit
def synthetic_circuit(n_gates, seed=0x5EED, n_pub=2):
    """Chain x_{i+1} = x_i·x_i + x_i + c_i built from `mul` / `add` gates — the gate shapes of
    /root/reference/src/zk/gadgets.rs:60,70,81 — with a few public inputs; total gate count = n_gates."""
    from model import random_fr
    comp = Composer()
    consts = random_fr(seed, 4)
    x = comp.add_input(consts[0])
    k = 0
    while comp.n + 2 <= n_gates - n_pub:
        sq = comp.mul(1, x, x, 0, 0)
        x = comp.add((1, sq), (1, x), (consts[1] + k) % R, 0)
        k += 1
    while comp.n < n_gates - n_pub:
        comp.boolean_gate(comp.zero_var)
    for j in range(n_pub):
        v = comp.values[x] if j == 0 else consts[2]
        var = x if j == 0 else comp.add_input(v)
        comp.constrain_to_constant(var, 0, -v)  # var − 0 + PI = 0 with PI = −v
    assert comp.n == n_gates and comp.check()
    return comp
