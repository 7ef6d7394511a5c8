"""Poseidon sponge over the BLS12-381 scalar field as the reference uses it: `dusk_poseidon::sponge::gadget(composer,
&[Variable]) -> Variable` (/root/reference/src/zk/circuits.rs:69-72; crates dusk-poseidon 0.22 / dusk-hades, pinned at
/root/reference/Cargo.toml:23) — the Hades permutation (width 5, 8 full + 59 partial rounds, quintic S-box, Cauchy MDS
matrix) inside a rate-4 sponge, on the host (`hash`) and as a circuit (`gadget`).

**UNPINNED, and more so than the rest of this repository**: the crates are not on disk, and this hash has no published
test vector that could be checked here.  Everything below is restated from memory of the crates' sources:
  * round constants: 960 scalars, c_k = from_bytes_wide(SHA-512 chain seeded with b"poseidon-for-plonk") + c_{k−1}, c_{−1} = 1
    (dusk-hades' `assets` generator);
  * MDS: M[i][j] = 1 / (x_i + y_j) with x_i = i, y_j = WIDTH + j;
  * permutation: 4 full rounds, 59 partial rounds (S-box on the last word), 4 full rounds; every round = add round keys,
    S-box, MDS multiplication;
  * sponge: state[0] is the capacity, messages are added to state[1..] in chunks of 4, a final `1` marks the end of the
    message (in the next free slot, or after an extra permutation when the last chunk is full), output state[1].
The gate layout of the gadget is this file's own (one `add` row per round key, three `mul` rows per S-box, two `big_add` rows
per output word of the matrix product); it constrains the same function `hash` computes, which tests/test_widgets_cpu.py
checks.  A maintainer with the crates on disk should diff constants, sponge padding and gate layout before relying on any
byte produced through this module (UPSTREAM_ASSUMPTIONS.md §Poseidon)."""
import hashlib

from .prover import R

WIDTH = 5
TOTAL_FULL_ROUNDS = 8
PARTIAL_ROUNDS = 59
CONSTANTS = 960

_cache = {}


def round_constants():
    if "ark" not in _cache:
        out, p, data = [], 1, b"poseidon-for-plonk"
        for _ in range(CONSTANTS):
            data = hashlib.sha512(data).digest()
            p = (int.from_bytes(data, "little") + p) % R      # BlsScalar::from_bytes_wide + previous constant
            out.append(p)
        _cache["ark"] = out
    return _cache["ark"]


def mds_matrix():
    if "mds" not in _cache:
        _cache["mds"] = [[pow(i + (WIDTH + j), -1, R) for j in range(WIDTH)] for i in range(WIDTH)]
    return _cache["mds"]


def permutation(words):
    """Hades permutation on a list of WIDTH integers (ScalarStrategy::perm)."""
    ark, m = iter(round_constants()), mds_matrix()
    w = [v % R for v in words]

    def round_(full):
        nonlocal w
        w = [(v + next(ark)) % R for v in w]
        if full:
            w = [pow(v, 5, R) for v in w]
        else:
            w[WIDTH - 1] = pow(w[WIDTH - 1], 5, R)
        w = [sum(m[j][k] * w[k] for k in range(WIDTH)) % R for j in range(WIDTH)]

    for _ in range(TOTAL_FULL_ROUNDS // 2):
        round_(True)
    for _ in range(PARTIAL_ROUNDS):
        round_(False)
    for _ in range(TOTAL_FULL_ROUNDS // 2):
        round_(True)
    return w


def _last_iteration(l):
    m = l // (WIDTH - 1)
    return m - 1 if l == m * (WIDTH - 1) else m


def hash(messages):
    """`sponge::hash(&[BlsScalar]) -> BlsScalar`."""
    assert len(messages) >= 1
    words = [0] * WIDTH
    last = _last_iteration(len(messages))
    for i in range(0, len(messages), WIDTH - 1):
        chunk = messages[i:i + WIDTH - 1]
        for k, c in enumerate(chunk):
            words[1 + k] = (words[1 + k] + c) % R
        if i // (WIDTH - 1) == last:
            if len(chunk) < WIDTH - 1:
                words[len(chunk) + 1] = (words[len(chunk) + 1] + 1) % R
            else:
                words = permutation(words)
                words[1] = (words[1] + 1) % R
        words = permutation(words)
    return words[1]


def _perm_gadget(cs, words):
    """GadgetStrategy::perm: the same rounds as `permutation`, as rows of the composer.  words: list of WIDTH variables."""
    ark, m = iter(round_constants()), mds_matrix()
    zero = cs.zero_var

    def round_(full):
        nonlocal words
        words = [cs.add((1, w), (0, zero), next(ark), None) for w in words]
        for k in (range(WIDTH) if full else (WIDTH - 1,)):
            v = words[k]
            v2 = cs.mul(1, v, v, 0, None)
            v4 = cs.mul(1, v2, v2, 0, None)
            words[k] = cs.mul(1, v4, v, 0, None)
        out = []
        for j in range(WIDTH):
            z3 = cs.big_add((m[j][0], words[0]), (m[j][1], words[1]), (m[j][2], words[2]), 0, None)
            out.append(cs.big_add((m[j][3], words[3]), (m[j][4], words[4]), (1, z3), 0, None))
        words = out

    for _ in range(TOTAL_FULL_ROUNDS // 2):
        round_(True)
    for _ in range(PARTIAL_ROUNDS):
        round_(False)
    for _ in range(TOTAL_FULL_ROUNDS // 2):
        round_(True)
    return words


def gadget(cs, messages):
    """`sponge::gadget(composer, &[Variable]) -> Variable`: constrains the returned variable to `hash` of the messages' values."""
    assert len(messages) >= 1
    words = [cs.zero_var] * WIDTH
    last = _last_iteration(len(messages))
    for i in range(0, len(messages), WIDTH - 1):
        chunk = messages[i:i + WIDTH - 1]
        for k, c in enumerate(chunk):
            words[1 + k] = cs.add((1, words[1 + k]), (1, c), 0, None)
        if i // (WIDTH - 1) == last:
            if len(chunk) < WIDTH - 1:
                k = len(chunk) + 1
                words[k] = cs.add((1, words[k]), (0, cs.zero_var), 1, None)
            else:
                words = _perm_gadget(cs, words)
                words[1] = cs.add((1, words[1]), (0, cs.zero_var), 1, None)
        words = _perm_gadget(cs, words)
    return words[1]
