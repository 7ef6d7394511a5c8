"""JubJub (dusk-jubjub 0.10, /root/reference/Cargo.toml:21) on the host: the twisted Edwards curve
−x² + y² = 1 + d·x²·y² over the BLS12-381 scalar field, d = −10240/10241, which the reference's gadgets use through
`GENERATOR_EXTENDED` / `GENERATOR_NUMS_EXTENDED` (/root/reference/src/zk/gadgets.rs:21,34,37; circuits.rs:64).

Only witness generation needs it (circuit synthesis is host-side bookkeeping in the reference too): the composer
computes the accumulator points of the fixed-base ladder and the sum of `point_addition_gate`; the constraints
themselves are checked on the GPU by the widgets in csrc/widgets.h.

Constants: GENERATOR is the point with y = 18 (the x below is the root of the curve equation for that y that lies in
the prime-order subgroup and starts 0x3fd2…); GENERATOR_NUMS as published by dusk-jubjub.  Both are verified on-curve
and of order JJ_ORDER by tests/test_widgets_cpu.py — pinned by the curve equation, not by memory."""
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
EDWARDS_D = (-10240 * pow(10241, -1, R)) % R
JJ_ORDER = 0x0E7DB4EA6533AFA906673B0101343B00A6682093CCC81082D0970E5ED6F72CB7
GENERATOR = (0x3FD2814C43AC65A6F1FBF02D0FD6CCE62E3EBB21FD6C54ED4DF7B7FFEC7BEACA, 0x12)
GENERATOR_NUMS = (0x5E67B8F316F414F7BD9514C773FD4456931E316A39FE4541921710179DF76377,
                  0x43D80EB3B2F3EB1B7B162DBEEB3B34FD9949BA0F82A5507A6705B707162E3EF8)
IDENTITY = (0, 1)


def add(p, q):
    (x1, y1), (x2, y2) = p, q
    t = EDWARDS_D * x1 % R * x2 % R * y1 % R * y2 % R
    return ((x1 * y2 + y1 * x2) * pow(1 + t, -1, R) % R, (y1 * y2 + x1 * x2) * pow(1 - t, -1, R) % R)


def neg(p):
    return ((-p[0]) % R, p[1])


def mul(p, k):
    acc = IDENTITY
    while k:
        if k & 1:
            acc = add(acc, p)
        p = add(p, p)
        k >>= 1
    return acc


def on_curve(p):
    x, y = p
    return (-x * x + y * y) % R == (1 + EDWARDS_D * x * x % R * y * y) % R


def wnaf2(k):
    """`Fr::compute_windowed_naf(2)`: 256 digits in {−1, 0, 1}, least significant first."""
    out = [0] * 256
    i = 0
    while k >= 1:
        if k & 1:
            d = 2 - (k & 3)
            out[i] = d
            k -= d
        k >>= 1
        i += 1
    return out
