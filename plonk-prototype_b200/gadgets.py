"""The reference's own gadgets and circuit, restated over the `StandardComposer` mirror so that they can be proved
through this backend: /root/reference/src/zk/gadgets.rs (`maybe_equal` :49-84, `range_check` :94-109, `min_bound`
:120-145, `max_bound` :151-184, `scalar_decomposition_gadget` :186-225, `scalar_to_bits` :228-237, `bits_count`
:240-248, `num_bits_closest_power_of_two` :252-256), /root/reference/src/zk/allocated_scalar.rs:25-38 and
`MockCircuit::valid_balance` (/root/reference/src/zk/circuits.rs:51-60).

These functions only append rows to the composer (host-side bookkeeping, as in the reference); behaviours a maintainer
would expect to carry over are kept on purpose: the decomposition allocates all 256 bit variables before truncating,
`min_bound` / `max_bound` pass a zero coefficient on the witness as their right operand, `max_bound` subtracts one from
the bound first, and `bits_count` works on the canonical value starting from one.

`commitment_gadget` (gadgets.rs:28-41) and `MockCircuit::prove_ownership` (circuits.rs:63-66) go through the composer's
`fixed_base_scalar_mul` / `point_addition_gate` / `assert_equal_public_point` and the GPU's fixed-base / variable-base
widgets (csrc/widgets.h).  `check_hash_inputs` (circuits.rs:69-72) uses the Poseidon sponge of
plonk-prototype_b200/poseidon.py, whose constants are regenerated from the recalled recipe of dusk-hades (the crate is not
on disk, Cargo.toml:23, and no test vector could be checked): it constrains the function that module's `hash` computes —
unpinned against the Rust crate.
"""
from .jubjub import GENERATOR as GENERATOR_EXTENDED, GENERATOR_NUMS as GENERATOR_NUMS_EXTENDED
from .prover import R


class AllocatedScalar:
    """A variable together with its witness assignment (allocated_scalar.rs:25-38)."""

    def __init__(self, var, scalar):
        self.var, self.scalar = var, scalar % R

    @classmethod
    def allocate(cls, composer, scalar):
        return cls(composer.add_input(scalar), scalar)


def scalar_to_bits(scalar):
    """256 little-endian bits of the canonical value (bit i of byte k = bit 8k + i)."""
    return [(scalar % R >> i) & 1 for i in range(256)]


def bits_count(scalar):
    scalar %= R
    counter = 1
    while scalar > 1:
        scalar >>= 1
        counter += 1
    return counter


def num_bits_closest_power_of_two(scalar):
    return bits_count(pow(2, bits_count(scalar), R))


def maybe_equal(composer, a, b):
    """1 if a = b, 0 otherwise: u = a − b, z = u⁻¹ (0 for 0), y = 1 − u·z, and y·u = 0."""
    u = composer.add((1, a.var), (-1, b.var), 0, None)
    u_scalar = (a.scalar - b.scalar) % R
    z = composer.add_input(pow(u_scalar, -1, R) if u_scalar else 0)
    y = composer.mul(-1, z, u, 1, None)
    composer.mul_gate(y, u, u, 1, 0, 0, None)
    return y


def scalar_decomposition_gadget(composer, num_bits, witness):
    bits = scalar_to_bits(witness.scalar)
    bit_vars = [composer.add_input(bit) for bit in bits][:num_bits]   # all 256 are allocated, then truncated
    acc = AllocatedScalar(composer.add_witness_to_circuit_description(0), 0)
    for power, bit in enumerate(bit_vars):
        composer.boolean_gate(bit)
        two_pow = pow(2, power, R)
        acc.var = composer.add((two_pow, bit), (1, acc.var), 0, None)
        acc.scalar = (acc.scalar + two_pow * bits[power]) % R
    return maybe_equal(composer, acc, witness), bit_vars


def range_proof(composer, value, num_bits):
    is_equal, _ = scalar_decomposition_gadget(composer, num_bits, value)
    return is_equal


def min_bound(composer, min_range, witness, num_bits):
    """1 if witness ≥ min_range (as num_bits-bit quantities), else 0."""
    var = composer.add((1, witness.var), (0, witness.var), -min_range, None)
    return range_proof(composer, AllocatedScalar(var, witness.scalar - min_range), num_bits)


def max_bound(composer, max_range, witness):
    """(1 if witness < max_range else 0, number of bits used)."""
    max_range = (max_range - 1) % R
    num_bits = num_bits_closest_power_of_two(max_range)
    var = composer.add((-1, witness.var), (0, witness.var), max_range, None)
    return range_proof(composer, AllocatedScalar(var, max_range - witness.scalar), num_bits), num_bits


def range_check(composer, min_range, max_range, witness):
    y1, num_bits = max_bound(composer, max_range, witness)
    y2 = min_bound(composer, min_range, witness, num_bits)
    return composer.mul(1, y1, y2, 0, None)


def commitment_gadget(composer, value, blinder):
    """In-circuit Pedersen commitment value·G + blinder·G_nums (gadgets.rs:28-41); `value`, `blinder` are variables."""
    p1 = composer.fixed_base_scalar_mul(value, GENERATOR_EXTENDED)
    p2 = composer.fixed_base_scalar_mul(blinder, GENERATOR_NUMS_EXTENDED)
    return composer.point_addition_gate(p1, p2)


class MockCircuit:
    def __init__(self, note_value, private_key=None, hash_inputs=(), public_key=None):
        self.note_value, self.private_key, self.hash_inputs, self.public_key = note_value, private_key, list(hash_inputs), public_key

    def valid_balance(self, composer, tx_value, gas_fee):
        """The note value covers the transaction and its gas: min_bound(tx + gas, note, 30 bits).  As in the reference the
        0/1 output is returned to the caller and not itself constrained."""
        return min_bound(composer, (tx_value + gas_fee) % R, self.note_value, 30)

    def prove_ownership(self, composer):
        """Public-key ownership (circuits.rs:63-66): private_key·G computed in circuit equals the public key (public input)."""
        circuit_pk = composer.fixed_base_scalar_mul(self.private_key, GENERATOR_EXTENDED)
        composer.assert_equal_public_point(circuit_pk, self.public_key)

    def check_hash_inputs(self, composer, public_hash):
        """Constrains a public hash to the Poseidon sponge of the private `hash_inputs` (circuits.rs:69-72).  The sponge is
        plonk-prototype_b200/poseidon.py — constants and padding restated from memory of dusk-poseidon / dusk-hades, unpinned."""
        from . import poseidon
        h = poseidon.gadget(composer, self.hash_inputs)
        composer.constrain_to_constant(h, 0, -public_hash)
