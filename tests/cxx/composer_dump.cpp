// Builds a fixed gadget sequence with pb200::StandardComposer (include/pb200.hpp) and prints the column images, so that
// tests/test_cxx_host_layer_cpu.py can compare them with the Python mirror and the model composer.  No GPU, no ABI call.
#include <cstdio>
#include <string>

#include "../../include/pb200.hpp"

using namespace pb200;

static void hex(const BlsScalar &s) {
    for (int i = 0; i < 4; i++) printf("%016llx", (unsigned long long)s.v.l[i]);
}

static void dump(const StandardComposer &c);

int main(int argc, char **argv) {
    if (argc > 1 && std::string(argv[1]) == "ecc") {
        // commitment_gadget (/root/reference/src/zk/gadgets.rs:28-41) + assert_equal_public_point (circuits.rs:65)
        StandardComposer c;
        const Variable value = c.add_input(BlsScalar::from(0xC0FFEE)), blinder = c.add_input(BlsScalar::from(0xB200B200B200ull));
        const Point p1 = c.fixed_base_scalar_mul(value, jubjub::generator());
        const Point p2 = c.fixed_base_scalar_mul(blinder, jubjub::generator_nums());
        const Point p3 = c.point_addition_gate(p1, p2);
        c.assert_equal_public_point(p3, {c.value_of(p3.x), c.value_of(p3.y)});
        dump(c);
        return 0;
    }
    StandardComposer c;
    const Variable a = c.add_input(BlsScalar::from(1234)), b = c.add_input(BlsScalar::from(1200));
    const Variable u = c.add({BlsScalar::one(), a}, {-BlsScalar::one(), b}, BlsScalar::zero());
    const auto inv = BlsScalar::from(34).invert();
    const Variable z = c.add_input(inv.second);
    const Variable y = c.mul(-BlsScalar::one(), z, u, BlsScalar::one());
    c.mul_gate(y, u, u, BlsScalar::one(), BlsScalar::zero(), BlsScalar::zero());
    Variable acc = c.add_witness_to_circuit_description(BlsScalar::zero());
    for (int k = 0; k < 8; k++) {
        const Variable bit = c.add_input(BlsScalar::from((34 >> k) & 1));
        c.boolean_gate(bit);
        acc = c.add({BlsScalar::from(2).pow(k), bit}, {BlsScalar::one(), acc}, BlsScalar::zero());
    }
    const BlsScalar pi = -BlsScalar::from(34);
    c.constrain_to_constant(acc, BlsScalar::zero(), &pi);
    dump(c);
    // BlsScalar helpers used by the reference's host-side code (gadgets.rs:230-256)
    const auto bytes = BlsScalar::from(0x1234).to_bytes();
    printf("to_bytes %02x%02x%02x\n", bytes[0], bytes[1], bytes[2]);
    const auto red = BlsScalar::pow_of_2(70).reduce();
    printf("reduce %llx %llx\n", (unsigned long long)red[0], (unsigned long long)red[1]);
    // fft::Polynomial: zero coefficients at the top are dropped, degree() of the zero polynomial is 0
    const Polynomial p = Polynomial::from_coefficients_vec({BlsScalar::from(5), BlsScalar::zero(), BlsScalar::from(7), BlsScalar::zero(), BlsScalar::zero()});
    const Polynomial pz = Polynomial::from_coefficients_vec({BlsScalar::zero(), BlsScalar::zero()});
    printf("poly %zu %zu %d %zu %d\n", p.len(), p.degree(), (int)p.is_zero(), pz.degree(), (int)pz.is_zero());
    return 0;
}

static void dump(const StandardComposer &c) {
    const pb200_circuit circ = c.circuit();
    printf("n_gates %zu n_vars %zu\n", circ.n_gates, circ.n_vars);
    for (int k = 0; k < 11; k++) {
        printf("sel %d", k);
        if (!circ.selectors[k]) {
            printf(" null\n");
            continue;
        }
        for (size_t i = 0; i < circ.n_gates; i++) {
            printf(" ");
            for (int l = 0; l < 4; l++) printf("%016llx", (unsigned long long)circ.selectors[k][4 * i + l]);
        }
        printf("\n");
    }
    for (int k = 0; k < 4; k++) {
        printf("wire %d", k);
        for (size_t i = 0; i < circ.n_gates; i++) printf(" %u", circ.wires[k][i]);
        printf("\n");
    }
    printf("vars");
    for (const auto &v : c.variables()) {
        printf(" ");
        hex(v);
    }
    printf("\n");
    for (const auto &kv : c.public_inputs_sparse_store()) {
        printf("pi %u ", kv.first);
        hex(kv.second);
        printf("\n");
    }
}
