"""Host-side mirror of dusk-plonk 0.8.2 `fft::EvaluationDomain` (pinned at /root/reference/Cargo.toml:19;
interface restated in SURVEY.md §8b and App. B.2) on top of the C ABI.

Same names, argument meaning and error behaviour as the Rust type, so the parity tests read like the
upstream ones: `EvaluationDomain(n)` raises for log2(size) ≥ 32; `fft / ifft / coset_fft / coset_ifft`
take a slice of BlsScalar (here an (k, 4) uint64 array of Montgomery limbs, k ≤ size), zero-pad it to
`size`, and return a new vector in natural order.  All arithmetic happens on the GPU.
"""
import ctypes

import numpy as np

from . import _native


class InvalidEvalDomainSize(ValueError):
    """`Error::InvalidEvalDomainSize` of dusk-plonk."""


class EvaluationDomain:
    def __init__(self, num_coeffs, ctx=None):
        log_n = ctypes.c_uint32()
        if _native.lib().pb200_domain_log_size(int(num_coeffs), ctypes.byref(log_n)) != 0:
            raise InvalidEvalDomainSize("log2(size) >= 32 for %d coefficients" % num_coeffs)
        self.log_size_of_group = log_n.value
        self.size = 1 << log_n.value
        self.ctx = ctx or default_context()

    def _run(self, coeffs, inverse, coset):
        a = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, 4)
        if a.shape[0] > self.size:
            raise ValueError("more coefficients than the domain size")
        buf = np.zeros((self.size, 4), np.uint64)  # `resize(size, zero)`
        buf[: a.shape[0]] = a
        self.ctx.ntt(buf, self.log_size_of_group, inverse, coset)
        return buf

    def fft(self, coeffs):
        return self._run(coeffs, False, False)

    def ifft(self, evals):
        return self._run(evals, True, False)

    def coset_fft(self, coeffs):
        return self._run(coeffs, False, True)

    def coset_ifft(self, evals):
        return self._run(evals, True, True)


class Polynomial:
    """`fft::Polynomial { coeffs: Vec<BlsScalar> }` (dusk-plonk 0.8.2 `fft/polynomial.rs`, pinned at
    /root/reference/Cargo.toml:19): dense coefficient form, lowest degree first, no trailing zeros.
    Coefficients are an (k, 4) uint64 array of Montgomery limbs.  `evaluate` and `ruffini` run on the GPU
    (`pb200_kzg_witness_dev`: the evaluation and the quotient by X − z come out of the same Ruffini pass)."""

    def __init__(self, coeffs, ctx=None):
        self.coeffs = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, 4)
        self.ctx = ctx
        assert self.coeffs.shape[0] == 0 or self.coeffs[-1].any(), "leading coefficient must be non-zero"

    @classmethod
    def zero(cls, ctx=None):
        return cls(np.zeros((0, 4), np.uint64), ctx)

    @classmethod
    def from_coefficients_vec(cls, coeffs, ctx=None):
        """Drops the zero coefficients at the top (`truncate_leading_zeros`)."""
        a = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, 4)
        nz = np.flatnonzero(a.any(axis=1))
        return cls(a[: (nz[-1] + 1) if nz.size else 0].copy(), ctx)

    from_coefficients_slice = from_coefficients_vec

    def is_zero(self):
        return self.coeffs.shape[0] == 0

    def degree(self):
        """0 for the zero polynomial, like upstream."""
        return 0 if self.is_zero() else self.coeffs.shape[0] - 1

    def __len__(self):
        return self.coeffs.shape[0]

    def _ruffini(self, z_mont):
        ctx = self.ctx or default_context()
        n = self.coeffs.shape[0]
        d_p, d_q = ctx.malloc(32 * n), ctx.malloc(32 * n)
        try:
            ctx.h2d(d_p, self.coeffs)
            ev = ctx.kzg_witness_dev(d_p, n, z_mont, d_q)
            q = np.empty((n, 4), np.uint64)
            ctx.d2h(q, d_q)
        finally:
            ctx.free(d_p)
            ctx.free(d_q)
        return ev, q

    def evaluate(self, point_mont):
        """p(point); `point_mont` and the result are 4 Montgomery limbs."""
        if self.is_zero():
            return np.zeros(4, np.uint64)
        return self._ruffini(np.ascontiguousarray(point_mont, dtype=np.uint64).reshape(4))[0]

    def ruffini(self, z_mont):
        """(p(X) − p(z)) / (X − z) as a Polynomial (`Polynomial::ruffini`)."""
        if self.is_zero():
            return Polynomial.zero(self.ctx)
        q = self._ruffini(np.ascontiguousarray(z_mont, dtype=np.uint64).reshape(4))[1]
        return Polynomial.from_coefficients_vec(q[:-1], self.ctx)   # the library writes n scalars, the top one zero


class Evaluations:
    """`fft::Evaluations { evals, domain }` (dusk-plonk 0.8.2 `fft/evaluations.rs`): values over the whole domain."""

    def __init__(self, evals, domain):
        self.evals = np.ascontiguousarray(evals, dtype=np.uint64).reshape(-1, 4)
        self.domain = domain

    from_vec_and_domain = classmethod(lambda cls, evals, domain: cls(evals, domain))

    def interpolate_by_ref(self):
        return Polynomial.from_coefficients_vec(self.domain.ifft(self.evals), self.domain.ctx)

    interpolate = interpolate_by_ref


_default = None


def default_context():
    global _default
    if _default is None:
        _default = _native.Context(0)
    return _default
