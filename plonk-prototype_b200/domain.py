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


_default = None


def default_context():
    global _default
    if _default is None:
        _default = _native.Context(0)
    return _default
