"""Times pb200_preprocess / pb200_prove on the synthetic arithmetic circuit at 2^k gates.
usage: python scripts/prove_bench.py 16 18 20"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import plonk_prototype_b200 as pb  # noqa: E402
from plonk_prototype_b200.synth import synthetic_circuit_columns  # noqa: E402


def main():
    ctx = pb.Context(0)
    for log_n in [int(a) for a in sys.argv[1:]] or [16]:
        n = 1 << log_n
        t0 = time.time()
        sel, wires, values, pi_pos, pi_vals = synthetic_circuit_columns(n)
        t_build = time.time() - t0
        pp = pb.PublicParameters(n - 1, 0xB200 + log_n, ctx)
        ctx.sync()
        t0 = time.time()
        pk, vk = ctx.preprocess(pp.srs, sel, wires, values.shape[0], b"pb200-bench")
        t_pre = time.time() - t0
        ctx.profile_enable(True)
        times = []
        for it in range(4):
            t0 = time.time()
            proof = ctx.prove(pp.srs, pk, values, pi_pos, pi_vals)
            times.append(time.time() - t0)
        rounds = {k: ctx.profile_ms("prove.round%d" % k) for k in range(1, 6)}
        ctx.profile_enable(False)
        t0 = time.time()
        proof2 = ctx.prove(pp.srs, pk, values, pi_pos, pi_vals)
        t_plain = time.time() - t0
        assert proof2 == proof
        print(json.dumps({"log_n": log_n, "build_s": round(t_build, 2), "preprocess_ms": round(1e3 * t_pre, 1),
                          "prove_ms": [round(1e3 * t, 2) for t in times], "prove_ms_noprofile": round(1e3 * t_plain, 2),
                          "rounds_ms": rounds, "pk_gib": round(ctx.prover_key_bytes(pk) / 2**30, 2),
                          "proof_head": proof[:8].hex()}), flush=True)
        ctx.prover_key_free(pk)
        pp.close()


if __name__ == "__main__":
    main()
