"""Phase times of the batched, pre-doubled MSM (the prover's commit shape) as the per-GPU slice shrinks: what a rank of a
sharded prove runs at N = 1, 2, 4, 8 (slice = 2^20 / N points), batch 4 / 1 / 2.  One GPU.
    python scripts/msm_tail_scaling.py > gpurun_out/msm_tail_scaling.json"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import plonk_prototype_b200 as pb  # noqa: E402

ctx = pb.Context(0)
out = []
for log_n in ([int(v) for v in sys.argv[1:]] or [20, 19, 18, 17, 16]):
    n = 1 << log_n
    pp = pb.PublicParameters(n - 1, 0xB2 + log_n, ctx)
    for batch in (4, 1):
        sc = bench.random_fr_limbs(0x5CA1 + log_n, n * batch)
        d = ctx.malloc(sc.nbytes)
        ctx.h2d(d, sc)
        for _ in range(3):
            ctx.msm_batch_dev(pp.srs, d, n, batch, n)
        ctx.profile_enable(True)
        ctx.profile_reset()
        reps = 10
        for _ in range(reps):
            ctx.msm_batch_dev(pp.srs, d, n, batch, n)
        row = {"log_n": log_n, "batch": batch}
        for k in ("sort", "accumulate", "partials", "reduce", "total"):
            row[k + "_ms"] = round(ctx.profile_sum_ms("msm." + k)[0] / reps, 4)
        ctx.profile_enable(False)
        ctx.free(d)
        out.append(row)
        print(json.dumps(row), flush=True)
    pp.close()
ctx.close()
