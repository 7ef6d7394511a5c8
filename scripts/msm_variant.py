"""Times pb200_msm_g1_dev at a few sizes (with / without pre-doubled copies).  Used to compare kernel variants selected by
environment switches (e.g. PB200_MSM_ACC_3CTA=1).  usage: python scripts/msm_variant.py"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import plonk_prototype_b200 as pb  # noqa: E402
from bench import random_fr_limbs  # noqa: E402


def main():
    ctx = pb.Context(0)
    stream = torch.cuda.ExternalStream(ctx.stream)
    out = {"variant": "3cta" if os.environ.get("PB200_MSM_ACC_3CTA") else "2cta"}
    for L, pre in ((20, True), (22, True), (24, False), (26, False)):
        n = 1 << L
        bases = ctx.malloc(n * 96)
        ctx.synthetic_bases_dev(bases, n, 0xB2000001, 0x9E3779B1)
        srs = ctx.srs_wrap_dev(bases, n)
        if pre:
            ctx.srs_precompute(srs)
        s = random_fr_limbs(0xB2000000 + L, n)
        sd = ctx.malloc(n * 32)
        ctx.h2d(sd, s)
        first = ctx.msm_dev(srs, sd, n)
        ctx.msm_dev(srs, sd, n)
        ctx.profile_enable(True)
        ctx.msm_dev(srs, sd, n)
        acc = ctx.profile_ms("msm.accumulate")
        ctx.profile_enable(False)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record(stream)
        for _ in range(reps):
            last = ctx.msm_dev(srs, sd, n)
        e1.record(stream)
        ctx.sync()
        out["2^%d%s" % (L, "pre" if pre else "")] = {"ms": round(e0.elapsed_time(e1) / reps, 3), "accumulate_ms": round(acc, 3),
                                                    "xyz0": int(last[0]) == int(first[0])}
        ctx.srs_free(srs)
        ctx.free(bases)
        ctx.free(sd)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
