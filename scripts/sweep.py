"""Size sweep (BASELINE.json configs[1] and [2]): MSM 2^16…2^26 with per-phase device times, NTT 2^16…2^26
for the four EvaluationDomain variants.  Device-resident inputs, CUDA events.  Prints one JSON object."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import plonk_prototype_b200 as pb

ctx = pb.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream)
peak, _ = ctx.imad_peak()
out = {"imad_peak_T": peak / 1e12, "msm": [], "ntt": []}
sizes = [int(x) for x in os.environ.get("SIZES", "16,18,20,22,24,26").split(",")]
maxl = max(sizes)
bases = ctx.malloc((96 << maxl))
ctx.synthetic_bases_dev(bases, 1 << maxl)
s = bench.random_fr_limbs(0xB2000000 + maxl, 1 << maxl)
sd = ctx.malloc(32 << maxl)
ctx.h2d(sd, s)
for L in ([] if os.environ.get("NTT_ONLY") else sizes):
    n = 1 << L
    srs = ctx.srs_wrap_dev(bases, n)
    if os.environ.get("PRECOMP") and L <= 22:
        ctx.srs_precompute(srs)
    for _ in range(2):
        ctx.msm_dev(srs, sd, n)
    ctx.profile_enable(True)
    reps = 5 if L <= 22 else 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ph = {k: 0.0 for k in ("msm.sort", "msm.accumulate", "msm.partials", "msm.reduce")}
    e0.record(stream)
    for _ in range(reps):
        ctx.msm_dev(srs, sd, n)
        for k in ph:
            ph[k] += ctx.profile_ms(k) / reps
    e1.record(stream)
    ctx.sync()
    ctx.profile_enable(False)
    ms = e0.elapsed_time(e1) / reps
    wm = -(-256 // max(L - 4, 1))
    out["msm"].append({"log_n": L, "ms": ms, "mpts": n / ms / 1e3, "window": int(pb._native.lib().pb200_msm_window_bits(n)),
                       "imad_frac": n * wm * 3000.0 / (ms * 1e-3) / peak, **{k.split(".")[1] + "_ms": v for k, v in ph.items()}})
    ctx.srs_free(srs)
    print(json.dumps(out["msm"][-1]), file=sys.stderr, flush=True)
ctx.free(bases)
for L in sizes:
    n = 1 << L
    d = ctx.malloc(32 << L)
    ctx.h2d(d, s[:n])
    row = {"log_n": L}
    for name, inv, cos in (("fft", 0, 0), ("ifft", 1, 0), ("coset_fft", 0, 1), ("coset_ifft", 1, 1)):
        for _ in range(3):
            ctx.ntt_dev(d, L, inv, cos)
        ctx.sync()
        reps = 20 if L <= 22 else 8
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            ctx.ntt_dev(d, L, inv, cos)
        e1.record(stream)
        ctx.sync()
        ms = e0.elapsed_time(e1) / reps
        row[name + "_ms"] = ms
        if name == "fft":
            row["melem"] = n / ms / 1e3
            row["imad_frac"] = (n / 2) * L * 136.0 / (ms * 1e-3) / peak
            row["hbm_frac"] = 64.0 * n * (-(-L // 12)) / (ms * 1e-3) / 1e9 / 6451.5
    out["ntt"].append(row)
    print(json.dumps(row), file=sys.stderr, flush=True)
    ctx.free(d)
print(json.dumps(out))
