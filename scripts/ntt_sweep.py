"""NTT kernel / plan comparison on one GPU (run under gpurun): forward transform of 2^L device-resident scalars, CUDA events,
for every kernel configuration, each in its own process (the switches are read once per process).

    python scripts/ntt_sweep.py [--logs 16,18,20,22,24,26] [--out gpurun_out/ntt_sweep.json]

configurations:  legacy        PB200_NTT_PLAN=legacy            (round-1 plan: ≤ 2^11-point tiles, ntt_pass_kernel)
                 tma_shape_old PB200_NTT_KERNEL=legacy          (tma-shaped plan — 4-column tiles — on ntt_pass_kernel)
                 tma128        (default)                        ntt_pass_tma_kernel, 128 registers / 512 threads per SM
                 tma168        PB200_NTT_TMA_REGS=168           ntt_pass_tma_kernel, 168 registers / 384 threads per SM
Each line also carries a round-trip check and, for L ≤ 22, equality with the CPU oracle."""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CONFIGS = {"legacy": {"PB200_NTT_PLAN": "legacy"}, "tma_shape_old": {"PB200_NTT_KERNEL": "legacy"}, "tma128": {},
           "tma168": {"PB200_NTT_TMA_REGS": "168"}}


def worker(logs, batch4):
    import numpy as np
    import torch
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import plonk_prototype_b200 as pb
    import pyoracle as O
    ctx = pb.Context(0)
    stream = torch.cuda.ExternalStream(ctx.stream)
    out = []
    for L in logs:
        n = 1 << L
        x = O.random_fr(0xF1F00000 + L, n)
        d = ctx.malloc(x.nbytes)
        ctx.h2d(d, x)
        res = {"log_n": L}
        for name, inv, cos in (("fft", 0, 0), ("coset_fft", 0, 1)):
            ctx.h2d(d, x)
            ctx.ntt_dev(d, L, inv, cos)
            got = np.empty_like(x)
            ctx.d2h(got, d)
            if L <= 22:
                res[name + "_equals_oracle"] = bool((got == O.ntt(x, inv, cos, threads=os.cpu_count() or 8)).all())
            ctx.ntt_dev(d, L, 1, cos)
            ctx.d2h(got, d)
            res[name + "_roundtrip"] = bool((got == x).all())
            for _ in range(3):
                ctx.ntt_dev(d, L, inv, cos)
            ctx.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 20 if L <= 22 else 10
            e0.record(stream)
            for _ in range(reps):
                ctx.ntt_dev(d, L, inv, cos)
            e1.record(stream)
            ctx.sync()
            res[name + "_ms"] = e0.elapsed_time(e1) / reps
        ctx.free(d)
        if batch4 and L <= 22:   # the prover's shape: four vectors per launch
            d = ctx.malloc(4 * x.nbytes)
            for k in range(4):
                ctx.h2d(d + k * x.nbytes, x)
            for _ in range(3):
                ctx.ntt_batch_dev(d, L, 4, 1, 0)
            ctx.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(10):
                ctx.ntt_batch_dev(d, L, 4, 1, 0)
            e1.record(stream)
            ctx.sync()
            res["ifft_batch4_ms"] = e0.elapsed_time(e1) / 10
            ctx.free(d)
        out.append(res)
    ctx.close()
    print("RESULT " + json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--logs", default="16,18,20,22,24,26")
    ap.add_argument("--configs", default=",".join(CONFIGS))
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ntt_sweep.json"))
    ap.add_argument("--worker", default=None)
    args = ap.parse_args()
    logs = [int(v) for v in args.logs.split(",")]
    if args.worker:
        worker(logs, True)
        return
    results = {}
    for name in args.configs.split(","):
        env = dict(os.environ, **CONFIGS[name])
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", name, "--logs", args.logs], env=env, capture_output=True,
                           text=True, timeout=900)
        line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
        results[name] = json.loads(line[-1][7:]) if line else {"error": (r.stdout[-500:], r.stderr[-1500:]), "rc": r.returncode}
        print(name, json.dumps(results[name])[:600], flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(results, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
