#!/usr/bin/env python
"""bench.py — headline measurement for the MSM / NTT hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--log-n L]

A "step" is one G1 MSM of 2^L points (default L = 26: the largest size of BASELINE.json configs[1],
"standalone G1 MSM sweep 2^16–2^26") per GPU, with synthetic seeded inputs:
    bases   P_i = (a + i·d)·G generated on the GPU and kept resident (the SRS / CommitKey),
    scalars uniform Montgomery limbs < r from SplitMix64.
`value`  = points/s over all ranks with scalars already in HBM (CUDA events on the library's stream),
`e2e`    = the same through pb200_msm_g1 with the scalars in pinned HOST memory (H2D + result D2H timed);
           the bases are the resident SRS, exactly as CommitKey::powers_of_g is across dusk-plonk commits.
The same JSON line carries the NTT figures (2^24 forward transform, device-resident and host-to-host), the
integer roofline of the dominant kernel against a live IMAD.WIDE microbenchmark, the HBM reading for the NTT,
and the CPU baseline (the restated upstream algorithm in oracle/, all host cores, bounded sample).
Every full-size result is checked inside the run against the closed form Σ sᵢ·(a + i·d)·G.

N > 1 (torchrun): point-range sharding, one MSM of 2^L points per rank (weak scaling), partial results
all-gathered over NCCL and summed with pb200_g1_sum on rank 0.
`--impl reference` times the CPU restatement alone (the Rust reference cannot be built here).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
FR_MONT_R = (1 << 256) % R_MOD
A0, D0 = 0xB2000001, 0x9E3779B1
M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


# ------------------------------------------------------------------------------------------ inputs
def splitmix64_block(seed, start, count):
    """Outputs start … start+count−1 of SplitMix64(seed) — the state is a counter, so this vectorises."""
    with np.errstate(over="ignore"):
        k = np.arange(start + 1, start + count + 1, dtype=np.uint64)
        z = np.uint64(seed) + k * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def random_fr_limbs(seed, n, out=None):
    """n values uniform in [0, r) as (n, 4) uint64 limbs; identical stream to oracle/model random_fr."""
    r_limbs = [(R_MOD >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]
    res = out if out is not None else np.empty((n, 4), np.uint64)
    filled, cand = 0, 0
    chunk = 1 << 22
    while filled < n:
        raw = splitmix64_block(seed, 4 * cand, 4 * chunk).reshape(chunk, 4)
        cand += chunk
        raw[:, 3] &= np.uint64(0x7FFFFFFFFFFFFFFF)
        lt = np.zeros(chunk, bool)
        eq = np.ones(chunk, bool)
        for i in (3, 2, 1, 0):
            lt |= eq & (raw[:, i] < np.uint64(r_limbs[i]))
            eq &= raw[:, i] == np.uint64(r_limbs[i])
        ok = raw[lt]
        take = min(n - filled, ok.shape[0])
        res[filled:filled + take] = ok[:take]
        filled += take
    return res


def closed_form_scalar(scalars_mont, a, d):
    """k with Σ sᵢ·Pᵢ = k·G for Pᵢ = (a + i·d)·G (exact; 16-bit partial dot products)."""
    s = np.ascontiguousarray(scalars_mont, dtype=np.uint64).reshape(-1, 4)
    n = s.shape[0]
    k = np.uint64(a) + np.arange(n, dtype=np.uint64) * np.uint64(d)
    s16 = s.view(np.uint16).reshape(n, 16)
    k16 = k.view(np.uint16).reshape(n, 4).astype(np.uint64)
    total = 0
    for i in range(16):
        col = s16[:, i].astype(np.uint64)
        for j in range(4):
            total += int(np.dot(col, k16[:, j])) << (16 * (i + j))
    return total * pow(FR_MONT_R, -1, R_MOD) % R_MOD


# -------------------------------------------------------------------------------------- sharding
def shard_params(rank, n):
    """Point-range sharding: rank r owns global indices [r·n, (r+1)·n) of the synthetic SRS,
    i.e. bases (a_r + i·d)·G with a_r = a + r·n·d, and its own scalar stream."""
    return {"a": A0 + rank * n * D0, "d": D0, "scalar_seed": 0xB2000000 + (n.bit_length() - 1) + 1000 * rank}


def gather_partials(dist, local_u64x18, world, device):
    """All-gather the 144-byte partial results (as int64) — the only collective on the MSM path."""
    import torch
    mine = torch.from_numpy(np.ascontiguousarray(local_u64x18).view(np.int64).copy()).to(device)
    bufs = [torch.zeros(18, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(bufs, mine)
    return torch.stack(bufs).cpu().numpy().view(np.uint64)


# ---------------------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.FIELDS,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons, pw = [], [], set(), []
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        loaded = [c for c, p in zip(sm, pw) if p > 0.5 * max(pw)] if pw else sm
        return {"sm_mhz": statistics.median(loaded) if loaded else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------- reference arm
def cpu_msm_sample(sample_log, threads, seed, reps=1):
    """Time the restated upstream msm_variable_base (oracle/) on 2^sample_log points. → (pts/s, seconds)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as O  # CPU baseline leg: the one place bench.py may execute oracle/
    n = 1 << sample_log
    # bases: a stride of the same synthetic family (cost of an MSM does not depend on which points)
    pts = O.synthetic_bases(min(n, 1 << 12))
    pts = np.ascontiguousarray(np.tile(pts, (n // pts.shape[0] + 1, 1))[:n])
    s = random_fr_limbs(seed, n)
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        O.msm_variable_base(pts, s, threads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return n / best, best


def cpu_ntt_sample(log_n, threads, seed):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as O
    x = random_fr_limbs(seed, 1 << log_n)
    t0 = time.perf_counter()
    O.ntt(x, 0, 0, threads)
    dt = time.perf_counter() - t0
    return (1 << log_n) / dt, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample_log = min(args.log_n, args.cpu_sample_log)
    for _ in range(args.warmup and 1):
        cpu_msm_sample(min(sample_log, 14), cores, 1)
    t = []
    for k in range(args.steps):
        pps, dt = cpu_msm_sample(sample_log, cores, 0xB2000000 + args.log_n + k)
        t.append(dt)
    ms = 1e3 * sum(t) / len(t)
    value = (1 << sample_log) / (ms * 1e-3) / 1e6
    line = {
        "impl": "reference", "metric": "G1 MSM throughput (BLS12-381, 2^%d points per GPU)" % args.log_n,
        "value": value, "unit": "Mpts/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 limbs (Fp 12x32, Fr 8x32)",
        "data": "synthetic",
        "config": {"workload": "standalone G1 MSM, 2^%d points (BASELINE.json configs[1])" % args.log_n,
                   "points_per_gpu": 1 << args.log_n},
        "cpu_baseline": {"value": value, "unit": "Mpts/s", "cores": cores, "kind": "port",
                         "sample": "each step = msm_variable_base restatement (oracle/oracle.c, SURVEY App. B.1) on 2^%d "
                                   "points of the workload, %d threads over windows; the Rust reference cannot be built "
                                   "here (no cargo, crates not vendored)" % (sample_log, cores)},
        "e2e": {"value": value, "unit": "Mpts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if not args.skip_prove:
        # the CPU prover on a bounded circuit (2^16 gates: seconds), same restatement as the `prove.cpu_baseline` of our arm
        cl = min(args.prove_log_n, 16)
        t_pre, t_prove = cpu_prove_sample(cl, cores)
        line["prove"] = {"metric": "PLONK prove, synthetic arithmetic circuit, 2^%d gates, CPU restatement" % cl, "value": 1e3 * t_prove,
                         "unit": "ms", "preprocess_ms": 1e3 * t_pre, "cores": cores, "kind": "port"}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------- our arm
def run_ours(args, rank, world, local_rank):
    import torch
    import plonk_prototype_b200 as pb

    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    ctx = pb.Context(local_rank)  # raises without a B200: no fallback
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))
    L = args.log_n
    n = 1 << L
    shard = shard_params(rank, n)
    a_rank = shard["a"]

    # --- inputs (untimed): resident SRS shard, scalars in pinned host memory and in HBM
    bases = ctx.malloc(n * 96)
    ctx.synthetic_bases_dev(bases, n, a_rank, D0)
    srs = ctx.srs_wrap_dev(bases, n)
    pinned = torch.empty((n, 4), dtype=torch.int64, pin_memory=True)
    s_host = pinned.numpy().view(np.uint64)
    random_fr_limbs(shard["scalar_seed"], n, out=s_host)
    s_dev = ctx.malloc(n * 32)
    ctx.h2d(s_dev, s_host)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.sync()

    def step_device():
        out = ctx.msm_dev(srs, s_dev, n)
        return combine(out)

    def step_e2e():
        out = ctx.msm(srs, s_host)
        return combine(out)

    def combine(out):
        if dist is None:
            return out
        parts = gather_partials(dist, out, world, torch.device("cuda", local_rank))  # 144 B per rank over NVLink
        if rank == 0:
            return ctx.g1_sum(parts)
        return out

    # --- correctness of the full-size result (untimed): closed form over the global range
    res = step_device()
    k_local = closed_form_scalar(s_host, a_rank, D0)
    if dist is not None:
        ks = [None] * world
        dist.all_gather_object(ks, k_local)
        k_total = sum(ks) % R_MOD
    else:
        k_total = k_local
    verified = None
    if rank == 0:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import pyoracle as O  # checker only (cpu_baseline leg + result verification), never the measured path
        import model
        want = model.g1_mul(model.G1_GEN, k_total)
        got = O.g1_proj_to_affine_canonical(res)
        verified = bool(got == want)
        if not verified:
            raise SystemExit("MSM result does not match the closed form — refusing to report a number")

    # --- timed: device-resident
    for _ in range(args.warmup):
        step_device()
    ctx.profile_enable(True)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = ctx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    acc_ms, sort_ms, red_ms, part_ms = [], [], [], []
    ev0.record(stream)
    for _ in range(args.steps):
        step_device()
        acc_ms.append(ctx.profile_ms("msm.accumulate"))
        sort_ms.append(ctx.profile_ms("msm.sort"))
        red_ms.append(ctx.profile_ms("msm.reduce"))
        part_ms.append(ctx.profile_ms("msm.partials"))
    ev1.record(stream)
    barrier()
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    dev_ms = ev0.elapsed_time(ev1) / args.steps
    ctx.profile_enable(False)

    # --- timed: end to end through the host-buffer C ABI call
    for _ in range(min(args.warmup, 2)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps

    if dist is not None:
        t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms = t.tolist()

    # --- NTT and roofline context (rank 0, N = 1 only)
    extra = {}
    if rank == 0 and world == 1:
        imad_peak, _ = ctx.imad_peak()
        w_model = -(-256 // max(L - 4, 1))
        alg_imad = n * w_model * 10 * 300.0            # SURVEY §8d model in IMAD.WIDE units (600 lo+hi ops = 300 wide)
        acc = sum(acc_ms) / len(acc_ms)
        extra["roofline"] = {
            "bound": "imad", "kernel": "msm_accumulate_kernel",
            "achieved": alg_imad / (acc * 1e-3) / 1e12, "peak": imad_peak / 1e12, "unit": "T IMAD.WIDE.U32 lane-op/s",
            "frac": alg_imad / (acc * 1e-3) / imad_peak,
            # dram__bytes_read.sum + dram__bytes_write.sum of one launch at 2^26, ncu --set full (profiles/ncu_full_r01_summary.txt):
            # 161.2 + 5.8 GB against ~84 GB of entries + gathered bases — irrelevant next to the integer work
            "traffic": 167.0e9 if L == 26 else None,
            "peak_source": "pb200_imad_peak: IMAD.WIDE.U32.X carry-chain microbenchmark, this run (MEASURED_PEAKS.json has no "
                           "integer peak); see profiles/imad_explore_r01.txt",
            "model": "N*ceil(256/(log2N-4))*10 Fp mul * 300 IMAD.WIDE (SURVEY.md §8d: 600 lo+hi lane-ops at 64/clk/SM "
                     "== 300 IMAD.WIDE at the measured 32/clk/SM)",
            "kernel_ms": acc, "share_of_step": acc / dev_ms,
            "whole_msm_frac": alg_imad / (dev_ms * 1e-3) / imad_peak,
            "phases_ms": {"sort": sum(sort_ms) / len(sort_ms), "accumulate": acc,
                          "partials": sum(part_ms) / len(part_ms), "reduce+combine": sum(red_ms) / len(red_ms)},
        }
        extra["ntt"] = bench_ntt(ctx, stream, args, imad_peak)
        if not args.skip_prove:
            extra["prove"] = bench_prove(ctx, stream, args)
        cores = os.cpu_count() or 1
        sample_log = min(L, args.cpu_sample_log)
        pps, secs = cpu_msm_sample(sample_log, cores, 0xB2000000 + L)
        pps1, secs1 = cpu_msm_sample(min(L, 16), 1, 0xB2000000 + L)
        extra["cpu_baseline"] = {
            "value": pps / 1e6, "unit": "Mpts/s", "cores": cores, "kind": "port",
            "sample": "msm_variable_base restatement (oracle/oracle.c) on 2^%d points, %d threads: %.2f s; "
                      "single thread on 2^%d points: %.3f Mpts/s; Rust reference not buildable here"
                      % (sample_log, cores, secs, min(L, 16), pps1 / 1e6),
            "single_thread_mpts": pps1 / 1e6,
        }

    if world > 1:
        sharded = bench_ntt_sharded(ctx, dist, rank, world, args)
        if rank == 0:
            extra["ntt_sharded"] = sharded
        if not args.skip_prove:
            ps = bench_prove_sharded(ctx, dist, rank, world, local_rank, args)
            if rank == 0:
                extra["prove_sharded"] = ps

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        line = {
            "metric": "G1 MSM throughput (BLS12-381, 2^%d points per GPU)" % L,
            "value": world * n / (dev_ms * 1e-3) / 1e6, "unit": "Mpts/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32 limbs (Fp 12x32, Fr 8x32)", "data": "synthetic",
            "config": {"workload": "standalone G1 MSM, 2^%d points per GPU (BASELINE.json configs[1], largest size)" % L,
                       "points_per_gpu": n, "window_bits": int(pb._native.lib().pb200_msm_window_bits(n)),
                       "l2": "inputs larger than L2 (scalars %.1f GiB + bases %.1f GiB per GPU vs 126 MB)"
                             % (n * 32 / 2**30, n * 96 / 2**30),
                       "sharding": "point range per rank; 144 B partial results all-gathered over NCCL, summed on rank 0"
                                   if world > 1 else "single GPU"},
            "e2e": {"value": world * n / (e2e_ms * 1e-3) / 1e6, "unit": "Mpts/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": n * 32, "d2h_bytes_per_step": 144,
                    "note": "pb200_msm_g1: scalars from pinned host memory, result to host; bases are the resident SRS"},
            "gpu_launches": int(launches), "clocks": clocks, "result_verified": verified,
            "hbm_peak_gbs": peaks.get("hbm_gbs"),
        }
        line.update(extra)
        print(json.dumps(line), flush=True)
    ctx.srs_free(srs)
    ctx.free(bases)
    ctx.free(s_dev)
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


def bench_ntt_sharded(ctx, dist, rank, world, args):
    """Sharded four-step NTT of ONE 2^L vector over all ranks: column pass + NCCL all-to-all + batched
    rows.  Strong scaling by nature (the domain is fixed); verified by a sharded ifft(fft(x)) = x round trip here and
    against the single-GPU transform in scripts/dist_ntt_check.py / tests."""
    import torch
    import plonk_prototype_b200 as pb
    L = args.ntt_dist_log_n
    be = pb.GpuBackend(ctx, dist, torch)
    dom = pb.DistributedDomain(L, rank, world, be)
    spec = dom.spec
    shard = random_fr_limbs(0xF1F00000 + L + 1000 * rank, spec.local)
    buf = torch.from_numpy(shard.view(np.int64).reshape(-1).copy()).cuda()
    tmp = torch.empty_like(buf)
    torch.cuda.synchronize()   # `buf` was filled on torch's stream
    dom.fft(buf, tmp)
    dom.ifft(buf, tmp)
    ctx.sync(); torch.cuda.synchronize()
    ok = bool((buf.cpu().numpy().view(np.uint64).reshape(-1, 4) == shard).all())
    for _ in range(3):
        dom.fft(buf, tmp)
    dist.barrier(); ctx.sync(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(args.steps, 10)
    e0.record(be.stream)
    for _ in range(reps):
        dom.fft(buf, tmp)
    e1.record(be.stream)
    ctx.sync(); torch.cuda.synchronize(); dist.barrier()
    nccl_ms = e0.elapsed_time(e1) / reps
    # fused variant: the column kernel stores straight into the peers' row buffers (CUDA IPC over NVLink); wall clock
    # per transform including the cross-rank barrier, median of 10, input restored between repetitions
    peers = pb.PeerBuffers(ctx, dist, rank, world, spec.local)
    col = torch.from_numpy(shard.view(np.int64).reshape(-1).copy()).cuda()
    work = torch.empty_like(col)
    ts = []
    for i in range(13):
        work.copy_(col); torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        dom.fft_fused(work, peers)
        ctx.sync()
        if i >= 3:
            ts.append((time.perf_counter() - t0) * 1e3)
    fused_ms = sorted(ts)[len(ts) // 2]
    # the fused result must equal the NCCL result (itself checked against the single-GPU transform in tests)
    buf.copy_(col)
    torch.cuda.synchronize()   # the copy runs on torch's stream, the transform on the library's
    dom.fft(buf, tmp)
    ctx.sync(); torch.cuda.synchronize()
    fused_out = np.empty((spec.local, 4), np.uint64)
    ctx.d2h(fused_out, peers.mine.ptr)
    same = bool((fused_out == buf.cpu().numpy().view(np.uint64).reshape(-1, 4)).all())
    peers.close()
    t = torch.tensor([nccl_ms, fused_ms, 0.0 if (ok and same) else 1.0], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    nccl_ms, fused_ms, bad = t.tolist()
    ms = min(nccl_ms, fused_ms)
    return {"metric": "sharded four-step NTT, 2^%d points over %d GPUs (forward)" % (L, world), "ms": ms,
            "value": (1 << L) / (ms * 1e-3) / 1e6, "unit": "Melem/s", "scaling": "strong", "n1": spec.n1,
            "fused_peer_store_ms": fused_ms, "nccl_all_to_all_ms": nccl_ms,
            "exchange_bytes_per_gpu": spec.local * 32 * (world - 1) // world,
            "verified": bad == 0.0, "roundtrip_ok": ok, "fused_equals_nccl": same, "note": "fused: column kernel writes the owners' row buffers over NVLink peer memory, "
            "wall clock incl. the cross-rank barrier; nccl: column kernel + all_to_all_single + transpose kernel, CUDA events"}


def bench_ntt(ctx, stream, args, imad_peak):
    import torch
    L = args.ntt_log_n
    n = 1 << L
    pinned = torch.empty((n, 4), dtype=torch.int64, pin_memory=True)
    x = pinned.numpy().view(np.uint64)
    random_fr_limbs(0xF1F00000 + L, n, out=x)
    keep = x.copy()
    d = ctx.malloc(n * 32)
    ctx.h2d(d, x)
    # round trip check at full size (untimed)
    ctx.ntt_dev(d, L, 0, 0)
    ctx.ntt_dev(d, L, 1, 0)
    back = np.empty_like(keep)
    ctx.d2h(back, d)
    ok = bool((back == keep).all())
    for _ in range(max(args.warmup, 3)):
        ctx.ntt_dev(d, L, 0, 0)
    ctx.sync()
    import torch.cuda as tc
    e0, e1 = tc.Event(enable_timing=True), tc.Event(enable_timing=True)
    reps = max(args.steps, 10)
    l0 = ctx.launch_count()
    e0.record(stream)
    for _ in range(reps):
        ctx.ntt_dev(d, L, 0, 0)
    e1.record(stream)
    ctx.sync()
    ms = e0.elapsed_time(e1) / reps
    launches = (ctx.launch_count() - l0) // reps
    # host-to-host through pb200_ntt
    ctx.ntt(x, L, 0, 0)
    t0 = time.perf_counter()
    for _ in range(3):
        ctx.ntt(x, L, 0, 0)
    e2e_ms = (time.perf_counter() - t0) * 1e3 / 3
    ctx.free(d)
    passes = 1 if L <= 11 else (2 if L <= 22 else 3)
    alg_bytes = 64.0 * n * passes                      # 32 B read + 32 B write per element per pass
    alg_bytes_survey = 64.0 * n * (-(-L // 12))
    alg_imad = (n / 2) * L * 136.0                      # 272 lo/hi lane-ops = 136 IMAD.WIDE per Fr mul
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    cores = os.cpu_count() or 1
    cpu_log = min(L, 22)
    cpu_eps, cpu_s = cpu_ntt_sample(cpu_log, cores, 0xF1F00000 + L)
    return {
        "metric": "NTT throughput (BLS12-381 Fr, 2^%d, forward, device-resident)" % L,
        "value": n / (ms * 1e-3) / 1e6, "unit": "Melem/s", "ms": ms, "kernels_per_transform": int(launches),
        "roundtrip_verified": ok,
        "e2e": {"value": n / (e2e_ms * 1e-3) / 1e6, "unit": "Melem/s", "ms": e2e_ms, "h2d_bytes_per_step": n * 32,
                "d2h_bytes_per_step": n * 32},
        "roofline": {"bound": "hbm", "achieved": alg_bytes_survey / (ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                     "frac": alg_bytes_survey / (ms * 1e-3) / 1e9 / hbm,
                     # per launch of ntt_pass_kernel at 2^24 (one of three passes): 0.54 GB read + 0.64 GB written (ncu)
                     "traffic": 1.175e9 if L == 24 else None, "traffic_unit": "bytes per pass launch (algorithmic: 64 B * n)",
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                     "model": "64 B * n * ceil(log2 n / 12) (SURVEY.md §8d); this build moves 64 B * n * %d" % passes,
                     "moved_bytes_frac": alg_bytes / (ms * 1e-3) / 1e9 / hbm},
        "roofline_imad": {"bound": "imad", "achieved": alg_imad / (ms * 1e-3) / 1e12, "peak": imad_peak / 1e12,
                          "unit": "T IMAD.WIDE.U32 lane-op/s", "frac": alg_imad / (ms * 1e-3) / imad_peak,
                          "model": "(n/2)*log2(n) Fr mul * 136 IMAD.WIDE (SURVEY.md §8d) — the roofline that binds"},
        "cpu_baseline": {"value": cpu_eps / 1e6, "unit": "Melem/s", "cores": cores, "kind": "port",
                         "sample": "best_fft/parallel_fft restatement (oracle/oracle.c) on 2^%d, %d threads: %.2f s"
                                   % (cpu_log, cores, cpu_s)},
    }


def cpu_prove_sample(log_gates, threads):
    """Time the C restatement of the dusk-plonk prover (oracle/plonk_oracle.inc) on the synthetic circuit with
    2^log_gates gates.  → (preprocess s, prove s).  Bases: a tiled synthetic family (prover cost does not depend on
    which points the commit key holds; the proof is timed, not checked — parity is tests/test_prover_*.py's job)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as O  # CPU baseline leg
    from plonk_prototype_b200.synth import synthetic_circuit_columns
    n = 1 << log_gates
    sel, wires, values, pi_pos, pi_vals = synthetic_circuit_columns(n)
    pts = O.synthetic_bases(min(n, 1 << 12))
    srs = np.ascontiguousarray(np.tile(pts, (n // pts.shape[0] + 1, 1))[:n])
    _, _, t_pre, t_prove = O.plonk_prove(sel, wires, values, pi_pos, pi_vals, srs, b"pb200-bench", threads=threads)
    return t_pre, t_prove


def bench_prove(ctx, stream, args, with_cpu=True):
    """BASELINE.json configs[3]: full PLONK prove of the synthetic 2^20-gate arithmetic circuit on one B200, through
    pb200_prove (witness from host memory, 1040-byte proof back on the host) — rounds 1-5 with device-resident
    polynomials, 11 MSM + 10 NTT(n) + 8 coset NTT(4n) + the pointwise kernels, Merlin transcript on the host."""
    import plonk_prototype_b200 as pb
    from plonk_prototype_b200.synth import synthetic_circuit_columns
    L = args.prove_log_n
    n = 1 << L
    import torch
    sel, wires, values, pi_pos, pi_vals = synthetic_circuit_columns(n)
    pinned = torch.empty(values.shape, dtype=torch.int64, pin_memory=True)   # the witness comes from pinned host memory
    pinned.numpy().view(np.uint64)[:] = values
    values = pinned.numpy().view(np.uint64)
    pp = pb.PublicParameters(n - 1, 0xB2000000 + L, ctx)
    t0 = time.perf_counter()
    pk, vk = ctx.preprocess(pp.srs, sel, wires, values.shape[0], b"pb200-bench")
    ctx.sync()
    pre_ms = 1e3 * (time.perf_counter() - t0)
    proofs = set()
    for _ in range(3):
        proofs.add(ctx.prove(pp.srs, pk, values, pi_pos, pi_vals))
    launches0 = ctx.launch_count()
    t = []
    for _ in range(max(args.steps, 3)):
        ctx.sync()
        t0 = time.perf_counter()
        proofs.add(ctx.prove(pp.srs, pk, values, pi_pos, pi_vals))
        t.append(1e3 * (time.perf_counter() - t0))
    launches = (ctx.launch_count() - launches0) // len(t)
    ctx.profile_enable(True)
    ctx.prove(pp.srs, pk, values, pi_pos, pi_vals)
    rounds = {"round%d" % k: round(ctx.profile_ms("prove.round%d" % k), 3) for k in range(1, 6)}
    ctx.profile_enable(False)
    out = {"metric": "PLONK prove, synthetic arithmetic circuit, 2^%d gates, 1 GPU (BASELINE.json configs[3])" % L,
           "value": sum(t) / len(t), "unit": "ms", "higher_is_better": False, "min_ms": min(t), "steps": len(t),
           "preprocess_ms": pre_ms, "rounds_ms": rounds, "gpu_launches_per_prove": int(launches),
           "deterministic": len(proofs) == 1, "prover_key_gib": ctx.prover_key_bytes(pk) / 2**30,
           "e2e": {"value": sum(t) / len(t), "unit": "ms", "h2d_bytes_per_step": int(values.nbytes + pi_vals.nbytes + pi_pos.nbytes),
                   "d2h_bytes_per_step": 1040 + 11 * 144 + 17 * 32,
                   "note": "pb200_prove is the end-to-end call: witness values from pinned host memory, proof bytes on the host"},
           "parity": "proofs byte-identical with the CPU restatement up to 2^13 gates and accepted by the pairing verifier "
                     "up to 2^20 gates (tests/test_prover_gpu.py); parity unpinned against the Rust reference (not buildable here)"}
    ctx.prover_key_free(pk)
    pp.close()
    if with_cpu:
        cores = os.cpu_count() or 1
        # bounded sample: the largest circuit ≤ 2^L whose CPU prove is expected to stay under ~40 s on this host
        est = lambda lg: 4.8 * (1 << (lg - 16)) * 8.0 / min(cores, 20)  # noqa: E731  (2^16: 4.8 s on 8 threads)
        cl = L
        while cl > 12 and est(cl) > 40.0:
            cl -= 1
        t_pre, t_prove = cpu_prove_sample(cl, cores)
        gpu_same = None
        if cl != L:
            gpu_same = gpu_prove_ms(ctx, cl)
        # what /root/reference/Cargo.toml:19 actually builds (default-features = false: no rayon) is single-threaded:
        # timed at 2^14 gates (≈ 8 s), with the GPU prove of the same circuit beside it
        _, t_prove_1 = cpu_prove_sample(14, 1)
        gpu_14 = gpu_prove_ms(ctx, 14)
        out["cpu_baseline"] = {"value": 1e3 * t_prove, "unit": "ms", "cores": cores, "kind": "port", "log_gates": cl,
                               "preprocess_ms": 1e3 * t_pre, "gpu_ms_same_size": gpu_same,
                               "single_thread": {"log_gates": 14, "cpu_ms": 1e3 * t_prove_1, "gpu_ms": gpu_14,
                                                 "note": "1 thread = the reference's own feature set (no rayon)"},
                               "sample": "C restatement of dusk-plonk 0.8 prove_with_preprocessed (oracle/plonk_oracle.inc: same NTT/MSM "
                                         "call list, window-parallel MSM and parallel FFT/quotient loop as the rayon build) on the "
                                         "same synthetic circuit at 2^%d gates, %d threads; Rust reference not buildable here" % (cl, cores)}
    return out


def bench_prove_sharded(ctx, dist, rank, world, local_rank, args):
    """BASELINE.json configs[4] shape: the same prove with every commitment sharded by point range over the ranks
    (commit-key slice per GPU, 144-byte partial sums all-gathered over NCCL); NTTs and pointwise kernels replicated.
    Strong scaling: total work fixed.  Timed as the max over ranks."""
    import torch
    import plonk_prototype_b200 as pb
    from plonk_prototype_b200.synth import synthetic_circuit_columns
    dev = torch.device("cuda", local_rank)
    out = []
    for L in sorted(set([args.prove_log_n, args.prove_dist_log_n])):
        n = 1 << L
        sel, wires, values, pi_pos, pi_vals = synthetic_circuit_columns(n)
        sp = pb.ShardedParameters(n, 0xB2000000 + L, rank, world, ctx)
        pk, _ = ctx.preprocess(sp.srs, sel, wires, values.shape[0], b"pb200-bench", shard=(rank, world, pb.torch_allgather(dist, dev)) + pb.torch_device_collectives(dist, dev))
        proofs = {ctx.prove(sp.srs, pk, values, pi_pos, pi_vals) for _ in range(2)}
        t = []
        for _ in range(max(args.steps, 3)):
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            proofs.add(ctx.prove(sp.srs, pk, values, pi_pos, pi_vals))
            dt = torch.tensor([time.perf_counter() - t0], device=dev)
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            t.append(1e3 * float(dt))
        ctx.prover_key_free(pk)
        sp.close()
        out.append({"log_gates": L, "ms": sum(t) / len(t), "min_ms": min(t), "deterministic": len(proofs) == 1, "scaling": "strong"})
    return {"metric": "sharded PLONK prove, %d GPUs (MSMs by point range + all-gather of partial commitments; round 3 as sharded four-step transforms)" % world,
            "unit": "ms", "sizes": out,
            "parity": "byte-identical with the single-GPU proof (scripts/dist_prove_check.py, profiles/dist_prove_*)"}


def gpu_prove_ms(ctx, L):
    import plonk_prototype_b200 as pb
    from plonk_prototype_b200.synth import synthetic_circuit_columns
    n = 1 << L
    sel, wires, values, pi_pos, pi_vals = synthetic_circuit_columns(n)
    pp = pb.PublicParameters(n - 1, 0xB2000000 + L, ctx)
    pk, _ = ctx.preprocess(pp.srs, sel, wires, values.shape[0], b"pb200-bench")
    for _ in range(2):
        ctx.prove(pp.srs, pk, values, pi_pos, pi_vals)
    t = []
    for _ in range(3):
        t0 = time.perf_counter()
        ctx.prove(pp.srs, pk, values, pi_pos, pi_vals)
        t.append(1e3 * (time.perf_counter() - t0))
    ctx.prover_key_free(pk)
    pp.close()
    return sum(t) / len(t)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log-n", type=int, default=26, help="log2 of MSM points per GPU")
    ap.add_argument("--ntt-log-n", type=int, default=24)
    ap.add_argument("--ntt-dist-log-n", type=int, default=26, help="log2 of the sharded NTT domain (N > 1 only)")
    ap.add_argument("--cpu-sample-log", type=int, default=20, help="log2 of the CPU baseline's bounded sample")
    ap.add_argument("--skip-prove", action="store_true")
    ap.add_argument("--prove-log-n", type=int, default=20, help="log2 of the gate count of the timed PLONK prove")
    ap.add_argument("--prove-dist-log-n", type=int, default=22, help="log2 of the gate count of the second sharded prove (N > 1)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
