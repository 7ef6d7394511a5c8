"""torchrun script: sharded four-step NTT over all visible GPUs (NCCL all-to-all), verified against the single-GPU
transform and timed.   torchrun --nproc-per-node G scripts/dist_ntt_check.py [log_n]"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import plonk_prototype_b200 as pb

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = pb.Context(local)
be = pb.GpuBackend(ctx, dist, torch)
dom = pb.DistributedDomain(log_n, rank, world, be)
spec = dom.spec
x = bench.random_fr_limbs(0xF1F00000 + log_n, 1 << log_n)          # every rank derives the same global vector
buf = torch.from_numpy(spec.scatter(x, rank, "column").view(np.int64).reshape(-1).copy()).cuda()
tmp = torch.empty_like(buf)
dom.fft(buf, tmp)
ctx.sync(); torch.cuda.synchronize()
mine = buf.cpu().numpy().view(np.uint64).reshape(-1, 4)
ok = True
if rank == 0:
    ref = x.copy()
    d = ctx.malloc(ref.nbytes)
    ctx.h2d(d, ref)
    ctx.ntt_dev(d, log_n, False, False)
    ctx.d2h(ref, d)
    ctx.free(d)
    ok = bool((spec.scatter(ref, 0, "row") == mine).all())
    ref_all = ref
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.broadcast(flag, 0)
# every rank checks its own shard against rank 0's reference result
ref_t = torch.from_numpy((ref_all if rank == 0 else np.empty((1 << log_n, 4), np.uint64)).view(np.int64)).cuda()
dist.broadcast(ref_t, 0)
ok_local = bool((spec.scatter(ref_t.cpu().numpy().view(np.uint64), rank, "row") == mine).all())
dom.ifft(buf, tmp)
ctx.sync(); torch.cuda.synchronize()
back_ok = bool((buf.cpu().numpy().view(np.uint64).reshape(-1, 4) == spec.scatter(x, rank, "column")).all())
oks = [None] * world
dist.all_gather_object(oks, (ok_local, back_ok))
# timing: max over ranks, CUDA events on the library stream
stream = be.stream
for _ in range(3):
    dom.fft(buf, tmp)
dist.barrier(); ctx.sync(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
e0.record(stream)
for _ in range(reps):
    dom.fft(buf, tmp)
e1.record(stream)
ctx.sync(); torch.cuda.synchronize(); dist.barrier()
ms = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda", dtype=torch.float64)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
# ---- fused variant: the column kernel stores straight into the peers' row buffers (CUDA IPC over NVLink)
peers = pb.PeerBuffers(ctx, dist, rank, world, spec.local)
col = torch.from_numpy(spec.scatter(x, rank, "column").view(np.int64).reshape(-1).copy()).cuda()
work = torch.empty_like(col)
work.copy_(col)
torch.cuda.synchronize(); dist.barrier()
dom.fft_fused(work, peers)
ctx.sync()
fused_out = np.empty((spec.local, 4), np.uint64)
ctx.d2h(fused_out, peers.mine.ptr)
fused_ok = bool((spec.scatter(ref_t.cpu().numpy().view(np.uint64), rank, "row") == fused_out).all())
foks = [None] * world
dist.all_gather_object(foks, fused_ok)
for _ in range(3):
    work.copy_(col); torch.cuda.synchronize(); dist.barrier()
    dom.fft_fused(work, peers)
import time
ts = []
for _ in range(10):
    work.copy_(col); torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    dom.fft_fused(work, peers)
    ctx.sync()
    ts.append((time.perf_counter() - t0) * 1e3)
fms = torch.tensor([sorted(ts)[len(ts) // 2]], device="cuda", dtype=torch.float64)
dist.all_reduce(fms, op=dist.ReduceOp.MAX)
# same wall-clock protocol for the NCCL variant, for a like-for-like comparison
ts = []
for _ in range(10):
    buf.copy_(col); torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    dom.fft(buf, tmp)
    ctx.sync(); torch.cuda.synchronize()
    ts.append((time.perf_counter() - t0) * 1e3)
nms = torch.tensor([sorted(ts)[len(ts) // 2]], device="cuda", dtype=torch.float64)
dist.all_reduce(nms, op=dist.ReduceOp.MAX)
peers.close()
# ---- the same transform driven entirely by the library: pb200_comm + pb200_ntt_sharded_dev (NCCL inside libpb200.so)
ctx.comm_init_from_torch(dist)
d_data, d_tmp = ctx.malloc(spec.local * 32), ctx.malloc(spec.local * 32)
ctx.h2d(d_data, spec.scatter(x, rank, "column"))
ctx.ntt_sharded_dev(d_data, d_tmp, log_n, False)
lib_out = np.empty((spec.local, 4), np.uint64)
ctx.d2h(lib_out, d_data)
lib_fwd_ok = bool((spec.scatter(ref_t.cpu().numpy().view(np.uint64), rank, "row") == lib_out).all())
ctx.ntt_sharded_dev(d_data, d_tmp, log_n, True)
ctx.d2h(lib_out, d_data)
lib_back_ok = bool((lib_out == spec.scatter(x, rank, "column")).all())
for _ in range(3):
    ctx.ntt_sharded_dev(d_data, d_tmp, log_n, False)
dist.barrier(); ctx.sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(10):
    ctx.ntt_sharded_dev(d_data, d_tmp, log_n, False)
e1.record(stream)
ctx.sync()
lms = torch.tensor([e0.elapsed_time(e1) / 10], device="cuda", dtype=torch.float64)
dist.all_reduce(lms, op=dist.ReduceOp.MAX)
# sharded MSM through the same communicator: Σ over ranks of slice MSMs equals the single-GPU MSM over the whole range
n_msm = 1 << 16
per = n_msm // world
bases = ctx.malloc(per * 96)
ctx.synthetic_bases_dev(bases, per, bench.A0 + rank * per * bench.D0, bench.D0)
srs = ctx.srs_wrap_dev(bases, per)
sc = bench.random_fr_limbs(0xABC, n_msm)
d_sc = ctx.malloc(per * 32)
ctx.h2d(d_sc, np.ascontiguousarray(sc[rank * per:(rank + 1) * per]))
total = ctx.msm_sharded_dev(srs, d_sc, per)
msm_ok = True
if rank == 0:
    allb = ctx.malloc(n_msm * 96)
    ctx.synthetic_bases_dev(allb, n_msm, bench.A0, bench.D0)
    srs1 = ctx.srs_wrap_dev(allb, n_msm)
    want = ctx.msm(srs1, sc)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import pyoracle as O
    msm_ok = O.g1_proj_to_affine_canonical(total) == O.g1_proj_to_affine_canonical(want)
    ctx.srs_free(srs1); ctx.free(allb)
loks = [None] * world
dist.all_gather_object(loks, (lib_fwd_ok, lib_back_ok, msm_ok))
ctx.srs_free(srs); ctx.free(bases); ctx.free(d_sc); ctx.free(d_data); ctx.free(d_tmp)
if rank == 0:
    print(json.dumps({"what": "library-driven sharded NTT (pb200_ntt_sharded_dev) and sharded MSM (pb200_msm_g1_sharded_dev) over pb200_comm",
                      "log_n": log_n, "n_gpus": world, "ms": lms.item(), "matches_single_gpu": all(o[0] for o in loks),
                      "ifft_roundtrip": all(o[1] for o in loks), "sharded_msm_equals_single_gpu": all(o[2] for o in loks)}))
if rank == 0:
    print(json.dumps({"what": "sharded NTT, exchange fused into the column kernel (peer stores) vs NCCL all-to-all + transpose",
                      "log_n": log_n, "n_gpus": world, "fused_wall_ms": fms.item(), "nccl_wall_ms": nms.item(),
                      "fused_matches_single_gpu": all(foks)}))
    print(json.dumps({"what": "sharded four-step NTT, forward, column layout -> row layout", "log_n": log_n, "n_gpus": world,
                      "n1": spec.n1, "ms": ms.item(), "melem_per_s": (1 << log_n) / ms.item() / 1e3,
                      "all_shards_match_single_gpu": all(o[0] for o in oks), "ifft_roundtrip": all(o[1] for o in oks)}))
dist.destroy_process_group()
