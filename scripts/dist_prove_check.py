"""Sharded PLONK prove under torchrun (one process per GPU): every rank must return the same proof, and at sizes one GPU
can prove alone rank 0 checks it against the single-GPU proof byte for byte.
usage: torchrun --nproc-per-node G scripts/dist_prove_check.py LOG_GATES [LOG_GATES …]"""
import hashlib
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import plonk_prototype_b200 as pb  # noqa: E402
from plonk_prototype_b200.synth import synthetic_circuit_columns  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = pb.Context(local)
    gather = pb.torch_allgather(dist, dev)
    a2a_dev, ag_dev = pb.torch_device_collectives(dist, dev)
    replicate_round3 = bool(os.environ.get("PB200_REPLICATE_ROUND3"))
    use_comm = os.environ.get("PB200_DIST_MODE", "comm") == "comm"   # collectives by the library's own NCCL communicator
    if use_comm:
        ctx.comm_init_from_torch(dist)
    for L in [int(a) for a in sys.argv[1:]] or [16]:
        n, tau, label = 1 << L, 0xB2000000 + L, b"pb200-dist"
        sel, wires, values, pi_pos, pi_vals = synthetic_circuit_columns(n)
        sp = pb.ShardedParameters(n, tau, rank, world, ctx)
        if use_comm:
            pk, vk = ctx.preprocess_comm(sp.srs, sel, wires, values.shape[0], label)
        else:
            pk, vk = ctx.preprocess(sp.srs, sel, wires, values.shape[0], label, shard=(rank, world, gather) if replicate_round3 else (rank, world, gather, a2a_dev, ag_dev))
        proof = ctx.prove(sp.srs, pk, values, pi_pos, pi_vals)
        t = []
        for _ in range(3):
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            proof2 = ctx.prove(sp.srs, pk, values, pi_pos, pi_vals)
            dt = torch.tensor([time.perf_counter() - t0], device=dev)
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            t.append(1e3 * float(dt))
            assert proof2 == proof
        digest = torch.frombuffer(bytearray(hashlib.sha256(proof + vk).digest()), dtype=torch.uint8).to(dev)
        alld = torch.empty(world * 32, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(alld, digest)
        same = all(bytes(alld[32 * r:32 * r + 32].cpu().numpy()) == bytes(digest.cpu().numpy()) for r in range(world))
        ctx.prover_key_free(pk)
        sp.close()
        single_ms, matches = None, None
        if rank == 0 and L <= 24:
            pp = pb.PublicParameters(n - 1, tau, ctx)
            pk1, vk1 = ctx.preprocess(pp.srs, sel, wires, values.shape[0], label)
            p1 = ctx.prove(pp.srs, pk1, values, pi_pos, pi_vals)
            t0 = time.perf_counter()
            ctx.prove(pp.srs, pk1, values, pi_pos, pi_vals)
            single_ms = 1e3 * (time.perf_counter() - t0)
            matches = (p1 == proof) and (vk1 == vk)
            ctx.prover_key_free(pk1)
            pp.close()
        dist.barrier()
        if rank == 0:
            print(json.dumps({"log_gates": L, "world": world, "prove_ms": sum(t) / len(t), "min_ms": min(t),
                              "round3": "replicated" if replicate_round3 else "sharded", "collectives": "library NCCL (pb200_comm)" if use_comm else "host callbacks (torch.distributed)", "all_ranks_same_proof": same, "single_gpu_ms": single_ms, "equals_single_gpu_proof": matches,
                              "proof_sha256": hashlib.sha256(proof).hexdigest()[:16]}), flush=True)
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
