"""Multi-process GPU tests (one process per GPU over NCCL, launched with torchrun from inside pytest): the sharded
four-step NTT against the single-GPU transform and the sharded PLONK prove against the single-GPU proof.  Skipped on
boxes with fewer than two GPUs (the 1-GPU round-end run); `gpurun --gpus 2 -- python -m pytest tests -m gpu` runs them."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _world():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    return 8 if n >= 8 else (4 if n >= 4 else 2)


def _torchrun(world, script, *args, env=None, timeout=900):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", "29617", os.path.join(ROOT, "scripts", script)] + [str(a) for a in args]
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, timeout=timeout, env=dict(os.environ, **(env or {})))
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    return [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]


def test_sharded_ntt_equals_single_gpu_transform():
    world = _world()
    for log_n in (20, 24):
        out = _torchrun(world, "dist_ntt_check.py", log_n)
        lib, fused, nccl = out[-3], out[-2], out[-1]
        assert lib["matches_single_gpu"] is True and lib["ifft_roundtrip"] is True and lib["sharded_msm_equals_single_gpu"] is True
        assert nccl["log_n"] == log_n and nccl["n_gpus"] == world
        assert nccl["all_shards_match_single_gpu"] is True and nccl["ifft_roundtrip"] is True
        assert fused["fused_matches_single_gpu"] is True


@pytest.mark.parametrize("mode", ["comm-fused", "comm-nccl", "callbacks-fused", "callbacks-nccl"])
def test_sharded_prove_equals_single_gpu_proof(mode):
    """comm: every collective by the library's own NCCL communicator (pb200_preprocess_comm), stream-ordered; callbacks: through the
    host's torch.distributed callbacks.  fused: round-3 forward exchange as peer stores; nccl: as an all-to-all."""
    world = _world()
    env = {"PB200_DIST_MODE": mode.split("-")[0]}
    if mode.endswith("nccl"):
        env["PB200_ROUND3_NCCL"] = "1"
    out = _torchrun(world, "dist_prove_check.py", 12, 16, 20, env=env)
    assert [o["log_gates"] for o in out] == [12, 16, 20]
    for o in out:
        assert o["world"] == world and o["all_ranks_same_proof"] is True and o["equals_single_gpu_proof"] is True, o
