"""Runs only bench.py's `prove` object (bench_prove) at a chosen size: python scripts/bench_prove_only.py [LOG_GATES]"""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import plonk_prototype_b200 as pb  # noqa: E402

ctx = pb.Context(0)
args = types.SimpleNamespace(prove_log_n=int(sys.argv[1]) if len(sys.argv) > 1 else 20, steps=3)
print(json.dumps(bench.bench_prove(ctx, None, args)))
