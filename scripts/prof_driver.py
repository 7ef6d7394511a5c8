"""Small driver for ncu captures: one MSM (2^LOG points) then one forward NTT (2^NTT_LOG), device-resident."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import plonk_prototype_b200 as pb

L = int(os.environ.get("LOG", "22"))
NL = int(os.environ.get("NTT_LOG", "24"))
ctx = pb.Context(0)
n = 1 << L
bases = ctx.malloc(n * 96)
ctx.synthetic_bases_dev(bases, n)
PRE = os.environ.get("PRECOMP")
srs = ctx.srs_wrap_dev(bases, n)
if PRE:
    ctx.srs_precompute(srs)
s = bench.random_fr_limbs(0xB2000000 + L, n)
sd = ctx.malloc(n * 32)
ctx.h2d(sd, s)
for _ in range(int(os.environ.get("REPS", "1"))):
    ctx.msm_dev(srs, sd, n)
x = bench.random_fr_limbs(0xF1F00000 + NL, 1 << NL)
d = ctx.malloc(32 << NL)
ctx.h2d(d, x)
for _ in range(int(os.environ.get("REPS", "1"))):
    ctx.ntt_dev(d, NL, 0, 0)
ctx.sync()
print("driver done; launches:", ctx.launch_count())
