#!/usr/bin/env python
"""SASS opcode histograms of the kernels in libpb200.so (cuobjdump -sass), one block per kernel.

    python scripts/sass_hist.py [kernel-name-regex ...] > profiles/sass_rNN.txt

Per kernel: instruction count, the IMAD.WIDE share (the instruction the integer rooflines count), other
FMA-pipe integer instructions (plain IMAD / IMAD.MOV / IMAD.IADD …, which compete with IMAD.WIDE for the
same pipe), local-memory traffic (LDL / STL = spills), TMA / bulk-copy mnemonics (UTMALDG / UTMASTG / UBLKCP)
and the full opcode table.  Runs without a GPU.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "plonk-prototype_b200", "libpb200.so")


def demangle_short(name):
    m = re.search(r"\d+((?:ntt|msm|scan|quotient|poly|lincomb|perm|fr_|srs|witness|dist|gather|sigma|fill|pad|powers|scatter|imad|"
                  r"g1_|synthetic|block_|bucket|radix)[a-z0-9_]*)", name)
    return m.group(1) if m else name


def main():
    pats = [re.compile(p) for p in sys.argv[1:]] or [re.compile(".")]
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            kernels[cur][m.group(1).rstrip(";")] += 1
    for name, ops in kernels.items():
        short = demangle_short(name)
        if not any(p.search(short) for p in pats):
            continue
        total = sum(ops.values())
        wide = sum(v for k, v in ops.items() if k.startswith("IMAD.WIDE"))
        imad_other = sum(v for k, v in ops.items() if k.startswith("IMAD") and not k.startswith("IMAD.WIDE"))
        alu = sum(v for k, v in ops.items() if k.split(".")[0] in ("IADD3", "LOP3", "SHF", "PRMT", "SEL", "ISETP", "LEA", "IADD"))
        local = sum(v for k, v in ops.items() if k.split(".")[0] in ("LDL", "STL"))
        tma = {k: v for k, v in ops.items() if k.split(".")[0] in ("UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "SYNCS")}
        print("== %s  (%s)" % (short, name))
        print("   instructions %d | IMAD.WIDE %d (%.1f %%) | other IMAD %d | ALU-pipe int %d | LDL+STL %d | TMA/mbarrier %s"
              % (total, wide, 100.0 * wide / max(total, 1), imad_other, alu, local, tma or "none"))
        print("   " + "  ".join("%s:%d" % kv for kv in ops.most_common()))
        print()


if __name__ == "__main__":
    main()
