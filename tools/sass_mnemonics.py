#!/usr/bin/env python
"""Static counts of the tensor-core / TMEM / TMA / async-copy SASS mnemonics per kernel of the built library (runs here, no GPU):
cuobjdump -sass libavsep_b200.so.  usage: sass_mnemonics.py > profiles/r2_sass_mnemonics.txt"""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "audio-visual-sepatation-in-visual-agnostic-situtation_b200", "libavsep_b200.so")
COLS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "LDGSTS", "HMMA", "DFMA", "MUFU", "SHFL"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
blocks = re.split(r"(?m)^\s*Function : \S+\n", sass)[1:]
print("# Tensor-core / TMEM / TMA / async-copy SASS mnemonics per kernel of libavsep_b200.so (round 2, final kernels)")
print("# produced here (no GPU) with tools/sass_mnemonics.py: cuobjdump -sass libavsep_b200.so ; counts of static instructions")
print("# UTCHMMA = tcgen05.mma (kind::f16), .2CTA = cta_group::2; LDTM = tcgen05.ld; UTMALDG / UTMASTG = cp.async.bulk.tensor load / store (TMA);")
print("# UBLKCP = cp.async.bulk (1-D TMA bulk copy); SYNCS = mbarrier ops; LDGSTS = cp.async; HMMA would be the legacy mma.sync path (none)\n")
print(f"{'kernel':48s}" + "".join(f"{c:>13s}" for c in COLS))
for name, body in zip(names, blocks):
    short = re.sub(r"^void ", "", name).replace("avsep::", "")
    short = re.sub(r"\(.*$", "", short)
    counts = []
    for c in COLS:
        if c == "UTCHMMA.2CTA":
            counts.append(len(re.findall(r"\bUTCHMMA\.2CTA\b", body)))
        else:
            counts.append(len(re.findall(rf"\b{re.escape(c)}\b", body)))
    print(f"{short[:47]:48s}" + "".join(f"{n:13d}" for n in counts))
