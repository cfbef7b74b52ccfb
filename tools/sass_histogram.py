"""SASS opcode histogram per kernel of libdunet_b200.so (cuobjdump -sass): the evidence that the convolutions are tcgen05 / TMEM / TMA
kernels.  usage: python tools/sass_histogram.py > profiles/r2_sass_opcodes.txt   (CPU only; needs cuobjdump + c++filt)"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "diff-unet-amos_b200", "libdunet_b200.so")
KEYS = ["ACQBULK", "ATOM", "ELECT", "HMMA", "LDTM", "MEMBAR", "RED", "SYNCS", "UBLKCP", "UTCBAR", "UTCHMMA", "UTMALDG", "UTMASTG"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
kern, hist, count = None, {}, {}
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        hist[kern] = collections.Counter()
        count[kern] = 0
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        count[kern] += 1
        op = m.group(1)
        for k in KEYS:
            if op.startswith(k):
                hist[kern][k] += 1
names = subprocess.run(["c++filt"], input="\n".join(hist), capture_output=True, text=True).stdout.splitlines()
tot = collections.Counter()
for h in hist.values():
    tot.update(h)
print("# SASS opcode histogram per kernel of libdunet_b200.so (cuobjdump -sass, sm_100a), round-2 final build (tools/sass_histogram.py)")
print("# tcgen05.mma = UTCHMMA, tcgen05.commit = UTCBAR, tcgen05.ld = LDTM, TMA tensor load = UTMALDG, bulk copy = UBLKCP, mbarrier = SYNCS, warp mma.sync = HMMA")
print(f"# {len(hist)} kernels; totals: " + ", ".join(f"{k} {tot[k]}" for k in KEYS if tot[k]))
rows = []
for mangled, name in zip(hist, names):
    short = re.sub(r"^void dunet::", "", name)
    short = re.sub(r"\(.*$", "", short)
    rows.append((short, count[mangled], hist[mangled]))
for short, n, h in sorted(rows):
    print(f"{short:95s} instr {n:6d}  " + "  ".join(f"{k}:{h[k]}" for k in KEYS if h[k]))
