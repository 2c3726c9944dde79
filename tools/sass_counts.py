#!/usr/bin/env python
"""Per-kernel counts of the SASS opcodes that prove the Blackwell path (B200_PROFILING.md "What proves a Blackwell-native
kernel"): UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UBLKCP = TMA, UTCBAR = tcgen05.commit,
SYNCS = mbarrier, HMMA = legacy mma.sync (must be 0).  Usage: python tools/sass_counts.py > profiles/rN_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "speech-intent-recognizer_b200", "libsir_b200.so")
OPS = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "FFMA", "MUFU")

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
names = {}
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        counts[cur]["_total"] += 1
        for k in OPS:
            if op.startswith(k):
                counts[cur][k] += 1
demangled = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}: instruction counts per kernel (static, not executed counts)")
print("kernel | total | " + " | ".join(OPS))
for (mangled, c), name in zip(counts.items(), demangled):
    short = re.sub(r"\(.*\)$", "", name)
    print(f"{short} | {c['_total']} | " + " | ".join(str(c[k]) for k in OPS))
