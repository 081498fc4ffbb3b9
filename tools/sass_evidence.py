#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove the Blackwell paths (runs on the CPU container):
    python tools/sass_evidence.py > profiles/r2_sass_evidence.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vibravox_b200", "libvbx_b200.so")
WANT = ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMAPF", "UTCATOMSWS", "SYNCS", "HMMA", "REDG", "RED")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True,
                       text=True).stdout.split("\n")
print("SASS evidence (cuobjdump -sass vibravox_b200/libvbx_b200.so, sm_100a): tensor-core / TMEM / TMA / bulk-copy mnemonics per kernel")
print("UTCHMMA = tcgen05.mma kind::f16, UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st, UTMALDG = cp.async.bulk.tensor (TMA "
      "tensor map), UBLKCP = cp.async.bulk, SYNCS = mbarrier ops; HMMA (legacy mma.sync) must not appear\n")
blocks = sass.split("Function : ")[1:]
legacy = 0
for name, blk in zip(names, blocks):
    cnt = collections.Counter()
    for line in blk.split("\n"):
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            for w in WANT:
                if op == w or (w in ("UTCHMMA", "UTMALDG", "UBLKCP", "LDTM", "STTM", "HMMA") and op.startswith(w)):
                    cnt[w] += 1
    legacy += cnt.get("HMMA", 0)
    keep = {k: v for k, v in cnt.items() if k not in ("RED", "REDG")}
    if any(k in keep for k in ("UTCHMMA", "UTMALDG", "UBLKCP", "LDTM")):
        print(f"{name[:100]:<100} " + "  ".join(f"{k}={v}" for k, v in sorted(keep.items())))
print(f"\nlegacy HMMA instructions in the library: {legacy}")
