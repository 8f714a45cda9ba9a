#!/usr/bin/env python
"""Per-kernel counts of the Blackwell-native SASS opcodes in libdsact.so (B200_PROFILING.md "What proves a
Blackwell-native kernel"): UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA), UTCBAR
(tcgen05.commit), FFMA2 / FMUL2 / FADD2 (packed fp32), plus the legacy tensor opcodes that must NOT appear.

    python tools/sass_summary.py > profiles/r2_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, "dsac-v2_b200", "libdsact.so")
WANT = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "FFMA2", "FMUL2", "FADD2", "MUFU",
        "HMMA", "HGMMA"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n
print(f"# cuobjdump -sass {os.path.relpath(lib, REPO)} (sm_100a): opcode counts per kernel; HMMA / HGMMA (legacy mma.sync / wgmma) must be 0")
print("# " + " ".join(f"{w:>8s}" for w in WANT) + "  total  kernel")
for block in out.split("Function : ")[1:]:
    name = block.split("\n", 1)[0].strip()
    ops = collections.Counter()
    for line in block.split("\n"):
        m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(.*?);", line)
        if m:
            ins = re.sub(r"^@!?U?P\w+\s+", "", m.group(1).strip())
            ops[ins.split()[0].split(".")[0]] += 1
    print("  " + " ".join(f"{ops.get(w, 0):8d}" for w in WANT) + f"  {sum(ops.values()):5d}  {demangle(name)[:110]}")
