#!/usr/bin/env python
"""SASS opcode histogram per kernel of libseedvc_b200.so (cuobjdump, no GPU needed): the Blackwell-native evidence
(UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UTMAREDG = TMA load / store / reduce; HMMA
would be the legacy mma.sync path) kept reviewable without the binary.  usage: python scripts/sass_histogram.py > profiles/rN_sass_histogram.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "seed-vc_b200", "libseedvc_b200.so")
OPS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "HMMA", "MUFU", "FMNMX3", "FFMA2", "SYNCS"]

out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kern, hist, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern).replace("void ", "").replace("svc::", "")
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        op = m.group(1)
        hist[kern]["_n"] += 1
        for o in OPS:
            if op.startswith(o):
                hist[kern][o] += 1
                total[o] += 1
print("# SASS opcode histogram of `seed-vc_b200/libseedvc_b200.so` (sm_100a)\n")
print("Produced by `scripts/sass_histogram.py` (cuobjdump -sass).  `UTCHMMA` = tcgen05.mma kind::f16, `LDTM`/`STTM` = "
      "tcgen05.ld/st, `UTMALDG`/`UTMASTG`/`UTMAREDG` = TMA tensor load / store / reduce-add, `HMMA` = warp-level mma.sync: "
      "only `snake_mma_kernel` uses it, on purpose (banded-Toeplitz FIRs chained through the accumulator fragment, "
      "DESIGN section 5); every GEMM / attention kernel must show 0.\n")
print("Totals: " + ", ".join(f"{total[o]} `{o}`" for o in OPS) + f"; {len(hist)} kernels.\n")
print("| kernel | instr | " + " | ".join(OPS) + " |")
print("|---|---:|" + "---:|" * len(OPS))
for k, c in hist.items():
    if any(c[o] for o in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "HMMA")) or "--all" in sys.argv:
        print(f"| `{k}` | {c['_n']} | " + " | ".join(str(c[o]) for o in OPS) + " |")
print("\nKernels without tensor-core / TMA instructions (elementwise, FIR, SIMT fp32 paths): " +
      ", ".join(sorted({re.sub(r"<.*", "", k) for k, c in hist.items()
                        if not any(c[o] for o in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG"))})) + ".")
