#!/usr/bin/env python
"""Per-kernel SASS evidence (runs anywhere: no GPU needed).

    python tools/sass_opcodes.py > profiles/sass_opcodes_rN.txt

Disassembles the in-tree libb200np.so with cuobjdump and counts, per kernel, the mnemonics that prove the hardware
path (B200_PROFILING.md): UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor-map
copies, UBLKCP = cp.async.bulk, SYNCS = mbarrier, HMMA = legacy mma.sync (must be absent), FFMA = CUDA-core math.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "what-matters-for-meta-learning_b200", "csrc", "libb200np.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCOMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "SYNCS", "HMMA",
         "FFMA", "LDG", "STG", "LDS", "STS", "REDG", "ATOMG", "MUFU"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["total"] += 1
            for w in WATCH:
                if op.startswith(w):
                    cur[w] += 1
                    break
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"# SASS opcode counts per kernel of {os.path.relpath(SO, ROOT)} (cuobjdump -sass; sm_100a)")
    print("# " + " ".join(f"{w:>8s}" for w in ["total"] + WATCH) + "  kernel")
    for (name, c), dn in zip(kernels.items(), demangle):
        short = dn.replace("(anonymous namespace)::", "").replace("b200np::", "")
        short = re.sub(r"^void ", "", re.sub(r"\(.*", "", short))
        print("  " + " ".join(f"{c.get(w, 0):8d}" for w in ["total"] + WATCH) + "  " + short)
    return 0


if __name__ == "__main__":
    sys.exit(main())
