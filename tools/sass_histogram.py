#!/usr/bin/env python
"""SASS opcode histogram per kernel of libblr_cuda.so (cuobjdump needs no GPU): what each kernel is actually made of --
fp64 tensor-core MMAs (DMMA.8x8x4), TMA bulk copies (UBLKCP), mbarrier ops (SYNCS.*), shared-memory loads by width, the
fp64 vector instructions that share the pipe with DMMA (DFMA / DMUL / DADD), barriers, register hand-over (USETMAXREG).
Output: profiles/r02/sass_histogram.txt (committed per round)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "bayesianlinearregressors.jl_b200", "csrc", "libblr_cuda.so")
KEYS = ["DMMA", "UBLKCP", "UTMALDG", "UTMASTG", "UTCMMA", "LDTM", "SYNCS", "LDS.64", "LDS.128", "LDS", "STS", "LDG", "STG", "DFMA", "DMUL",
        "DADD", "MUFU", "BAR", "WARPSYNC", "SHFL", "USETMAXREG", "ATOM", "RED", "MEMBAR", "CCTL", "BRA"]


def main(out_path):
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            kernels[cur][op] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    with open(out_path, "w") as f:
        f.write("SASS opcode histogram of libblr_cuda.so (sm_100a), per kernel: total instructions, then selected opcode families\n")
        f.write("(prefix match: e.g. SYNCS counts SYNCS.ARRIVE.TRANS64 / SYNCS.PHASECHK...; LDS counts every shared load incl. .64/.128)\n\n")
        for (name, cnt), dn in zip(kernels.items(), demangle):
            short = re.sub(r"\(.*", "", dn)
            row = []
            for k in KEYS:
                v = sum(c for op, c in cnt.items() if op != "_total" and (op == k or op.startswith(k + ".") or (k in ("LDS", "DMMA", "SYNCS") and op.startswith(k))))
                if k in ("LDS.64", "LDS.128"):
                    v = sum(c for op, c in cnt.items() if op.startswith("LDS") and op.endswith(k[3:]))
                if v:
                    row.append(f"{k}={v}")
            f.write(f"{short}\n    total={cnt['_total']}  " + "  ".join(row) + "\n")
    print(open(out_path).read()[:3000])


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02", "sass_histogram.txt"))
