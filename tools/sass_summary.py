#!/usr/bin/env python
"""Per-kernel SASS evidence for libhgin.so: which kernels issue Blackwell tensor-core / TMA / TMEM instructions.

    python tools/sass_summary.py [path/to/libhgin.so] > profiles/rNN_sass_summary.txt

Counts, per kernel of the sm_100a image (cuobjdump -sass), the mnemonics that prove the data path:
UTCHMMA / UTCQMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / .st), UTMALDG / UTMASTG (TMA tensor load / store),
UBLKCP (cp.async.bulk), UTCBAR (tcgen05.commit), SYNCS (mbarrier), LDGSTS (cp.async), plus the instruction total.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "LDGSTS", "HMMA", "FFMA"]


def main():
    so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gnn_link_prediction_b200", "libhgin.so")
    out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    archs = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            while cur in kernels:            # the same instantiation emitted by two translation units: keep both
                cur += "'"
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P(?:\d+|T)\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            kernels[cur]["total"] += 1
            for k in KEYS:
                if op.startswith(k):
                    kernels[cur][k] += 1
    demangled = subprocess.run(["cu++filt"] + [k.rstrip("'") for k in kernels], capture_output=True, text=True).stdout.splitlines()
    print(f"# {os.path.relpath(so, ROOT)}: architectures {archs}; {len(kernels)} kernels")
    print("# " + " | ".join(["instr"] + KEYS + ["kernel"]))
    tot = collections.Counter()
    for (name, c), dm in zip(kernels.items(), demangled):
        tot.update(c)
        short = re.sub(r"\(.*", "", dm)
        print(" | ".join([f"{c['total']:6d}"] + [f"{c[k]:4d}" for k in KEYS] + [short[:110]]))
    print(" | ".join([f"{tot['total']:6d}"] + [f"{tot[k]:4d}" for k in KEYS] + ["TOTAL"]))


if __name__ == "__main__":
    main()
