"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file X.csv`) into
time per kernel and share of the total.  Usage: python tools/launch_list_summary.py X.csv "<command>" """
import csv
import io
import re
import sys
from collections import OrderedDict


def main(path, cmd):
    txt = open(path).read()
    rows = list(csv.reader(io.StringIO(txt[txt.index('"ID"'):])))
    h = rows[0]
    idx = {k: i for i, k in enumerate(h)}
    per = OrderedDict()
    n = 0
    for r in rows[1:]:
        if len(r) < len(h) or r[idx["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]])
        name = re.sub(r"void |hgin::|\(anonymous namespace\)::|<unnamed>::|tcgemm::|thin::|_GLOBAL__N__\w+::", "", name)
        v = float(r[idx["Metric Value"]].replace(",", ""))
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r[idx["Metric Unit"]], 1e-6)
        t, c = per.get(name, (0.0, 0))
        per[name] = (t + v * scale, c + 1)
        n += 1
    total = sum(t for t, _ in per.values())
    print(f"# ncu launch list summary: {cmd}, {n} launches after warm-up")
    print(f"# cold-cache serialised times; compare SHARES, not absolutes. total {total:.3f} ms")
    print("ms | launches | share | kernel")
    for name, (t, c) in sorted(per.items(), key=lambda kv: -kv[1][0]):
        print(f"{t:8.3f} | {c:4d} | {100 * t / total:5.1f}% | {name[:100]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
