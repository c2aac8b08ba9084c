#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel shares."""
import collections
import csv
import sys


def main(path, out=None):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    cols, data = rows[hdr], rows[hdr + 1:]
    ik, iv = cols.index("Kernel Name"), cols.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        try:
            v = float(r[iv].replace(",", ""))
        except Exception:
            continue
        k = r[ik].replace("b200np::<unnamed>::", "").replace("<unnamed>::", "").replace("void ", "")[:90]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v for _, v in agg.values())
    lines = [f"# {path}", f"total {tot / 1e6:.3f} ms over {sum(c for c, _ in agg.values())} launches "
             "(ncu: serialised, cold cache -- compare SHARES, not absolutes)", "",
             f"{'ms':>9} {'share':>6} {'n':>5}  kernel"]
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{v / 1e6:9.3f} {100 * v / tot:5.1f}% {c:5d}  {k}")
    txt = "\n".join(lines) + "\n"
    if out:
        open(out, "w").write(txt)
    print(txt)


if __name__ == "__main__":
    main(*sys.argv[1:3])
