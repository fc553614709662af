"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals for ONE training step
(the launches between the last two adam_kernel launches).  usage: summarize_launches.py launches.csv [out.txt]"""
import collections
import csv
import re
import sys


def main(path, out=None):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [r["Kernel Name"] for r in rows]
    idx = [i for i, n in enumerate(names) if "adam" in n]
    s, e = (idx[-2] + 1, idx[-1] + 1) if len(idx) >= 2 else (0, len(rows))
    tot = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[s:e]:
        n = re.sub(r"\(.*", "", r["Kernel Name"])
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        v = v / 1e3 if u.startswith("n") else (v * 1e3 if u.startswith("m") else v)
        tot[n][0] += 1
        tot[n][1] += v
    T = sum(v[1] for v in tot.values())
    text = [f"one training step: {e - s} launches, {T:.1f} us summed kernel time (ncu, serialised, cold-cache: compare SHARES)"]
    for n, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        text.append(f"{t:10.1f} us {100 * t / T:5.1f}%  x{c:4d}  {n[:120]}")
    text = "\n".join(text)
    print(text)
    if out:
        open(out, "w").write(text + "\n")


if __name__ == "__main__":
    main(*sys.argv[1:3])
