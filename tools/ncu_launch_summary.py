"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py into a markdown table:
per-kernel totals over the whole run and the per-launch times of the LAST solve pass (one u_solve)."""
import collections
import csv
import re
import sys

path, nlast = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 42      # launches of one u_solve at n = 4 (+ the top-level u_hat)
lines = [l for l in open(path) if not l.startswith("==")]
items = [it for it in csv.DictReader(lines) if it["Metric Name"] == "gpu__time_duration.sum"]
rows = [(it["Kernel Name"], float(it["Metric Value"].replace(",", "")), it["Grid Size"]) for it in items]


def short(n):
    n = re.sub(r"scasml::|\(anonymous namespace\)::|<unnamed>::|void ", "", n)
    return re.sub(r"\(.*", "", n)


tot = collections.defaultdict(lambda: [0, 0.0])
for n, v, g in rows:
    tot[short(n)][0] += 1
    tot[short(n)][1] += v
allns = sum(v for _, v, _ in rows)
print("| kernel | launches | total ms | share of run |\n|---|---|---|---|")
for k, (c, v) in sorted(tot.items(), key=lambda x: -x[1][1]):
    print(f"| `{k}` | {c} | {v/1e6:.3f} | {100*v/allns:.1f}% |")
pic = [(n, v, g) for n, v, g in rows if re.search("sample_|eval_tc|eval_f64|reduce_|point_weights|row_setup|mlp_terminal", n)]
last = pic[-nlast:]
t = sum(v for _, v, _ in last)
print(f"\nLast solve pass ({len(last)} launches, {t/1e6:.3f} ms of kernel time):\n")
print("| # | kernel | grid | ms | share |\n|---|---|---|---|---|")
for i, (n, v, g) in enumerate(last):
    print(f"| {i} | `{short(n)}` | {g} | {v/1e6:.3f} | {100*v/t:.1f}% |")
agg = collections.defaultdict(float)
for n, v, g in last:
    agg[short(n)] += v
print("\n| kernel | ms in the pass | share |\n|---|---|---|")
for k, v in sorted(agg.items(), key=lambda x: -x[1]):
    print(f"| `{k}` | {v/1e6:.3f} | {100*v/t:.1f}% |")
