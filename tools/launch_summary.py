"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel (last two sweeps)."""
import collections, csv, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith('==')]
rows = list(csv.DictReader(lines))
names = [re.sub(r'\(.*', '', x['Kernel Name'])[:60] for x in rows]
vals = [float(x['Metric Value'].replace(',', '')) for x in rows]
idx = [i for i, n in enumerate(names) if 'partial_gemm' in n and (', 0>' in n or '<0>' in n)]   # pass A of each sweep (DMMA or INT8 kernel)
start = idx[-2] - 2 if len(idx) >= 2 else 0
tail = list(zip(names, vals))[start:]
tot = collections.Counter(); cnt = collections.Counter()
for n, v in tail:
    tot[n] += v; cnt[n] += 1
s = sum(tot.values())
print(f"{'kernel':60s} {'calls/sweep':>11s} {'us/sweep':>10s} {'share':>7s}")
for n, v in tot.most_common():
    print(f"{n:60s} {cnt[n] / 2:11.1f} {v / 2e3:10.1f} {100 * v / s:6.1f}%")
print(f"{'total':60s} {'':11s} {s / 2e3:10.1f}")
