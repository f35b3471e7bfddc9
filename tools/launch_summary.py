"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (and optionally print every launch)."""
import collections
import csv
import re
import sys


def load(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    out = []
    for row in csv.DictReader(lines):
        name = re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', '')
        v = float(row['Metric Value'].replace(',', '')) * {'us': 1e-3, 'ns': 1e-6, 'ms': 1, 'usecond': 1e-3, 'nsecond': 1e-6, 'msecond': 1}[row['Metric Unit']]
        out.append((name, v, row.get('Grid Size', '')))
    return out


if __name__ == '__main__':
    rows = load(sys.argv[1])
    agg = collections.OrderedDict()
    tot = 0.0
    for n, v, _ in rows:
        a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
    print(f'Total {tot:.2f} ms over {len(rows)} launches\n\n| ms | share | launches | kernel |\n|---|---|---|---|')
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'| {t:.3f} | {100 * t / tot:.1f}% | {n} | `{k}` |')
    if len(sys.argv) > 2:
        for i, (n, v, g) in enumerate(rows):
            print(i, f'{v * 1e3:9.1f} us', g, n)
