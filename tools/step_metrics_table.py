"""Per-kernel-class table out of an ncu --metrics CSV of ONE train step (tools/one_step.py):

    python tools/step_metrics_table.py gpurun_out/step_metrics.csv [hbm_gbs]

columns: launches, total us, share, DRAM MB read / written, achieved DRAM GB/s (bytes / duration), fraction of the measured copy
bandwidth, tensor-pipe active % (time-weighted), L2 traffic MB.  ncu serialises the kernels and measures each cold: compare
SHARES; the byte counts are exact."""
import csv
import json
import os
import re
import sys
from collections import OrderedDict, defaultdict


def load(path):
    rows = defaultdict(dict)
    names = {}
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        i = int(r['ID'])
        names[i] = r['Kernel Name']
        v = r['Metric Value'].replace(',', '')
        try:
            v = float(v)
        except ValueError:
            continue
        unit = r['Metric Unit']
        scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1.0)
        rows[i][r['Metric Name']] = v * scale
    return names, rows


def short(name):
    m = re.match(r'(?:void )?([A-Za-z0-9_]+)(<[^(]*>)?\(', name)
    base = m.group(1) if m else name.split('(')[0]
    tmpl = (m.group(2) or '') if m else ''
    return base + tmpl.replace(' ', '')


def main():
    path = sys.argv[1]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hbm = float(sys.argv[2]) if len(sys.argv) > 2 else json.load(open(os.path.join(root, 'MEASURED_PEAKS.json')))['hbm_gbs']
    names, rows = load(path)
    agg = OrderedDict()
    for i in sorted(rows):
        k = short(names[i])
        a = agg.setdefault(k, dict(n=0, us=0.0, rd=0.0, wr=0.0, l2=0.0, tens=0.0))
        m = rows[i]
        us = m.get('gpu__time_duration.sum', 0.0)
        a['n'] += 1; a['us'] += us; a['rd'] += m.get('dram__bytes_read.sum', 0.0); a['wr'] += m.get('dram__bytes_write.sum', 0.0)
        a['l2'] += m.get('lts__t_bytes.sum', 0.0)
        a['tens'] += us * m.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0.0)
    tot = sum(a['us'] for a in agg.values())
    print(f'| kernel | launches | us | share | DRAM read MB | DRAM write MB | DRAM GB/s | of {hbm:.0f} GB/s | tensor pipe % | L2 MB |')
    print('|---|---|---|---|---|---|---|---|---|---|')
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]['us']):
        gbs = (a['rd'] + a['wr']) / (a['us'] * 1e-6) / 1e9 if a['us'] else 0.0
        print(f"| `{k}` | {a['n']} | {a['us']:.0f} | {100 * a['us'] / tot:.1f}% | {a['rd'] / 1e6:.0f} | {a['wr'] / 1e6:.0f} | {gbs:.0f} | {gbs / hbm:.2f} | "
              f"{a['tens'] / a['us'] if a['us'] else 0:.0f} | {a['l2'] / 1e6:.0f} |")
    rd, wr = sum(a['rd'] for a in agg.values()), sum(a['wr'] for a in agg.values())
    print(f"| **all {sum(a['n'] for a in agg.values())} launches** | | {tot:.0f} | | {rd / 1e6:.0f} | {wr / 1e6:.0f} | {(rd + wr) / (tot * 1e-6) / 1e9:.0f} | | | |")


if __name__ == '__main__':
    main()
