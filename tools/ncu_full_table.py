"""One table per kernel CLASS out of an `ncu --set full` report of a forward + backward (tools/one_step.py --mode fwdbwd):

    python tools/ncu_full_table.py gpurun_out/full.ncu-rep | gpurun_out/raw.csv (= `ncu -i REP --page raw --csv`) [hbm_gbs] > profiles/rN_ncu_all_kernels.md

per class (kernel name incl. template arguments): launches, total / mean duration, DRAM bytes read + written (sum over the launches),
achieved DRAM GB/s = bytes / duration against the measured copy bandwidth, ncu's DRAM-throughput %, tensor-pipe active %, L2 -> SM
bytes, registers per thread, achieved occupancy.  Durations under ncu are serialised, cold-cache and at ncu's clocks: compare
SHARES and RATIOS; byte counts are exact."""
import csv
import json
import os
import re
import subprocess
import sys
from collections import OrderedDict

M = {'dur': 'gpu__time_duration.sum', 'rd': 'dram__bytes_read.sum', 'wr': 'dram__bytes_write.sum',
     'dram_pct': 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'tensor': 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
     'l2sm': 'l1tex__m_xbar2l1tex_read_bytes.sum', 'regs': 'launch__registers_per_thread', 'occ': 'sm__warps_active.avg.pct_of_peak_sustained_active',
     'sm_pct': 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l2_pct': 'lts__throughput.avg.pct_of_peak_sustained_elapsed'}
SCALE = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6, 'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}


def short(name):
    m = re.match(r'(?:void )?([A-Za-z0-9_]+)(<[^(]*>)?\(', name)
    return (m.group(1) + (m.group(2) or '').replace(' ', '')) if m else name.split('(')[0]


def main():
    rep = sys.argv[1]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hbm = float(sys.argv[2]) if len(sys.argv) > 2 else json.load(open(os.path.join(root, 'MEASURED_PEAKS.json')))['hbm_gbs']
    txt = open(rep).read() if rep.endswith('.csv') else subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader([ln for ln in txt.splitlines() if ln.startswith('"')]))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {k: hdr.index(v) for k, v in M.items() if v in hdr}
    ni = hdr.index('Kernel Name')

    def val(r, k):
        if k not in col:
            return 0.0
        try:
            return float(r[col[k]].replace(',', '')) * SCALE.get(units[col[k]], 1.0)
        except ValueError:
            return 0.0
    agg = OrderedDict()
    for r in data:
        a = agg.setdefault(short(r[ni]), dict(n=0, us=0.0, rd=0.0, wr=0.0, l2sm=0.0, dram_pct=0.0, tensor=0.0, occ=0.0, regs=0, sm_pct=0.0, l2_pct=0.0))
        us = val(r, 'dur')
        a['n'] += 1; a['us'] += us; a['rd'] += val(r, 'rd'); a['wr'] += val(r, 'wr'); a['l2sm'] += val(r, 'l2sm')
        for k in ('dram_pct', 'tensor', 'occ', 'sm_pct', 'l2_pct'):
            a[k] += us * val(r, k)
        a['regs'] = max(a['regs'], int(val(r, 'regs')))
    tot = sum(a['us'] for a in agg.values())
    print(f'| kernel class | launches | total us | share | mean us | DRAM read MB | DRAM write MB | achieved GB/s | of {hbm:.0f} GB/s (measured copy) | ncu DRAM % | '
          'tensor pipe % | SM % | L2 % | L2->SM MB | regs | occupancy % |')
    print('|---|' + '---|' * 15)
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]['us']):
        gbs = (a['rd'] + a['wr']) / (a['us'] * 1e-6) / 1e9 if a['us'] else 0.0
        w = (lambda key: a[key] / a['us'] if a['us'] else 0.0)
        print(f"| `{k}` | {a['n']} | {a['us']:.0f} | {100 * a['us'] / tot:.1f}% | {a['us'] / a['n']:.1f} | {a['rd'] / 1e6:.0f} | {a['wr'] / 1e6:.0f} | {gbs:.0f} | "
              f"{gbs / hbm:.2f} | {w('dram_pct'):.0f} | {w('tensor'):.0f} | {w('sm_pct'):.0f} | {w('l2_pct'):.0f} | {a['l2sm'] / 1e6:.0f} | {a['regs']} | {w('occ'):.0f} |")
    rd, wr = sum(a['rd'] for a in agg.values()), sum(a['wr'] for a in agg.values())
    print(f"| **all {sum(a['n'] for a in agg.values())} launches** | | {tot:.0f} | | | {rd / 1e6:.0f} | {wr / 1e6:.0f} | {(rd + wr) / (tot * 1e-6) / 1e9:.0f} | | | | | | | | |")


if __name__ == '__main__':
    main()
