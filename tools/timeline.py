"""Kernel timeline of ONE train step in its real (multi-lane, graphed) schedule, from CUPTI through torch.profiler:

    python tools/timeline.py [--batch 128] [--graph 1] --out gpurun_out/timeline.csv

Writes one row per kernel (name, stream, start us, duration us) and prints where the step's time goes: the union of the
intervals in which a tensor-core kernel is running, the rest ("tensor-idle" time), and which kernels fill the idle gaps.
This is the measurement ncu's serialised launch list cannot give: how much of the non-GEMM work the lanes really hide."""
import argparse
import collections
import csv
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from spatial_intention_maps_b200 import networks, synth, train as T  # noqa: E402

TENSOR = re.compile(r'(conv\w*_umma_kernel|wgrad\w*_umma_kernel)')


def short(name):
    return re.sub(r'\(.*', '', name).replace('void ', '')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=128)
    ap.add_argument('--warm', type=int, default=6)
    ap.add_argument('--out', default='gpurun_out/timeline.csv')
    a = ap.parse_args()
    dev = torch.device('cuda', 0)
    B = a.batch
    pol, tgt, opt = bench.make_nets(networks, torch, dev, bench.C_IN, bench.A_OUT, B)
    hb = T.HostBatch(B, bench.C_IN).fill(synth.synth_batch(B, bench.C_IN, bench.A_OUT, 1234, terminal_every=bench.TERMINAL_EVERY))
    db = T.DeviceBatch(B, bench.C_IN, dev).upload(hb)

    def run():
        T.train_step_device(pol, tgt, opt, db, B, bench.GAMMA, 100, True)
    for _ in range(a.warm):
        run()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):                          # three back-to-back steps: the middle one is in steady state
            run()
        torch.cuda.synchronize()
    ev = []
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start:
            ev.append((e.time_range.start, e.time_range.end, short(e.name), getattr(e, 'device_index', 0)))
    ev = [x for x in ev if 'memcpy' not in x[2].lower() and 'memset' not in x[2].lower()]
    ev.sort()
    if not ev:
        print('no CUDA kernels recorded')
        return
    # split into the three steps at the sgd_update kernel (last kernel of a step)
    ends = [i for i, x in enumerate(ev) if 'sgd_update' in x[2]]
    if len(ends) >= 3:
        ev = ev[ends[0] + 1: ends[1] + 1]
    t0 = ev[0][0]
    os.makedirs(os.path.dirname(a.out) or '.', exist_ok=True)
    with open(a.out, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['start_us', 'dur_us', 'kernel'])
        for s, e, n, _ in ev:
            w.writerow([f'{s - t0:.1f}', f'{e - s:.1f}', n])
    span = max(e for _, e, _, _ in ev) - t0
    # union of tensor-kernel intervals
    tens = sorted((s, e) for s, e, n, _ in ev if TENSOR.search(n))
    merged = []
    for s, e in tens:
        if merged and s <= merged[-1][1]:
            merged[-1][1] = max(merged[-1][1], e)
        else:
            merged.append([s, e])
    busy = sum(e - s for s, e in merged)
    tens_sum = sum(e - s for s, e in tens)
    print(f'step span {span / 1e3:.2f} ms, {len(ev)} kernels; tensor-core kernels: sum of durations {tens_sum / 1e3:.2f} ms, union {busy / 1e3:.2f} ms; '
          f'tensor-idle {(span - busy) / 1e3:.2f} ms')
    # gaps and what runs in them
    gaps = []
    prev = t0
    for s, e in merged:
        if s > prev:
            gaps.append((prev, s))
        prev = max(prev, e)
    if prev < t0 + span:
        gaps.append((prev, t0 + span))
    fill = collections.Counter()
    empty = 0.0
    for gs, ge in gaps:
        covered = []
        for s, e, n, _ in ev:
            if TENSOR.search(n) or e <= gs or s >= ge:
                continue
            lo, hi = max(s, gs), min(e, ge)
            fill[n] += hi - lo
            covered.append((lo, hi))
        covered.sort()
        c = 0.0
        p = gs
        for lo, hi in covered:
            if hi > p:
                c += hi - max(lo, p)
                p = hi
        empty += (ge - gs) - c
    print(f'{len(gaps)} gaps; {empty / 1e3:.2f} ms of them with NO kernel running at all (launch / dependency latency)')
    print('| kernel running while no tensor-core kernel runs | us |')
    print('|---|---|')
    for n, v in fill.most_common(25):
        print(f'| `{n}` | {v:.0f} |')
    big = sorted(gaps, key=lambda g: g[0] - g[1])[:15]
    print('largest gaps (start us, length us, tensor kernel before -> after):')
    for gs, ge in sorted(big):
        before = [n for s, e, n, _ in ev if TENSOR.search(n) and abs(e - gs) < 0.5]
        after = [n for s, e, n, _ in ev if TENSOR.search(n) and abs(s - ge) < 0.5]
        print(f'  {gs - t0:9.1f} {ge - gs:7.1f}  {before[:1]} -> {after[:1]}')


if __name__ == '__main__':
    main()
