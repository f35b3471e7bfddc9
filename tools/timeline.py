"""Kernel timeline of ONE train step in its real (multi-lane, graphed) schedule, from CUPTI through torch.profiler:

    python tools/timeline.py [--batch 128] --out gpurun_out/timeline.csv        # on a GPU box (SIMQ_GRAPH=0 for eager launches)
    python tools/timeline.py --csv profiles/r2_timeline_after.csv               # re-analyse a committed timeline, no GPU

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

TENSOR = re.compile(r'(conv\w*_umma_kernel|wgrad\w*_umma_kernel)')


def short(name):
    return re.sub(r'\(.*', '', name).replace('void ', '')


def analyse(ev):
    """ev: (start us, end us, kernel name) of every kernel of one step.  Returns the step span, the union of the intervals in
    which a tensor-core kernel runs, the tensor-idle time, how much of it has no kernel at all, what fills the rest, and the gaps."""
    ev = sorted(ev)
    t0 = ev[0][0]
    span = max(e for _, e, _ in ev) - t0
    tens = sorted((s, e) for s, e, n in ev if TENSOR.search(n))
    merged = []
    for s, e in tens:
        if merged and s <= merged[-1][1]:
            merged[-1][1] = max(merged[-1][1], e)
        else:
            merged.append([s, e])
    busy = sum(e - s for s, e in merged)
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
        for s, e, n in ev:
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
    return {'t0': t0, 'span': span, 'kernels': len(ev), 'tensor_sum': sum(e - s for s, e in tens), 'tensor_union': busy,
            'tensor_idle': span - busy, 'no_kernel': empty, 'fill': fill, 'gaps': gaps, 'ev': ev}


def report(r):
    out = [f"step span {r['span'] / 1e3:.2f} ms, {r['kernels']} kernels; tensor-core kernels: sum of durations {r['tensor_sum'] / 1e3:.2f} ms, "
           f"union {r['tensor_union'] / 1e3:.2f} ms; tensor-idle {r['tensor_idle'] / 1e3:.2f} ms",
           f"{len(r['gaps'])} gaps; {r['no_kernel'] / 1e3:.2f} ms of them with NO kernel running at all (launch / dependency latency)",
           '| kernel running while no tensor-core kernel runs | us |', '|---|---|']
    for n, v in r['fill'].most_common(25):
        out.append(f'| `{n}` | {v:.0f} |')
    out.append('largest gaps (start us, length us, tensor kernel before -> after):')
    big = sorted(r['gaps'], key=lambda g: g[0] - g[1])[:15]
    for gs, ge in sorted(big):
        before = [n for s, e, n in r['ev'] if TENSOR.search(n) and abs(e - gs) < 0.5]
        after = [n for s, e, n in r['ev'] if TENSOR.search(n) and abs(s - ge) < 0.5]
        out.append(f"  {gs - r['t0']:9.1f} {ge - gs:7.1f}  {before[:1]} -> {after[:1]}")
    return '\n'.join(out)


def load_csv(path):
    """rows written by main(): start_us, dur_us, kernel"""
    with open(path, newline='') as f:
        return [(float(r['start_us']), float(r['start_us']) + float(r['dur_us']), r['kernel']) for r in csv.DictReader(f)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=128)
    ap.add_argument('--warm', type=int, default=6)
    ap.add_argument('--out', default='gpurun_out/timeline.csv')
    ap.add_argument('--csv', default=None, help='re-analyse a timeline written earlier (no GPU needed)')
    a = ap.parse_args()
    if a.csv:
        print(report(analyse(load_csv(a.csv))))
        return
    import torch
    import bench
    from spatial_intention_maps_b200 import networks, synth, train as T
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
    rep = analyse([(s_, e_, n_) for s_, e_, n_, _ in ev])
    print(report(rep))


if __name__ == '__main__':
    main()
