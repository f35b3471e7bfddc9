"""A/B of the two step schedules (simq_set_schedule) with nvidia-smi clock / power sampling during each timed region:
tells a time-bound step (lanes hide the elementwise kernels) from a power-bound one (same energy -> same time).

    python tools/ab_schedule.py [--batch 128] [--steps 60]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from spatial_intention_maps_b200 import networks, synth, train as T  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=128)
    ap.add_argument('--steps', type=int, default=60)
    args = ap.parse_args()
    B, dev = args.batch, torch.device('cuda', 0)
    torch.manual_seed(0)
    pol = networks.FCN(bench.C_IN, bench.A_OUT, max_batch=B).to(dev).train()
    tgt = networks.FCN(bench.C_IN, bench.A_OUT, max_batch=B)
    tgt.load_state_dict(pol.state_dict())
    tgt = tgt.to(dev).eval()
    opt = torch.optim.SGD(pol.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    hb = T.HostBatch(B, bench.C_IN).fill(synth.synth_batch(B, bench.C_IN, bench.A_OUT, 1234, terminal_every=bench.TERMINAL_EVERY))
    db = T.DeviceBatch(B, bench.C_IN, dev).upload(hb)
    out = {}
    for rep in range(2):
        for sched in ('serial', 'lanes'):
            pol.set_schedule(sched)
            for _ in range(5):
                T.train_step_device(pol, tgt, opt, db, B, bench.GAMMA, 100, True)
            torch.cuda.synchronize()
            smp = bench.ClockSampler(0)
            smp.start()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                T.train_step_device(pol, tgt, opt, db, B, bench.GAMMA, 100, True)
            e1.record()
            torch.cuda.synchronize()
            clk = smp.stop()
            pw = sorted(float(r[2]) for r in smp.rows if len(r) > 2)
            out[f'{sched}_{rep}'] = {'ms_per_step': e0.elapsed_time(e1) / args.steps, 'sm_mhz_median': clk['sm_mhz'],
                                    'power_w_median': pw[len(pw) // 2] if pw else None, 'power_w_max': clk['power_w_max'],
                                    'reasons': clk['reasons']}
    print(json.dumps({'batch': B, 'steps': args.steps, 'runs': out}))


if __name__ == '__main__':
    main()
