"""One train step (or one forward + backward) at the bench workload between cudaProfilerStart / Stop, for ncu:

    SIMQ_GRAPH=0 ncu --profile-from-start off --metrics ... python tools/one_step.py [--mode step|fwdbwd] [--batch 128]

Eager launches (SIMQ_GRAPH=0) so that every kernel of the step is a launch ncu can see; the warm-up steps run before the
profiler is switched on."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from spatial_intention_maps_b200 import networks, synth, train as T  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--mode', default='step', choices=['step', 'fwdbwd'])
    ap.add_argument('--batch', type=int, default=128)
    ap.add_argument('--warm', type=int, default=2)
    a = ap.parse_args()
    dev = torch.device('cuda', 0)
    B = a.batch
    pol, tgt, opt = bench.make_nets(networks, torch, dev, bench.C_IN, bench.A_OUT, B)
    hb = T.HostBatch(B, bench.C_IN).fill(synth.synth_batch(B, bench.C_IN, bench.A_OUT, 1234, terminal_every=bench.TERMINAL_EVERY))
    db = T.DeviceBatch(B, bench.C_IN, dev).upload(hb)
    x = db.s.permute(0, 3, 1, 2)
    dq = torch.zeros((B, bench.A_OUT, 96, 96), device=dev)
    dq.view(B, -1)[:, 7] = 1.0 / B

    def run():
        if a.mode == 'step':
            T.train_step_device(pol, tgt, opt, db, B, bench.GAMMA, 100, True)
        else:
            opt.zero_grad(set_to_none=True)
            pol(x).backward(dq)
    for _ in range(a.warm):
        run()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    run()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print('one_step done: loss', float(db.out2[0]))


if __name__ == '__main__':
    main()
