"""GPU probe for the two-term backward GEMMs (simq_set_backward_terms): gradient error against the fp32 reference (oracle) and
against its float64 twin, per setting (dgrad_terms, wgrad_terms), plus single-kernel errors and the step time at B=128.

    python tools/bwd_terms_probe.py > gpurun_out/bwd_terms.json
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from spatial_intention_maps_b200 import _lib, networks, synth, train as T  # noqa: E402
from tests import gpu_checks as G  # noqa: E402


def main():
    out = {'settings': {}}
    cases = {'c1': (4, 2, 16, 11, 0.75, 8), 'c2': (5, 2, 32, 12, 0.85, 16), 'cstar': (8, 2, 16, 15, 0.85, 8)}
    for d, w, mp in ((3, 3, 0), (3, 2, 0), (2, 2, 512), (2, 2, 256)):
        rec = {}
        for name, (C, A, B, seed, gamma, te) in cases.items():
            r = G.train_step_check(C, A, B, seed, gamma, te, 1, fused=True, with_fp64=True, setup=lambda p: p.set_backward_terms(d, w, mp))
            gn = r['grad_norm_ref']
            big = [n for n in r['grad_rel_l2'] if r['grad_ref_norm'][n] >= 1e-6 * gn]
            worst = max(big, key=lambda n: r['grad_rel_l2_64'][n])
            rec[name] = {'flat_vs_fp32_ref': r['flat_grad_rel_l2'], 'flat_vs_fp64': r['flat_grad_rel_l2_64'], 'ref32_vs_fp64': r['flat_ref32_rel_l2_64'],
                         'worst_tensor_vs_fp64': [worst, r['grad_rel_l2_64'][worst], r['ref32_rel_l2_64'][worst]],
                         'worst_tensor_vs_fp32_ref': max(r['grad_rel_l2'][n] for n in big), 'loss': r['loss'][0], 'loss_ref': r['loss_ref'][0],
                         'grad_norm': r['grad_norm'], 'grad_norm_ref': gn, 'param_rel_l2_max': max(r['param_rel_l2'].values())}
        # single kernels vs torch fp64 (dgrad / wgrad of 256->256 3x3 and 64->64 3x3)
        ctx = _lib.Ctx(0, 4, 2, 3)
        _lib.check(_lib.lib().simq_set_backward_terms(ctx.handle, d, w, 0), 'set')
        rec['kernel_err'] = {f'{m}_{ci}x{co}': G.conv_check(ci, co, 3, mode, 0, B=3, ctx=ctx)
                             for m, mode in (('dgrad', 1), ('wgrad', 2)) for ci, co in ((64, 64), (256, 256), (512, 512))}
        ctx.close()
        # speed
        dev = torch.device('cuda', 0)
        B = 128
        pol, tgt, opt = bench.make_nets(networks, torch, dev, bench.C_IN, bench.A_OUT, B)
        pol.set_backward_terms(d, w, mp)
        dbs = [T.DeviceBatch(B, bench.C_IN, dev).upload(T.HostBatch(B, bench.C_IN).fill(
            synth.synth_batch(B, bench.C_IN, bench.A_OUT, 1234 + i, terminal_every=64))) for i in range(4)]
        for i in range(10):
            T.train_step_device(pol, tgt, opt, dbs[i % 4], B, bench.GAMMA, 100, True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for i in range(40):
            T.train_step_device(pol, tgt, opt, dbs[i % 4], B, bench.GAMMA, 100, True)
        e1.record(); torch.cuda.synchronize()
        rec['ms_per_step_b128'] = e0.elapsed_time(e1) / 40
        del pol, tgt, opt, dbs
        torch.cuda.empty_cache()
        out['settings'][f'dgrad{d}_wgrad{w}_min{mp}'] = rec
        print(f'dgrad{d}_wgrad{w}_min{mp}', json.dumps(rec), file=sys.stderr, flush=True)
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
