"""GPU diagnostic sweep (run on the B200 box): conv unit checks for both back-ends, per-layer forward
traces, one train step -- prints every metric, never stops at the first failure.
    python tools/diag_gpu.py [conv] [fwd] [step]   > gpurun_out/diag.log
"""
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from tests import gpu_checks as G  # noqa: E402
from spatial_intention_maps_b200 import _lib  # noqa: E402

parts = sys.argv[1:] or ['conv', 'fwd', 'step']
BK = {'fma': _lib.BACKEND_FMA, 'umma': _lib.BACKEND_UMMA}


def guard(label, fn):
    t = time.time()
    try:
        r = fn()
        print(f'[ok  ] {label}: {r}  ({time.time() - t:.1f}s)', flush=True)
        return r
    except Exception as e:  # noqa: BLE001
        print(f'[FAIL] {label}: {type(e).__name__}: {e}', flush=True)
        traceback.print_exc()
        try:
            torch.cuda.synchronize()
        except Exception as e2:  # noqa: BLE001
            print('      device unusable after failure:', e2, flush=True)
            sys.exit(3)
        return None


print(torch.cuda.get_device_name(0), flush=True)
if 'conv' in parts:
    ctx = _lib.Ctx(0, 4, 2, 2)
    for bname in ('fma', 'umma'):
        for (ci, co, k) in G.CONV_SHAPES:
            for mode in (0, 1, 2):
                guard(f'conv {bname} {ci}->{co} k{k} mode{mode}', lambda: f'{G.conv_check(ci, co, k, mode, BK[bname], ctx=ctx):.3e}')
    ctx.close()

if 'fwd' in parts:
    for bname in ('fma', 'umma'):
        for training in (False, True):
            def run():
                errs, q, qr, bn = G.forward_trace_check(5, 2, 4, 105, training, BK[bname])
                worst = sorted(errs.items(), key=lambda kv: -kv[1])[:4]
                first_bad = next((n for n in errs if errs[n] > 1e-3), None)
                return (f"q={errs['q']:.3e} bn={bn:.3e} argmax={G.argmax_agreement(q, qr)} first_bad={first_bad} worst={worst} "
                        f"all={ {k: float(f'{v:.2e}') for k, v in errs.items()} }")
            guard(f'forward {bname} training={training}', run)

if 'step' in parts:
    for bname in ('fma', 'umma'):
        for fused in (True, False):
            def run():
                r = G.train_step_check(4, 2, 16, 11, 0.75, 8, 1, BK[bname], fused)
                g = r['grad_rel_l2']
                worst = sorted(g.items(), key=lambda kv: -kv[1])[:6]
                pw = sorted(r['param_rel_l2'].items(), key=lambda kv: -kv[1])[:3]
                return (f"FLAT_GRAD_REL_L2={r['flat_grad_rel_l2']:.3e} loss={r['loss']} ref={r['loss_ref']} td={r['td']} ref={r['td_ref']} gnorm={r['grad_norm']:.6g}/{r['grad_norm_ref']:.6g} "
                        f"bn={r['bn_err']:.2e} nbt_ok={r['nbt'] == r['nbt_ref']} fc={r['fc_untouched']} worst_grads={worst} worst_params={pw} "
                        f"mom={None if r['mom_rel_l2'] is None else max(r['mom_rel_l2'].values()):.2e}")
            guard(f'train step {bname} fused={fused}', run)
print('diag done', flush=True)
