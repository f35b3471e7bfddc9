"""c* at its full batch (C=8, A=2, B=128): gradient error of the default and the three-term backward against the fp32 oracle and its
float64 twin.   python tools/cstar128_probe.py > gpurun_out/cstar128.json"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import gpu_checks as G  # noqa: E402

out = {}
for name, setup in (('three_term', lambda p: p.set_backward_terms(3, 3, 0)), ('default', None)):
    r = G.train_step_check(8, 2, 128, 17, 0.85, 64, 1, fused=True, with_fp64=True, setup=setup)
    gn = r['grad_norm_ref']
    big = [n for n in r['grad_rel_l2'] if r['grad_ref_norm'][n] >= 1e-6 * gn]
    w = max(big, key=lambda n: r['grad_rel_l2_64'][n])
    out[name] = {'flat_vs_fp32_ref': r['flat_grad_rel_l2'], 'flat_vs_fp64': r['flat_grad_rel_l2_64'], 'ref32_vs_fp64': r['flat_ref32_rel_l2_64'],
                 'worst_tensor_vs_fp64': [w, r['grad_rel_l2_64'][w], r['ref32_rel_l2_64'][w]],
                 'worst5_vs_fp32': sorted(((r['grad_rel_l2'][n], n) for n in big), reverse=True)[:5]}
    print(name, out[name], file=sys.stderr, flush=True)
print(json.dumps(out, indent=1))
