"""Measured gradient errors of the default mode against the fp32 oracle for every golden case (to set the test bars):
flat rel-L2, worst tensor rel-L2, gradient-norm error.   python tools/grad_bars_probe.py > gpurun_out/grad_bars.json"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from tests import gpu_checks as G  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')
FILES = {'cstar': 'steps_cstar.npz', 'cstar128': 'steps_cstar128.npz', 'clip1': 'steps_clip.npz'}
out = {}
for key in ('c1', 'c2', 'c3', 'traj', 'cstar', 'cstar128', 'clip1'):
    g = np.load(os.path.join(GOLD, FILES.get(key, 'steps.npz')))
    C, A, B, nsteps, seed, te = [int(v) for v in g[key + '_cfg']]
    clip = 100.0 if key != 'clip1' else 1.0
    for fused in (True, False):
        r = G.train_step_check(C, A, B, seed, float(g[key + '_gamma']), te, 1, fused=fused, grad_clip=clip)
        gn = r['grad_norm_ref']
        big = [n for n in r['grad_rel_l2'] if r['grad_ref_norm'][n] >= 1e-6 * gn]
        worst = max(big, key=lambda n: r['grad_rel_l2'][n])
        out[f'{key}_{"fused" if fused else "autograd"}'] = {
            'flat': r['flat_grad_rel_l2'], 'worst': [worst, r['grad_rel_l2'][worst]], 'grad_norm_err': abs(r['grad_norm'] - gn) / gn,
            'param_rel_l2_max': max(r['param_rel_l2'].values()), 'bn_err': r['bn_err'],
            'mom_max': max(v for n, v in r['mom_rel_l2'].items() if r['grad_ref_norm'][n] >= 1e-6 * gn) if r['mom_rel_l2'] else None,
            'loss_err': abs(r['loss'][0] - float(g[key + '_loss'][0])) / abs(float(g[key + '_loss'][0]))}
        print(key, fused, out[f'{key}_{"fused" if fused else "autograd"}'], file=sys.stderr, flush=True)
print(json.dumps(out, indent=1))
