"""The oracle (oracle/fcn_oracle.py) against fixtures produced by the reference itself
(tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import fcn_oracle as O
from spatial_intention_maps_b200 import synth

G = os.path.join(os.path.dirname(__file__), 'golden')


def test_manifest_matches_reference_state_dict():
    m = np.load(os.path.join(G, 'manifest_C5_A2.npz'))
    spec = O.state_spec(5, 2)
    assert len(spec) == 138
    assert ['module.' + n for n, _, _ in spec] == list(m['names'])
    assert [str(tuple(s)) for _, s, _ in spec] == list(m['shapes'])
    assert sum(k == 'param' for _, _, k in spec) == 72          # incl. the never-executed resnet18.fc.*
    assert len(O.trainable_names(5, 2)) == 70
    n_train = sum(int(np.prod(s)) for n, s, k in spec if k == 'param' and not n.startswith('resnet18.fc.'))
    assert n_train == 11252962


@pytest.mark.parametrize('C,A', [(4, 2), (5, 2), (5, 1), (8, 2), (3, 2), (10, 2)])
def test_forward_matches_reference(C, A):
    g = np.load(os.path.join(G, 'forward.npz'))
    key = f'C{C}_A{A}'
    seed = int(g[key + '_seed'])
    st = O.make_state(C, A, seed)
    x = O.hwc_to_nchw(list(synth.synth_states(2, C, seed)))
    with torch.no_grad():
        q_eval = O.forward(st, x, False).numpy()
        q_train = O.forward(st, x, True).numpy()
    # same library primitives as the reference => essentially bit-equal; allow thread-count jitter
    for mine, ref in ((q_eval, g[key + '_q_eval']), (q_train, g[key + '_q_train'])):
        assert np.abs(mine - ref).max() <= 2e-5 * np.abs(ref).max()
        assert (mine.reshape(2, -1).argmax(1) == ref.reshape(2, -1).argmax(1)).all()
    np.testing.assert_allclose(st['bn1.running_mean'].numpy(), g[key + '_bn1_rm'], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(st['bn1.running_var'].numpy(), g[key + '_bn1_rv'], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(st['resnet18.layer4.1.bn2.running_mean'].numpy(), g[key + '_l4_rm'], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(st['resnet18.layer4.1.bn2.running_var'].numpy(), g[key + '_l4_rv'], rtol=1e-4, atol=1e-6)
    assert int(st['bn2.num_batches_tracked']) == int(g[key + '_nbt'])


def _batch_tensors(batch):
    s = O.hwc_to_nchw(list(batch.state))
    nf = [n for n in batch.next_state if n is not None]
    ns = O.hwc_to_nchw(nf) if nf else torch.zeros(0, *s.shape[1:])
    mask = torch.tensor([n is not None for n in batch.next_state])
    return s, torch.tensor(batch.action), torch.tensor(batch.reward, dtype=torch.float32), ns, mask


def run_oracle_steps(C, A, B, nsteps, seed, te, gamma, clip=100.0):
    pol = O.make_state(C, A, seed)
    tgt = O.clone_state(pol)
    mom, first = None, None
    infos = []
    for step in range(nsteps):
        batch = synth.synth_batch(B, C, A, seed + 1000 * step, terminal_every=te)
        r = O.dqn_step(pol, tgt, mom, *_batch_tensors(batch), discount=gamma, grad_clip=clip)
        mom = r['momentum']
        infos.append((r['loss'], r['td_error']))
        if first is None:
            first = r
    return pol, mom, first, infos


GOLDEN_FILE = {'cstar': 'steps_cstar.npz', 'cstar128': 'steps_cstar128.npz', 'clip1': 'steps_clip.npz', 'clip5': 'steps_clip.npz',
               'clipnone': 'steps_clip.npz'}


@pytest.mark.parametrize('key', ['c1', 'traj', 'c2', 'cstar', 'clip1', 'clip5', 'clipnone', 'cstar128'])
def test_dqn_step_matches_reference(key):
    """clip1 / clip5: clip_grad_norm_ ENGAGED (train.py:133-134, coef < 1); clipnone: ``grad_norm_clipping: None`` (the branch
    that skips it); cstar128: the north_star's headline shape at its full batch (C=8, A=2, B=128)."""
    g = np.load(os.path.join(G, GOLDEN_FILE.get(key, 'steps.npz')))
    C, A, B, nsteps, seed, te = [int(v) for v in g[key + '_cfg']]
    clip = 100.0
    if key + '_clip' in g:
        clip = None if float(g[key + '_clip']) < 0 else float(g[key + '_clip'])
    pol, mom, first, infos = run_oracle_steps(C, A, B, nsteps, seed, te, float(g[key + '_gamma']), clip)
    if clip is not None:                       # the golden gradients are the ones clip_grad_norm_ rescaled in place
        first = dict(first, grad_norm=min(first['grad_norm'], clip * first['grad_norm'] / (first['grad_norm'] + 1e-6)))
        if clip < 100.0:
            assert first['grad_norm'] < 1.0001 * clip and float(g[key + '_grad_norm']) > 0.999 * clip   # the clip really is active
    np.testing.assert_allclose([i[0] for i in infos], g[key + '_loss'], rtol=2e-4)
    np.testing.assert_allclose([i[1] for i in infos], g[key + '_td'], rtol=2e-4)
    names = O.trainable_names(C, A)
    gd = np.stack([O.digest(first['grads'][n]) for n in names])
    # compare L2 norms tightly, sums loosely (cancellation)
    np.testing.assert_allclose(gd[:, 2], g[key + '_grad_digest'][:, 2], rtol=2e-3, atol=1e-7)
    assert abs(first['grad_norm'] - float(g[key + '_grad_norm'])) <= 1e-3 * float(g[key + '_grad_norm'])
    pd = np.stack([O.digest(pol[n]) for n in names])
    ptol = 1e-5 if nsteps == 1 else 1e-4      # multi-step: sensitivity to 1-ulp differences grows per step
    np.testing.assert_allclose(pd[:, 2], g[key + '_param_digest'][:, 2], rtol=ptol)
    np.testing.assert_allclose(pd[:, 3:], g[key + '_param_digest'][:, 3:], rtol=1e-3, atol=20 * ptol)
    md = np.stack([O.digest(mom[n]) for n in names])
    np.testing.assert_allclose(md[:, 2], g[key + '_mom_digest'][:, 2], rtol=5e-3 if nsteps == 1 else 3e-2, atol=1e-7)
    bn = np.concatenate([pol[n].numpy().ravel() for n, _, k in O.state_spec(C, A) if k == 'buffer'])
    np.testing.assert_allclose(bn, g[key + '_bn'], rtol=1e-3, atol=1e-4)
    nbt = [int(pol[n]) for n, _, k in O.state_spec(C, A) if k == 'nbt']
    assert nbt == list(g[key + '_nbt'])            # +2 per train() call
    np.testing.assert_allclose(O.digest(pol['resnet18.fc.weight']), g[key + '_fc_w_digest'])  # untouched


def test_greedy_action_matches_reference_policy_step():
    g = np.load(os.path.join(G, 'policy_step.npz'))
    C, A, seed = [int(v) for v in g['cfg']]
    st = O.make_state(C, A, seed)
    states = synth.synth_states(16, C, seed)
    for i in range(16):
        a, q = O.greedy_action(st, states[i])
        assert a == int(g['actions'][i])
        assert abs(float(q.max()) - float(g['qmax'][i])) <= 1e-5 * max(1.0, abs(float(g['qmax'][i])))


def test_intention_step_matches_reference():
    """train.train_intention (train.py:143-158) run by the reference itself (tests/golden/intention.npz)."""
    g = np.load(os.path.join(G, 'intention.npz'))
    C, B, seed = [int(v) for v in g['cfg']]
    net = O.make_state(C - 1, 1, seed)
    mom, losses, first = None, [], None
    for step in range(2):
        batch = synth.synth_batch(B, C, 2, seed + 1000 * step, terminal_every=None)
        r = O.intention_step(net, mom, batch.state)
        mom = r['momentum']
        losses.append(r['loss_intention'])
        first = first or r
    np.testing.assert_allclose(losses, g['loss'], rtol=1e-5)
    names = O.trainable_names(C - 1, 1)
    gd = np.stack([O.digest(first['grads'][n]) for n in names])
    np.testing.assert_allclose(gd[:, 2], g['grad_digest'][:, 2], rtol=2e-3, atol=1e-9)
    pd = np.stack([O.digest(net[n]) for n in names])
    np.testing.assert_allclose(pd[:, 2], g['param_digest'][:, 2], rtol=1e-5)
    assert [int(net[n]) for n, _, k in O.state_spec(C - 1, 1) if k == 'nbt'] == list(g['nbt'])
