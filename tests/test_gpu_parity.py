"""Parity of the CUDA path (through the C-ABI) with the CPU oracle and the committed golden
fixtures.  Runs on the B200 box: ``pytest -m gpu``.

Tolerances (BASELINE.json north_star): Q-map max|d|/max|Q| <= 1e-3 in fp32, per-sample arg-max action
index equal (a differing index is accepted only if the oracle's own Q-values at the two indices
differ by <= 1e-6 max|Q|: fp32 reduction order cannot be matched bit-for-bit)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

QTOL = 1e-3
G = None


@pytest.fixture(scope='module', autouse=True)
def _need_gpu():
    global G
    assert torch.cuda.is_available(), 'gpu-marked tests need a CUDA device'
    from tests import gpu_checks
    G = gpu_checks


GOLD = os.path.join(os.path.dirname(__file__), 'golden')


@pytest.mark.parametrize('backend', [0, 1], ids=['umma', 'fma'])
@pytest.mark.parametrize('mode', [0, 1, 2], ids=['fwd', 'dgrad', 'wgrad'])
def test_conv_kernels_match_torch_fp64(backend, mode):
    from spatial_intention_maps_b200 import _lib
    ctx = _lib.Ctx(0, 4, 2, 3)
    _lib.check(_lib.lib().simq_set_backward_terms(ctx.handle, 3, 3, 0), 'simq_set_backward_terms')     # the exact (three-term) kernels
    for (ci, co, k) in G.CONV_SHAPES:
        e = G.conv_check(ci, co, k, mode, backend, B=3, ctx=ctx)
        assert e < 2e-4, f'{ci}->{co} k{k} mode {mode} backend {backend}: rel err {e:.3e}'
    if backend == 0 and mode > 0:
        # two-term backward GEMMs (dy contributes its bf16 hi plane only): the error is dy's bf16 rounding, 2^-9 per element at most
        _lib.check(_lib.lib().simq_set_backward_terms(ctx.handle, 2, 2, 0), 'simq_set_backward_terms')
        for (ci, co, k) in G.CONV_SHAPES:
            e = G.conv_check(ci, co, k, mode, backend, B=3, ctx=ctx)
            assert 2e-4 < e < 4e-3, f'two-term {ci}->{co} k{k} mode {mode}: rel err {e:.3e}'
    ctx.close()


@pytest.mark.parametrize('C,A', [(4, 2), (5, 2), (5, 1), (8, 2), (3, 2), (10, 2)])
def test_forward_matches_golden(C, A):
    """Fixtures were produced by the reference's own networks.FCN (tests/golden/make_golden.py)."""
    from oracle import fcn_oracle as O
    from spatial_intention_maps_b200 import synth
    g = np.load(os.path.join(GOLD, 'forward.npz'))
    key = f'C{C}_A{A}'
    seed = int(g[key + '_seed'])
    net, st = G.make_net(C, A, seed, max_batch=2)
    x = O.hwc_to_nchw(list(synth.synth_states(2, C, seed))).to(G.DEV)
    with torch.no_grad():
        net.eval()
        q_eval = net(x).cpu()
        net.train()
        q_train = net(x).cpu()
    for mine, ref in ((q_eval, torch.from_numpy(g[key + '_q_eval'])), (q_train, torch.from_numpy(g[key + '_q_train']))):
        assert G.relerr(mine, ref) <= QTOL
        eq, near, B = G.argmax_agreement(mine, ref)
        assert eq + near == B
    sd = net.state_dict()
    np.testing.assert_allclose(sd['bn1.running_mean'].cpu().numpy(), g[key + '_bn1_rm'], rtol=2e-3, atol=1e-5)
    np.testing.assert_allclose(sd['bn1.running_var'].cpu().numpy(), g[key + '_bn1_rv'], rtol=2e-3, atol=1e-6)
    np.testing.assert_allclose(sd['resnet18.layer4.1.bn2.running_mean'].cpu().numpy(), g[key + '_l4_rm'], rtol=2e-3, atol=2e-5)
    np.testing.assert_allclose(sd['resnet18.layer4.1.bn2.running_var'].cpu().numpy(), g[key + '_l4_rv'], rtol=2e-3, atol=1e-6)
    assert int(sd['bn2.num_batches_tracked']) == int(g[key + '_nbt'])


@pytest.mark.parametrize('training', [False, True])
def test_forward_every_layer_matches_oracle(training):
    errs, q, qr, bn = G.forward_trace_check(5, 2, 4, 105, training)
    bad = {k: v for k, v in errs.items() if v > QTOL}
    assert not bad, f'layers above tolerance: {bad}'
    eq, near, B = G.argmax_agreement(q, qr)
    assert eq + near == B
    assert bn < 2e-3


def test_nchw_and_nhwc_inputs_agree():
    net, _ = G.make_net(5, 2, 7, max_batch=3)
    net.eval()
    from spatial_intention_maps_b200 import synth
    x = torch.from_numpy(synth.synth_states(3, 5, 7)).to(G.DEV)            # (3,96,96,5) NHWC
    with torch.no_grad():
        a = net(x.permute(0, 3, 1, 2))                                      # channels_last view
        b = net(x.permute(0, 3, 1, 2).contiguous())                         # NCHW copy
    assert torch.equal(a, b)


def _check_grads(r):
    """Gradients are ill-conditioned (ReLU / max-pool mask flips under rounding noise): the fp32
    reference itself is 3e-3 .. 5e-3 rel-L2 away from its own float64 twin (profiles/r2_bwd_terms_probe.md,
    r2_cstar128_gradient_probe.json).  Our operands carry 16 mantissa bits (bf16 hi+lo), so ~100x more pre-activations
    sit within rounding noise of zero than in fp32.  Measured over every golden case, fused and autograd routes
    (profiles/r2_v1_grad_errors_per_case.json): flat 1.2e-2 .. 2.8e-2 (worst: c* at batch 128), worst tensor 3.2e-2 =
    7.8x the fp32 reference's own error on that tensor.  Bars: whole gradient vector within 3e-2 rel-L2 of the
    reference, every tensor within 5e-2, and tensors whose true value is zero up to rounding (conv biases feeding a
    train-mode BN) within 1e-5 of the gradient norm in absolute terms."""
    gn = r['grad_norm_ref']
    assert r['flat_grad_rel_l2'] < 3e-2, f"flat gradient rel-L2 {r['flat_grad_rel_l2']:.3e}"
    for n, e in r['grad_rel_l2'].items():
        if r['grad_ref_norm'][n] < 1e-6 * gn:
            assert r['grad_abs'][n] <= 1e-5 * gn, f'{n}: abs err {r["grad_abs"][n]:.3e}'
        else:
            assert e < 5e-2, f'gradient {n} rel-L2 {e:.3e}'
    assert abs(r['grad_norm'] - gn) <= 2e-3 * gn


GOLDEN_FILE = {'cstar': 'steps_cstar.npz', 'cstar128': 'steps_cstar128.npz', 'clip1': 'steps_clip.npz', 'clip5': 'steps_clip.npz',
               'clipnone': 'steps_clip.npz'}


@pytest.mark.parametrize('key', ['c1', 'c2', 'c3', 'traj', 'cstar', 'cstar128', 'clip1', 'clip5', 'clipnone'])
@pytest.mark.parametrize('fused', [True, False], ids=['fused', 'autograd'])
def test_dqn_step_matches_oracle_and_golden(key, fused):
    """c1 / c2 / c3 (the full-size bench workload: B=128, C=5, A=1) of BASELINE.json, a 3-step trajectory, the north_star's headline shape
    c* (C=8, A=2) at B=16 and at its full batch 128, and the three branches of train.py:133-134: clip_grad_norm_ engaged (max-norm 1 and 5
    against a gradient norm of ~21: coef < 1), inactive (max-norm 100, every other case) and skipped (``grad_norm_clipping: None``).
    The golden losses / gradient norms come from the reference's own train.train (tests/golden/steps*.npz)."""
    g = np.load(os.path.join(GOLD, GOLDEN_FILE.get(key, 'steps.npz')))
    C, A, B, nsteps, seed, te = [int(v) for v in g[key + '_cfg']]
    clip = 100.0
    if key + '_clip' in g:
        clip = None if float(g[key + '_clip']) < 0 else float(g[key + '_clip'])
    r = G.train_step_check(C, A, B, seed, float(g[key + '_gamma']), te, nsteps, fused=fused, grad_clip=clip)
    np.testing.assert_allclose(r['loss'][0], g[key + '_loss'][0], rtol=1e-3)
    np.testing.assert_allclose(r['td'][0], g[key + '_td'][0], rtol=1e-3)
    if key.startswith('clip'):
        # the gradients left in .grad are the rescaled ones: their norm equals the reference's (= max-norm when engaged)
        np.testing.assert_allclose(r['grad_norm'], float(g[key + '_grad_norm']), rtol=2e-3)
        assert (r['clip_coef_ref'] < 0.5) == (clip is not None), r['clip_coef_ref']
        # and the parameter UPDATE (coef x lr x momentum rule) matches in size and direction
        assert abs(r['update_norm_ratio'] - 1.0) < 5e-3 and r['update_rel_l2'] < 3e-2, (r['update_norm_ratio'], r['update_rel_l2'])
        # step 2 starts from the oracle's state after ITS clipped update: equals the reference's second loss
        np.testing.assert_allclose(r['loss_ref'], g[key + '_loss'], rtol=2e-4)
    # multi-step: every step restarts from the oracle's state (teacher forcing, see gpu_checks.train_step_check):
    # free-running B=8 trajectories diverge chaotically (a Double-DQN arg-max flip moves the loss by several %;
    # make_golden.py notes 1e-5 -> 30 % by step 5 even for the fp32 oracle against the reference)
    np.testing.assert_allclose(r['loss'], r['loss_ref'], rtol=1e-3)
    np.testing.assert_allclose(r['td'], r['td_ref'], rtol=1e-3)
    _check_grads(r)
    worst = max(r['param_rel_l2'].items(), key=lambda kv: kv[1])
    assert worst[1] < (2e-3 if nsteps == 1 else 1e-2), f'parameter {worst[0]} rel-L2 {worst[1]:.3e}'
    assert all(e < 1e-2 for e in r.get('param_rel_l2_steps', [])), r['param_rel_l2_steps']
    assert r['bn_err'] < 5e-3
    assert r['nbt'] == r['nbt_ref'] == list(g[key + '_nbt'])
    assert r['fc_untouched']
    if r['mom_rel_l2'] is not None:
        assert max(v for n, v in r['mom_rel_l2'].items() if r['grad_ref_norm'][n] >= 1e-6 * r['grad_norm_ref']) < 5e-2


@pytest.mark.parametrize('case', ['all_terminal', 'batch1', 'plain_dqn', 'ragged_b37_c10_a1'])
def test_dqn_step_edge_cases(case):
    """Edges of train.py:108-141: every transition terminal (empty next-state batch: train.py:112 would cat an
    empty list, the target term vanishes), a single-sample batch (BN statistics over one image), the
    non-double branch (train.py:124), and a ragged batch (B=37: M = 23125 rows is not a tile multiple) with
    the widest input (C=10) and one output channel."""
    C, A, B, te, dd = {'all_terminal': (4, 2, 4, 1, True), 'batch1': (4, 2, 1, 2, True), 'plain_dqn': (5, 2, 8, 4, False),
                       'ragged_b37_c10_a1': (10, 1, 37, 5, True)}[case]
    for fused in (True, False):
        r = G.train_step_check(C, A, B, 40 + B, 0.85, te, 1, fused=fused, double_dqn=dd)
        np.testing.assert_allclose(r['loss'], r['loss_ref'], rtol=1e-3)
        np.testing.assert_allclose(r['td'], r['td_ref'], rtol=1e-3)
        assert r['flat_grad_rel_l2'] < (3e-2 if B > 1 else 1e-1), r['flat_grad_rel_l2']
        assert r['nbt'] == r['nbt_ref'] and r['fc_untouched'] and r['bn_err'] < 5e-3


def test_gradients_against_float64_twin():
    """c1: our gradient error against the float64 twin of the reference must be of the order of the
    fp32 reference's own error against it."""
    r = G.train_step_check(4, 2, 16, 11, 0.75, 8, 1, fused=True, with_fp64=True)
    assert abs(r['loss'][0] - r['loss_64']) <= 1e-4 * abs(r['loss_64'])
    assert r['flat_grad_rel_l2_64'] < 3e-2
    print('flat gradient rel-L2 vs float64: ours %.3e, fp32 reference %.3e' % (r['flat_grad_rel_l2_64'], r['flat_ref32_rel_l2_64']))
    gn = r['grad_norm_ref']
    for n, e in r['grad_rel_l2_64'].items():
        if r['grad_ref_norm'][n] >= 1e-6 * gn:
            assert e < max(1e-1, 10 * r['ref32_rel_l2_64'][n]), f'{n}: {e:.3e} (fp32 reference: {r["ref32_rel_l2_64"][n]:.3e})'


def test_default_backward_terms_meet_the_acceptance_rule():
    """The default backward scheme (weight gradients and layer-4 input gradients with two operand terms) was accepted under the
    rule: every gradient / parameter bar unchanged, and the gradient error against the reference's float64 twin may grow by at
    most 10 % over the three-term backward.  Re-measured here on c1 (measured: +5.9 %)."""
    full = G.train_step_check(4, 2, 16, 11, 0.75, 8, 1, fused=True, with_fp64=True, setup=lambda p: p.set_backward_terms(3, 3, 0))
    dflt = G.train_step_check(4, 2, 16, 11, 0.75, 8, 1, fused=True, with_fp64=True)
    print('flat gradient rel-L2 vs float64: three-term backward %.4e, default %.4e (%+.1f %%), fp32 reference %.4e' % (
        full['flat_grad_rel_l2_64'], dflt['flat_grad_rel_l2_64'], 100 * (dflt['flat_grad_rel_l2_64'] / full['flat_grad_rel_l2_64'] - 1),
        dflt['flat_ref32_rel_l2_64']))
    assert dflt['flat_grad_rel_l2_64'] <= 1.10 * full['flat_grad_rel_l2_64']
    assert full['loss'] == dflt['loss']                      # the forward passes are untouched
    _check_grads(dflt)


def test_intention_step_matches_oracle_and_golden():
    """train.train_intention (train.py:143-158): dense dL/dQ through the whole decoder backward; golden
    losses from the reference's own train_intention (tests/golden/intention.npz)."""
    g = np.load(os.path.join(GOLD, 'intention.npz'))
    C, B, seed = [int(v) for v in g['cfg']]
    r = G.intention_step_check(C, B, seed, nsteps=2)
    np.testing.assert_allclose(r['loss'], g['loss'], rtol=1e-3)
    np.testing.assert_allclose(r['loss'], r['loss_ref'], rtol=1e-3)
    assert r['flat_grad_rel_l2'] < 3e-2, r['flat_grad_rel_l2']
    for n, e in r['grad_rel_l2'].items():
        if r['grad_ref_norm'][n] >= 1e-6 * r['grad_norm_ref']:
            assert e < 1e-1, f'gradient {n} rel-L2 {e:.3e}'
    assert all(e < 1e-2 for e in r['param_rel_l2_steps']), r['param_rel_l2_steps']
    assert r['nbt'] == r['nbt_ref'] == list(g['nbt'])


def test_device_replay_buffer_feeds_the_same_update():
    """train.train on a DeviceSample (device gather, simq_gather_rows) == train.train on the equivalent host
    Transition batch, bit for bit."""
    import random
    from spatial_intention_maps_b200 import networks, synth, train as T
    from spatial_intention_maps_b200.replay import ReplayBuffer
    tr = synth.synth_batch(24, 5, 2, 9, terminal_every=5)
    buf = ReplayBuffer(20)                                      # wraps: 24 pushes into 20 slots
    for i in range(24):
        buf.push(tr.state[i], tr.action[i], tr.reward[i], tr.next_state[i])
    random.seed(3)
    sample = buf.sample(8)
    host_batch = sample.to_transition()
    losses = []
    for batch in (sample, host_batch):
        net, st = G.make_net(5, 2, 17, max_batch=8)
        tgt = networks.FCN(5, 2, max_batch=8)
        tgt.load_state_dict(st)
        tgt = tgt.to(G.DEV).eval()
        net.train()
        opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
        info = T.train(G.Cfg(8, 5), net, tgt, opt, batch, None, 0.85)
        losses.append((info['loss'], info['td_error'], net.flat_params.clone()))
    assert losses[0][0] == losses[1][0] and losses[0][1] == losses[1][1]
    assert torch.equal(losses[0][2], losses[1][2])


def test_cuda_graph_replay_matches_eager(monkeypatch):
    """simq_train_step replays the step as a CUDA graph from the second call on; six updates (two distinct
    non-terminal counts -> two graphs, first update eager) must leave exactly the parameters, BN buffers and
    losses of the eager launch path (SIMQ_GRAPH=0)."""
    from spatial_intention_maps_b200 import networks, synth, train as T
    results = []
    for mode in ('1', '0'):
        monkeypatch.setenv('SIMQ_GRAPH', mode)
        net, st = G.make_net(4, 2, 23, max_batch=8)
        tgt = networks.FCN(4, 2, max_batch=8)
        tgt.load_state_dict(st)
        tgt = tgt.to(G.DEV).eval()
        net.train()
        opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
        losses = []
        for step in range(6):
            batch = synth.synth_batch(8, 4, 2, 100 + step, terminal_every=4 if step % 2 else 8)
            if step == 4:                                       # a target sync in between (train.py:267-269)
                tgt.load_state_dict(net.state_dict())
            losses.append(T.train(G.Cfg(8, 4), net, tgt, opt, batch, None, 0.75)['loss'])
        results.append((losses, net.flat_params.clone(), net.flat_bn.clone(), net.flat_nbt.clone()))
    assert results[0][0] == results[1][0]
    assert torch.equal(results[0][1], results[1][1]) and torch.equal(results[0][2], results[1][2])
    assert torch.equal(results[0][3], results[1][3])


@pytest.mark.parametrize('graph', ['1', '0'], ids=['graph', 'eager'])
@pytest.mark.parametrize('B,C,A,double_dqn', [(8, 4, 2, True), (37, 10, 1, True), (8, 5, 2, False)])
def test_lane_schedule_is_bit_identical_to_serial(monkeypatch, graph, B, C, A, double_dqn):
    """The two-lane schedule (forward on s beside the forwards on s', weight gradients beside the dgrad chain;
    two branches of the CUDA graph) must leave exactly the losses, parameters, momentum, BN running statistics and
    counters of the serial schedule over six updates -- incl. an all-terminal batch (no s' lane), two non-terminal
    counts, a target sync, the plain-DQN branch (only the target pass on the s' lane), and the deferred
    running-statistics update of the concurrent s' pass."""
    from spatial_intention_maps_b200 import networks, synth, train as T
    monkeypatch.setenv('SIMQ_GRAPH', graph)
    results = []
    for sched in ('lanes', 'serial'):
        net, st = G.make_net(C, A, 29, max_batch=B)
        net.set_schedule(sched)
        tgt = networks.FCN(C, A, max_batch=B)
        tgt.load_state_dict(st)
        tgt = tgt.to(G.DEV).eval()
        net.train()
        opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
        cfg = G.Cfg(B, C)
        cfg.use_double_dqn = double_dqn
        nbt0 = int(net.flat_nbt[0])
        out = []
        for step in range(6):
            te = 1 if step == 3 else (4 if step % 2 else 8)       # step 3: every transition terminal
            batch = synth.synth_batch(B, C, A, 300 + step, terminal_every=te)
            if step == 4:
                tgt.load_state_dict(net.state_dict())
            r = T.train(cfg, net, tgt, opt, batch, None, 0.85)
            out.append((r['loss'], r['td_error']))
        results.append((out, net.flat_params.clone(), net.flat_bn.clone(), net.flat_nbt.clone(), net.flat_momentum.clone()))
    assert results[0][0] == results[1][0]
    for i, name in ((1, 'params'), (2, 'bn'), (3, 'nbt'), (4, 'momentum')):
        a, b = results[0][i], results[1][i]
        assert torch.equal(a, b), (name, int((a != b).sum()), float((a.double() - b.double()).abs().max()),
                                   (a != b).nonzero().flatten()[:8].tolist())
    assert int(results[0][3][0]) == nbt0 + 6 + (5 if double_dqn else 0)      # +1 per s pass, +1 per s' online pass


def test_lane_schedule_autograd_and_intention_paths():
    """The weight-gradient lane also serves simq_fcn_backward (stock autograd route) and simq_intention_step:
    gradients / updated parameters equal the serial schedule bit for bit."""
    from spatial_intention_maps_b200 import synth, train as T
    grads, params = [], []
    for sched in ('lanes', 'serial'):
        net, _ = G.make_net(5, 2, 31, max_batch=6)
        net.set_schedule(sched)
        net.train()
        x = torch.from_numpy(np.stack(synth.synth_states(6, 5, 77))).to(G.DEV).permute(0, 3, 1, 2)
        q = net(x)
        q.square().mean().backward()
        grads.append(net.flat_grad().clone())
        inet, _ = G.make_net(4, 1, 33, max_batch=6)
        inet.set_schedule(sched)
        inet.train()
        opt = torch.optim.SGD(inet.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
        batch = synth.synth_batch(6, 5, 1, 55, terminal_every=8)
        for _ in range(3):
            T.train_intention(inet, opt, batch, None)
        params.append(inet.flat_params.clone())
    assert torch.equal(grads[0], grads[1])
    assert torch.equal(params[0], params[1])


def test_c4_two_robot_groups_training_block_matches_oracle():
    """Config c4 (lifting_2_pushing_2: two robot groups -> an A=2 and an A=1 Q-network, gamma 0.85 each, predicted
    intention on -> two intention nets): ``train.train_groups`` -- the training block of the reference's main loop
    (train.py:253-263), all four updates enqueued back to back, one host sync -- returns the reference's
    ``all_train_info`` keys with losses within 1e-3 of the oracle's and leaves the networks bit-identical to four
    separate ``train`` / ``train_intention`` calls."""
    from oracle import fcn_oracle as O
    from spatial_intention_maps_b200 import policies, synth, train as T
    B, C = 8, 5
    cfg = G.Cfg(B, C)
    cfg.robot_config = [{'lifting_robot': 2}, {'pushing_robot': 2}]
    cfg.discount_factors = [0.85, 0.85]
    cfg.use_predicted_intention = True
    outs = []
    for grouped in (True, False):
        pol = policies.DQNIntentionPolicy(cfg, train=True, device=G.DEV)
        assert [n.module.num_output_channels for n in pol.policy_nets] == [2, 1]
        states = []
        for i, net in enumerate(pol.policy_nets + pol.intention_nets):
            st = O.make_state(net.module.num_input_channels, net.module.num_output_channels, 40 + i)
            net.module.load_state_dict(st)
            net.train()
            states.append(st)
        tgts = pol.build_policy_nets()
        for t, n in zip(tgts, pol.policy_nets):
            t.load_state_dict(n.state_dict()); t.eval()
        sgd = lambda n: torch.optim.SGD(n.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)   # noqa: E731
        opts, opts_i = [sgd(n) for n in pol.policy_nets], [sgd(n) for n in pol.intention_nets]
        batches = [synth.synth_batch(B, C, A, 70 + i, terminal_every=4) for i, A in enumerate((2, 1))]
        if grouped:
            info = T.train_groups(cfg, pol, tgts, opts, batches, opts_i)
        else:
            info = {}
            for i in range(2):
                r = T.train(cfg, pol.policy_nets[i], tgts[i], opts[i], batches[i], pol.apply_transform, cfg.discount_factors[i])
                r.update(T.train_intention(pol.intention_nets[i], opts_i[i], batches[i], pol.apply_transform))
                info.update({'{}/robot_group_{:02}'.format(k, i + 1): v for k, v in r.items()})
        outs.append((info, [n.module.flat_params.clone() for n in pol.policy_nets + pol.intention_nets]))
        if grouped:
            assert sorted(info) == sorted(f'{k}/robot_group_{g:02}' for k in ('loss', 'td_error', 'loss_intention') for g in (1, 2))
            for i in range(2):
                ref = O.dqn_step(O.clone_state(states[i]), O.clone_state(states[i]), None, *G.batch_tensors(batches[i]), discount=0.85)
                tag = f'robot_group_{i + 1:02}'
                assert abs(info[f'loss/{tag}'] - float(ref['loss'])) <= 1e-3 * abs(float(ref['loss'])), (i, info, float(ref['loss']))
                assert abs(info[f'td_error/{tag}'] - float(ref['td_error'])) <= 1e-3 * abs(float(ref['td_error']))
                refi = O.intention_step(O.clone_state(states[2 + i]), None, list(batches[i].state))
                assert abs(info[f'loss_intention/{tag}'] - refi['loss_intention']) <= 1e-3 * abs(refi['loss_intention'])
    assert outs[0][0] == outs[1][0]
    for a, b in zip(outs[0][1], outs[1][1]):
        assert torch.equal(a, b)


def test_bf16_mode_is_characterised_and_not_default():
    """simq_set_precision(BF16) -- one MMA per product -- is opt-in: its Q-map error is of the order SURVEY.md
    §7.2-1 predicts for bf16 operands (1e-2: outside the 1e-3 parity bar), the default mode is untouched by it."""
    net, st = G.make_net(5, 2, 105, max_batch=4)
    from oracle import fcn_oracle as O
    from spatial_intention_maps_b200 import synth
    x = O.hwc_to_nchw(list(synth.synth_states(4, 5, 105)))
    with torch.no_grad():
        ref = O.forward(O.clone_state(st), x, False)
        net.eval()
        q_parity = net(x.to(G.DEV)).cpu()
        net.set_precision('bf16')
        q_bf16 = net(x.to(G.DEV)).cpu()
        net.set_precision('parity')
        q_again = net(x.to(G.DEV)).cpu()
    assert G.relerr(q_parity, ref) <= QTOL and torch.equal(q_parity, q_again)
    e = G.relerr(q_bf16, ref)
    assert 1e-3 < e < 1e-1, e


def test_intention_policy_step_matches_oracle():
    """policies.DQNIntentionPolicy.step (policies.py:76-146) in evaluation mode: the predicted intention map
    sigmoid(FCN(C-1 -> 1)(s)) is appended to the state and the greedy action taken on the C-channel result."""
    from oracle import fcn_oracle as O
    from spatial_intention_maps_b200 import policies, synth
    C = 5
    pol = policies.DQNIntentionPolicy(G.Cfg(4, C), train=False)
    st_q, st_i = O.make_state(C, 2, 51), O.make_state(C - 1, 1, 52)
    pol.policy_nets[0].module.load_state_dict(st_q)
    pol.intention_nets[0].module.load_state_dict(st_i)
    for k in range(3):
        s = synth.synth_states(1, C - 1, 60 + k)[0]
        a, info = pol.step([[s]], exploration_eps=0.0, debug=True)
        with torch.no_grad():
            pred = torch.sigmoid(O.forward(O.clone_state(st_i), O.hwc_to_nchw([s]), False))[0, 0].numpy()
        assert np.abs(info['output_intention'][0][0] - pred).max() <= 1e-3
        a_ref, q_ref = O.greedy_action(O.clone_state(st_q), np.concatenate([s, pred[:, :, None]], axis=2))
        assert G.relerr(torch.from_numpy(info['output'][0][0]), torch.from_numpy(q_ref)) <= QTOL
        if a[0][0] != a_ref:
            assert q_ref.reshape(-1)[a_ref] - q_ref.reshape(-1)[a[0][0]] <= 1e-6 * np.abs(q_ref).max()


def test_policy_step_matches_golden():
    """policies.DQNPolicy.step greedy action == the reference's on 16 states."""
    from oracle import fcn_oracle as O
    from spatial_intention_maps_b200 import policies, synth
    g = np.load(os.path.join(GOLD, 'policy_step.npz'))
    C, A, seed = [int(v) for v in g['cfg']]
    pol = policies.DQNPolicy(G.Cfg(16, C), train=False)
    pol.policy_nets[0].module.load_state_dict(O.make_state(C, A, seed))
    states = synth.synth_states(16, C, seed)
    for i in range(16):
        a, info = pol.step([[states[i]]], exploration_eps=0.0, debug=True)
        q = info['output'][0][0]
        assert abs(float(q.max()) - float(g['qmax'][i])) <= QTOL * max(1.0, abs(float(g['qmax'][i])))
        if a[0][0] != int(g['actions'][i]):          # near-tie policy
            ref_q = O.greedy_action(O.make_state(C, A, seed), states[i])[1].reshape(-1)
            assert ref_q[int(g['actions'][i])] - ref_q[a[0][0]] <= 1e-6 * np.abs(ref_q).max()


def test_forward_full_batch128_matches_oracle():
    """c3 at its full size (B=128, C=5, A=1; M = 80000 rows -> the CTA-pair kernels with all 74 pairs busy): the
    train-mode Q-map against the CPU oracle directly, plus BN running statistics."""
    errs, q, qr, bn = G.forward_trace_check(5, 1, 128, 13, True)
    assert errs['q'] <= QTOL, errs['q']
    bad = {k: v for k, v in errs.items() if v > QTOL}
    assert not bad, bad
    eq, near, B = G.argmax_agreement(q, qr)
    assert eq + near == B, (eq, near, B)
    assert bn < 2e-3


@pytest.mark.parametrize('mode', [0, 1, 2], ids=['fwd', 'dgrad', 'wgrad'])
def test_pair_kernels_full_size_against_fma_comparator(mode):
    """B=128, 512 -> 512 3x3 (M = 80000 rows: every CTA pair busy for ~9 rounds; wgrad: 4 row splits of 20000 rows):
    the tcgen05 CTA-pair kernels against the fp32-FMA comparator kernels on the same split-bf16 operands
    (independent arithmetic), plus linearity out(a + b) == out(a) + out(b) for the forward conv."""
    import ctypes as C
    from spatial_intention_maps_b200 import _lib
    B, Ci, Co = 128, 512, 512
    ctx = _lib.Ctx(0, 4, 2, B)
    _lib.check(_lib.lib().simq_set_backward_terms(ctx.handle, 3, 3, 0), 'simq_set_backward_terms')     # compare the exact kernels
    g = torch.Generator(device=G.DEV).manual_seed(5 + mode)
    a = torch.randn(B, Ci, 24, 24, device=G.DEV, generator=g)
    a2 = torch.randn(B, Co, 24, 24, device=G.DEV, generator=g) if mode == 2 else None
    w = torch.randn(Co, Ci, 3, 3, device=G.DEV, generator=g) * (2.0 / (Ci * 9)) ** 0.5
    shape = (Co, Ci, 3, 3) if mode == 2 else (B, Co if mode == 0 else Ci, 24, 24)

    def run(backend, x):
        out = torch.empty(shape, dtype=torch.float32, device=G.DEV)
        _lib.check(_lib.lib().simq_test_conv(ctx.handle, backend, mode, B, Ci, Co, 3, _lib.ptr(x), _lib.ptr(a2), _lib.ptr(w), _lib.ptr(out),
                                             _lib.stream_ptr()), 'simq_test_conv')
        return out
    fast, slow = run(_lib.BACKEND_UMMA, a), run(_lib.BACKEND_FMA, a)
    assert float((fast - slow).abs().max() / slow.abs().max()) < (2e-4 if mode == 2 else 5e-5)
    if mode == 0:
        b = torch.randn(B, Ci, 24, 24, device=G.DEV, generator=g)
        lin = run(_lib.BACKEND_UMMA, a + b) - fast - run(_lib.BACKEND_UMMA, b)
        assert float(lin.abs().max() / fast.abs().max()) < 1e-4
    torch.cuda.synchronize()
    ctx.close()


def test_workspace_grows_with_the_batch():
    """A forward larger than the context's max_batch re-allocates the workspace transparently (same result as
    a network created for that batch)."""
    from spatial_intention_maps_b200 import synth
    small, _ = G.make_net(4, 2, 3, max_batch=2)
    big, _ = G.make_net(4, 2, 3, max_batch=6)
    x = torch.from_numpy(synth.synth_states(6, 4, 8)).to(G.DEV).permute(0, 3, 1, 2)
    small.eval(); big.eval()
    with torch.no_grad():
        small(x[:2])
        assert torch.equal(small(x), big(x)) and small.max_batch >= 6


def test_full_size_properties_b128():
    """c3 size (B=128, C=5, A=1): eval forward is per-sample independent (a sample's Q-map does not
    depend on its batch-mates), greedy_action == arg-max of the forward, and a train step is finite,
    moves the parameters and bumps num_batches_tracked by 2."""
    from spatial_intention_maps_b200 import networks, synth, train as T
    net, st = G.make_net(5, 1, 13, max_batch=128)
    x = torch.from_numpy(synth.synth_states(128, 5, 13)).to(G.DEV).permute(0, 3, 1, 2)
    net.eval()
    with torch.no_grad():
        q = net(x)
        q_sub = net(x[5:9])
        act, _ = net.greedy_action(x)
    assert G.relerr(q_sub, q[5:9]) < 1e-4      # not bitwise: small batches take the split-K path (different summation order)
    assert torch.equal(act, q.view(128, -1).argmax(1))
    tgt = networks.FCN(5, 1, max_batch=128)
    tgt.load_state_dict(st)
    tgt = tgt.to(G.DEV).eval()
    net.train()
    opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    before = net.flat_params.clone()
    info = T.train(G.Cfg(128, 5), net, tgt, opt, synth.synth_batch(128, 5, 1, 13), None, 0.85)
    assert np.isfinite(info['loss']) and np.isfinite(info['td_error'])
    assert torch.isfinite(net.flat_params).all() and not torch.equal(before, net.flat_params)
    assert int(net.bn1.num_batches_tracked) == 3 + 2


def test_out_of_range_action_is_reported_not_indexed():
    """train.py:115 gathers Q(s, a): an action index outside [0, A*96*96) raises in the reference.  Here the tail kernel must
    not index out of bounds: it flags the context's device error word, the step reports NaN, and train.train raises."""
    from spatial_intention_maps_b200 import _lib, networks, synth, train as T
    net, st = G.make_net(4, 2, 3, max_batch=4)
    tgt = networks.FCN(4, 2, max_batch=4)
    tgt.load_state_dict(st)
    tgt = tgt.to(G.DEV).eval()
    net.train()
    opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    good = synth.synth_batch(4, 4, 2, 3, terminal_every=2)
    bad = good._replace(action=(good.action[0], 2 * 9216, good.action[2], -1))
    with pytest.raises(_lib.SimqError, match='action index'):
        T.train(G.Cfg(4, 4), net, tgt, opt, bad, None, 0.75)
    info = T.train(G.Cfg(4, 4), net, tgt, opt, good, None, 0.75)          # the context stays usable, the flag was cleared
    assert np.isfinite(info['loss'])


def test_deepcopy_target_network_idiom():
    """``target = copy.deepcopy(policy_net)`` (the common PyTorch DQN idiom): the copy computes the same Q-map from its own flat
    vectors / context, and keeps computing the OLD Q-map after the original has been updated."""
    import copy
    from spatial_intention_maps_b200 import synth
    net, _ = G.make_net(4, 2, 9, max_batch=2)
    x = torch.from_numpy(synth.synth_states(2, 4, 9)).to(G.DEV).permute(0, 3, 1, 2)
    net.eval()
    with torch.no_grad():
        q0 = net(x)
        cp = copy.deepcopy(net)
        assert torch.equal(cp(x), q0)
        for p in net.parameters():
            p.mul_(0.5)
        assert not torch.equal(net(x), q0) and torch.equal(cp(x), q0)
        for p in cp.parameters():                      # a write through .data is invisible to the version counters ...
            p.data.mul_(0.5)
        cp.mark_params_changed()                       # ... so the caller says so (Polyak-style updates)
        assert torch.equal(cp(x), net(x))


def test_two_phase_step_is_bit_identical_to_the_whole_step():
    """simq_train_step_phase: phase 1 (forwards, tail, backward through head + layer 4) followed by phase 2 (rest of the backward)
    leaves exactly the gradients / BN statistics / report of the one-call step, eager and graph-replayed, and after phase 1 the
    tail of the gradient vector (layer 4 + head) is already final -- what the data-parallel path all-reduces early."""
    from spatial_intention_maps_b200 import _lib, networks, synth, train as T
    L = _lib.lib()
    B, C, A = 6, 5, 2
    outs = []
    for phased in (False, True):
        net, st = G.make_net(C, A, 61, max_batch=B)
        tgt = networks.FCN(C, A, max_batch=B)
        tgt.load_state_dict(st)
        tgt = tgt.to(G.DEV).eval()
        net.train()
        net.flat_momentum = torch.zeros_like(net.flat_params)
        db = T.DeviceBatch(B, C, G.DEV).upload(T.HostBatch(B, C).fill(synth.synth_batch(B, C, A, 62, terminal_every=3)))
        torch.cuda.synchronize()
        grads, out2 = net.flat_grad(), torch.zeros(2, device=G.DEV)
        split = net.grad_bucket_split()
        tails = []

        def call(phase):
            _lib.check(L.simq_train_step_phase(
                net.ctx(B).handle, _lib.ptr(net.flat_params), _lib.ptr(net.flat_bn), _lib.ptr(net.flat_nbt), _lib.ptr(tgt.flat_params),
                _lib.ptr(tgt.flat_bn), tgt.params_version, _lib.ptr(grads), _lib.ptr(net.flat_momentum), _lib.ptr(db.s), _lib.ptr(db.ns),
                _lib.X_NHWC, _lib.ptr(db.action), _lib.ptr(db.reward), _lib.ptr(db.nonfinal), B, db.Bn, 0.85, 0.01, 0.9, 1e-4, 100.0, 1, 1, 0,
                _lib.ptr(out2), phase, _lib.stream_ptr()), 'simq_train_step_phase')
        for rep in range(3):                                   # 1st call eager, 2nd captures, 3rd replays
            grads.zero_()
            if phased:
                call(1)
                torch.cuda.synchronize()
                tails.append(grads[split:].clone())
                call(2)
            else:
                call(0)
            torch.cuda.synchronize()
            outs.append((phased, rep, grads.clone(), net.flat_bn.clone(), net.flat_nbt.clone(), out2.clone()))
            if phased:
                assert torch.equal(tails[-1], grads[split:]) and float(grads[:split].abs().sum()) > 0
    for rep in range(3):
        a = next(o for o in outs if not o[0] and o[1] == rep)
        b = next(o for o in outs if o[0] and o[1] == rep)
        for i in range(2, 6):
            assert torch.equal(a[i], b[i]), (rep, i)


def test_free_running_trajectory_b128_drift_is_bounded():
    """Free-running (NOT teacher-forced) updates at the bench size: four consecutive train.train calls on c3 (B=128, C=5, A=1), ours
    and the oracle's each continuing from its OWN state.  The dynamics of these first steps are violent (lr 0.01 on a fresh net: the
    loss goes 0.155 -> 0.217 -> 0.200 -> 0.165), so the ~1.5 % gradient difference of one step is amplified: measured loss drift
    1.3e-5, 5.1e-3, 5.1e-2, 1.9e-2 at steps 1-4 (td_error 1.3e-5, 3.0e-3, 2.7e-2, 1.7e-2).  Bars: step 1 exact to 1e-3, step 2 to 2e-2,
    every step to 1e-1 -- a drift bound, not a parity claim (the parity of the update rule is what the teacher-forced test checks)."""
    r = G.train_step_check(5, 1, 128, 13, 0.85, 64, 4, fused=True, resync=False)
    drift = [abs(a - b) / abs(b) for a, b in zip(r['loss'], r['loss_ref'])]
    tdd = [abs(a - b) / abs(b) for a, b in zip(r['td'], r['td_ref'])]
    print('free-running loss drift per step:', ['%.2e' % d for d in drift], 'td:', ['%.2e' % d for d in tdd],
          'worst parameter rel-L2 after 4 steps: %.2e' % max(r['param_rel_l2'].values()))
    assert drift[0] < 1e-3 and tdd[0] < 1e-3
    assert drift[1] < 2e-2 and tdd[1] < 2e-2
    assert max(drift) < 1e-1 and max(tdd) < 1e-1
    assert max(r['param_rel_l2'].values()) < 5e-2
    assert r['nbt'] == r['nbt_ref']


def test_new_target_network_does_not_evict_the_policy_weights():
    """The library keeps two packed-weight slots per context.  Swapping in freshly built target networks (new parameter pointers) step
    after step must never evict the POLICY's slot in the middle of a step (round 1 raised a one-off 'packed policy weights evicted')."""
    from spatial_intention_maps_b200 import networks, synth, train as T
    net, st = G.make_net(4, 2, 71, max_batch=4)
    net.train()
    opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    batch = synth.synth_batch(4, 4, 2, 72, terminal_every=4)
    losses = []
    for k in range(5):
        tgt = networks.FCN(4, 2, max_batch=4)                # a NEW target object (new flat vectors) every step
        tgt.load_state_dict(st)
        tgt = tgt.to(G.DEV).eval()
        with torch.no_grad():
            net.eval(); net(torch.from_numpy(synth.synth_states(1, 4, k)).to(G.DEV).permute(0, 3, 1, 2)); net.train()   # policy.step between updates
        losses.append(T.train(G.Cfg(4, 4), net, tgt, opt, batch, None, 0.75)['loss'])
    assert all(np.isfinite(losses)) and len(set(losses)) > 1
