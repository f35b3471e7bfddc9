"""Shared helpers of the GPU parity tests and tools/diag_gpu.py: every function runs the CUDA path
through the C-ABI (via the Python host side) and the CPU oracle on the same seeded inputs and
returns error metrics.  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import fcn_oracle as O                                   # noqa: E402
from spatial_intention_maps_b200 import _lib, networks, synth, train as simq_train   # noqa: E402

DEV = 'cuda:0'


def relerr(a: torch.Tensor, b: torch.Tensor) -> float:
    """max|a-b| / max|b|  (the north_star's Q-map tolerance is 1e-3 of this)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def make_net(Cin, A, seed, max_batch=32, backend=_lib.BACKEND_UMMA):
    st = O.make_state(Cin, A, seed)
    net = networks.FCN(Cin, A, max_batch=max_batch)
    net.load_state_dict(st)
    net = net.to(DEV)
    net.set_backend(backend)
    return net, st


# ---------------------------------------------------------------------------------------------
# single convolutions through simq_test_conv
# ---------------------------------------------------------------------------------------------
def conv_check(Cin, Cout, k, mode, backend, B=2, seed=0, ctx=None):
    """mode 0 forward, 1 dgrad, 2 wgrad against torch fp64 on the CPU. Returns max-norm relative error."""
    g = torch.Generator().manual_seed(seed * 7919 + Cin * 31 + Cout * 17 + k + mode)
    x = torch.randn(B, Cin, 24, 24, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) * (2.0 / (Cin * k * k)) ** 0.5
    dy = torch.randn(B, Cout, 24, 24, generator=g)
    xd, wd, dyd = x.double().requires_grad_(True), w.double().requires_grad_(True), dy.double()
    y = F.conv2d(xd, wd, None, 1, k // 2)
    if mode == 0:
        ref, a, a2 = y.detach(), x, None
    else:
        gx, gw = torch.autograd.grad(y, [xd, wd], dyd)
        ref, a, a2 = (gx, dy, None) if mode == 1 else (gw, x, dy)
    own = ctx is None
    if own:
        ctx = _lib.Ctx(0, 4, 2, max(B, 2))
    out = torch.empty(ref.shape, dtype=torch.float32, device=DEV)
    a_d = a.to(DEV).contiguous()
    a2_d = a2.to(DEV).contiguous() if a2 is not None else None
    w_d = w.to(DEV).contiguous()
    _lib.check(_lib.lib().simq_test_conv(ctx.handle, backend, mode, B, Cin, Cout, k, _lib.ptr(a_d), _lib.ptr(a2_d), _lib.ptr(w_d),
                                         _lib.ptr(out), _lib.stream_ptr()), 'simq_test_conv')
    torch.cuda.synchronize()
    if own:
        ctx.close()
    return relerr(out, ref)


CONV_SHAPES = [(64, 64, 3), (64, 128, 3), (128, 128, 3), (128, 256, 3), (256, 256, 3), (256, 512, 3), (512, 512, 3),
               (64, 128, 1), (128, 256, 1), (256, 512, 1), (512, 128, 1)]


# ---------------------------------------------------------------------------------------------
# forward with per-layer trace
# ---------------------------------------------------------------------------------------------
DEBUG_IDS = {'raw0': 0, 'a0': 1, 'raw_h1': 34, 'u1': 35, 'raw_h2': 36, 't': 37}
for _b in range(8):
    for _j, _n in enumerate(('raw1', 'b1', 'raw2', 'out')):
        DEBUG_IDS[f'blk{_b}.{_n}'] = 2 + 4 * _b + _j
    if _b in (2, 4, 6):
        DEBUG_IDS[f'blk{_b}.rawd'] = 38 + _b


def debug_get(net, name, B, saved):
    ref_shapes = None
    chw = C.c_int64()
    buf = torch.empty(B * 512 * 48 * 48 // 4 + B * 128 * 48 * 48, dtype=torch.float32, device=DEV)
    _lib.check(_lib.lib().simq_debug_get(net.ctx().handle, 0 if saved else 1, DEBUG_IDS[name], B, _lib.ptr(buf), C.byref(chw),
                                         _lib.stream_ptr()), 'simq_debug_get')
    torch.cuda.synchronize()
    return buf[:B * chw.value].clone()


def forward_trace_check(Cin, A, B, seed, training, backend=_lib.BACKEND_UMMA, uniform=False):
    """Runs one forward on the GPU and the oracle; returns ({layer: relerr}, q_gpu, q_ref, bn_err)."""
    net, st = make_net(Cin, A, seed, max_batch=B, backend=backend)
    net.train(training)
    x_np = synth.synth_states(B, Cin, seed, uniform=uniform)
    x = O.hwc_to_nchw(list(x_np))
    st_ref = O.clone_state(st)
    with torch.no_grad():
        ref = O.forward_trace(st_ref, x, training)
        q = net(x.to(DEV))
    errs = {}
    for name in DEBUG_IDS:
        if name not in ref:
            continue
        if not training and backend == _lib.BACKEND_UMMA and name.endswith(('.raw1', '.raw2', '.rawd')):
            continue        # eval mode folds BN into the conv epilogues: the raw conv outputs never materialise
        mine = debug_get(net, name, B, saved=False).view(ref[name].shape)
        errs[name] = relerr(mine, ref[name])
    errs['q'] = relerr(q, ref['q'])
    bn_err = 0.0
    sd = net.state_dict()
    for n, _, kind in O.state_spec(Cin, A):
        if kind == 'buffer':
            bn_err = max(bn_err, float((sd[n].cpu() - st_ref[n]).abs().max() / (st_ref[n].abs().max() + 1e-6)))
        if kind == 'nbt':
            assert int(sd[n]) == int(st_ref[n]), f'{n}: {int(sd[n])} != {int(st_ref[n])}'
    return errs, q.cpu(), ref['q'], bn_err


def argmax_agreement(q, q_ref):
    """(#equal indices, #near-ties among the unequal ones, B): SURVEY.md §7.2-6 policy."""
    B = q.shape[0]
    a, b = q.reshape(B, -1).argmax(1), q_ref.reshape(B, -1).argmax(1)
    eq = int((a == b).sum())
    near = 0
    for i in range(B):
        if a[i] != b[i]:
            r = q_ref.reshape(B, -1)[i]
            if float(r[b[i]] - r[a[i]]) <= 1e-6 * float(q_ref.abs().max()):
                near += 1
    return eq, near, B


# ---------------------------------------------------------------------------------------------
# the DQN update
# ---------------------------------------------------------------------------------------------
def batch_tensors(batch):
    s = O.hwc_to_nchw(list(batch.state))
    nf = [n for n in batch.next_state if n is not None]
    ns = O.hwc_to_nchw(nf) if nf else torch.zeros(0, *s.shape[1:])
    mask = torch.tensor([n is not None for n in batch.next_state])
    return s, torch.tensor(batch.action), torch.tensor(batch.reward, dtype=torch.float32), ns, mask


class Cfg:
    def __init__(self, B, C):
        self.batch_size, self.num_input_channels = B, C
        self.use_double_dqn, self.grad_norm_clipping = True, 100
        self.robot_config = [{'lifting_robot': 1}]
        self.final_exploration, self.checkpoint_path, self.policy_path = 0.01, None, None


def train_step_check(Cin, A, B, seed, gamma, terminal_every, nsteps=1, backend=_lib.BACKEND_UMMA, fused=True, with_fp64=False, resync=True,
                     double_dqn=True, grad_clip=100.0, setup=None):
    """nsteps updates on the GPU (fused simq_train_step, or the autograd path with a stock SGD exactly as
    the reference's train.py drives it) and in the oracle.  Returns a dict of error metrics."""
    pol, st = make_net(Cin, A, seed, max_batch=B, backend=backend)
    tgt = networks.FCN(Cin, A, max_batch=B)
    tgt.load_state_dict(st)
    tgt = tgt.to(DEV).eval()
    pol.train()
    opt = torch.optim.SGD(pol.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)       # train.py:186
    cfg = Cfg(B, Cin)
    cfg.use_double_dqn = double_dqn
    cfg.grad_norm_clipping = grad_clip                    # train.py:133: None skips clip_grad_norm_
    if setup is not None:
        setup(pol)
    o_pol, o_tgt, o_mom = O.clone_state(st), O.clone_state(st), None
    out = {'loss': [], 'td': [], 'loss_ref': [], 'td_ref': []}
    names = O.trainable_names(Cin, A)
    for step in range(nsteps):
        batch = synth.synth_batch(B, Cin, A, seed + 1000 * step, terminal_every=terminal_every)
        if step == 0 and with_fp64:        # double-precision twin of the reference: the yardstick for gradients
            s_, a_, r_, ns_, m_ = batch_tensors(batch)
            r64 = O.dqn_step(O.clone_state(st, torch.float64), O.clone_state(st, torch.float64), None, s_.double(), a_,
                             r_.double(), ns_.double(), m_, discount=gamma, apply_update=False)
        r = O.dqn_step(o_pol, o_tgt, o_mom, *batch_tensors(batch), discount=gamma, double_dqn=double_dqn, grad_clip=grad_clip)
        o_mom = r['momentum']
        if fused:
            info = simq_train.train(cfg, pol, tgt, opt, batch, None, gamma)
            grads = pol.flat_grad()
            po = pol._layout[2]
            gmap = {n: grads[po[i]:po[i + 1]].view(p.shape) for i, (n, p) in enumerate(pol.trainable())}
        else:
            info = reference_style_train(cfg, pol, tgt, opt, batch, gamma)
            gmap = {n: p.grad for n, p in pol.trainable()}
        out['loss'].append(info['loss']); out['td'].append(info['td_error'])
        out['loss_ref'].append(r['loss']); out['td_ref'].append(r['td_error'])
        if resync and step + 1 < nsteps:
            # teacher forcing: continue from the ORACLE's state (parameters, BN buffers, momentum), so that step k
            # checks the k-th update itself (momentum rule, BN running statistics, counters) instead of the chaotic
            # divergence of two trajectories (a Double-DQN arg-max flip on a B=8 batch moves the loss by several %)
            out.setdefault('param_rel_l2_steps', []).append(
                max(rel_l2(pol.state_dict()[n], o_pol[n]) for n in names))
            pol.load_state_dict(o_pol)
            for n, p in pol.trainable():
                opt.state[p]['momentum_buffer'].copy_(o_mom[n])
        if step == 0:
            out['grad_rel_l2'] = {n: rel_l2(gmap[n], r['grads'][n]) for n in names}
            gn = float(torch.sqrt(sum((gmap[n].double() ** 2).sum() for n in names)))
            coef = 1.0 if grad_clip is None else min(1.0, float(grad_clip) / (r['grad_norm'] + 1e-6))
            out['grad_norm'], out['grad_norm_ref'], out['clip_coef_ref'] = gn, r['grad_norm'] * coef, coef
            # the UPDATE itself (clip coefficient x momentum rule x lr): p_after - p_before, ours vs the oracle's
            sd0 = pol.state_dict()
            d_mine = torch.cat([(sd0[n].detach().double().cpu() - st[n].double()).reshape(-1) for n in names])
            d_ref = torch.cat([(o_pol[n].double() - st[n].double()).reshape(-1) for n in names])
            out['update_rel_l2'] = float((d_mine - d_ref).norm() / d_ref.norm().clamp_min(1e-30))
            out['update_norm_ratio'] = float(d_mine.norm() / d_ref.norm().clamp_min(1e-30))
            flat = lambda d: torch.cat([d[n].detach().double().cpu().reshape(-1) for n in names])
            out['flat_grad_rel_l2'] = rel_l2(flat(gmap), flat(r['grads']))
            out['grad_abs'] = {n: float((gmap[n].double().cpu() - r['grads'][n].double()).abs().max()) for n in names}
            out['grad_ref_norm'] = {n: float(r['grads'][n].double().norm()) for n in names}
            if with_fp64:
                out['grad_rel_l2_64'] = {n: rel_l2(gmap[n], r64['grads'][n]) for n in names}
                out['ref32_rel_l2_64'] = {n: rel_l2(r['grads'][n], r64['grads'][n]) for n in names}
                out['flat_grad_rel_l2_64'] = rel_l2(flat(gmap), flat(r64['grads']))
                out['flat_ref32_rel_l2_64'] = rel_l2(flat(r['grads']), flat(r64['grads']))
                out['loss_64'] = r64['loss']
    sd = pol.state_dict()
    out['param_rel_l2'] = {n: rel_l2(sd[n], o_pol[n]) for n in names}
    out['bn_err'] = max(float((sd[n].cpu() - o_pol[n]).abs().max() / (o_pol[n].abs().max() + 1e-6))
                        for n, _, k in O.state_spec(Cin, A) if k == 'buffer')
    out['nbt'] = [int(sd[n]) for n, _, k in O.state_spec(Cin, A) if k == 'nbt']
    out['nbt_ref'] = [int(o_pol[n]) for n, _, k in O.state_spec(Cin, A) if k == 'nbt']
    out['fc_untouched'] = bool(torch.equal(sd['resnet18.fc.weight'].cpu(), st['resnet18.fc.weight']))
    mom = {n: opt.state[p]['momentum_buffer'] for n, p in pol.trainable() if 'momentum_buffer' in opt.state[p]}
    out['mom_rel_l2'] = {n: rel_l2(mom[n], o_mom[n]) for n in names} if len(mom) == len(names) else None
    return out


def reference_style_train(cfg, policy_net, target_net, optimizer, batch, discount_factor):
    """The body of the reference's train.train (train.py:108-141) written against the public
    nn.Module surface only -- exercises FCN's autograd.Function with a stock optimizer."""
    dev = DEV
    s, a, r, ns, mask = batch_tensors(batch)
    s, a, r, ns, mask = s.to(dev), a.to(dev), r.to(dev), ns.to(dev), mask.to(dev)
    output = policy_net(s)
    q = output.view(cfg.batch_size, -1).gather(1, a.unsqueeze(1)).squeeze(1)
    nv = torch.zeros(cfg.batch_size, dtype=torch.float32, device=dev)
    with torch.no_grad():
        if ns.shape[0] > 0:
            if cfg.use_double_dqn:
                best = policy_net(ns).view(ns.size(0), -1).max(1)[1].view(ns.size(0), 1)
                nv[mask] = target_net(ns).view(ns.size(0), -1).gather(1, best).view(-1)
            else:
                nv[mask] = target_net(ns).view(ns.size(0), -1).max(1)[0]
    y = r + discount_factor * nv
    td = torch.abs(q - y).detach()
    loss = F.smooth_l1_loss(q, y)
    optimizer.zero_grad()
    loss.backward()
    if cfg.grad_norm_clipping is not None:                                     # train.py:133
        torch.nn.utils.clip_grad_norm_(policy_net.parameters(), cfg.grad_norm_clipping)
    optimizer.step()
    return {'td_error': td.mean().item(), 'loss': loss.item()}


def intention_step_check(Ct, B, seed, nsteps=2, backend=_lib.BACKEND_UMMA):
    """train_intention on the GPU (simq_intention_step) vs the oracle, teacher-forced between steps."""
    net, st = make_net(Ct - 1, 1, seed, max_batch=B, backend=backend)
    net.train()
    opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)       # train.py:190
    o_net, o_mom = O.clone_state(st), None
    names = O.trainable_names(Ct - 1, 1)
    out = {'loss': [], 'loss_ref': []}
    for step in range(nsteps):
        batch = synth.synth_batch(B, Ct, 2, seed + 1000 * step, terminal_every=None)
        r = O.intention_step(o_net, o_mom, batch.state)
        o_mom = r['momentum']
        info = simq_train.train_intention(net, opt, batch, None)
        out['loss'].append(info['loss_intention']); out['loss_ref'].append(r['loss_intention'])
        if step == 0:
            grads, po = net.flat_grad(), net._layout[2]
            gmap = {n: grads[po[i]:po[i + 1]].view(p.shape) for i, (n, p) in enumerate(net.trainable())}
            flat = lambda d: torch.cat([d[n].detach().double().cpu().reshape(-1) for n in names])
            out['flat_grad_rel_l2'] = rel_l2(flat(gmap), flat(r['grads']))
            out['grad_rel_l2'] = {n: rel_l2(gmap[n], r['grads'][n]) for n in names}
            out['grad_ref_norm'] = {n: float(r['grads'][n].double().norm()) for n in names}
            out['grad_norm_ref'] = float(flat(r['grads']).norm())
        out.setdefault('param_rel_l2_steps', []).append(max(rel_l2(net.state_dict()[n], o_net[n]) for n in names))
        if step + 1 < nsteps:
            net.load_state_dict(o_net)
            for n, p in net.trainable():
                opt.state[p]['momentum_buffer'].copy_(o_mom[n])
    sd = net.state_dict()
    out['nbt'] = [int(sd[n]) for n, _, k in O.state_spec(Ct - 1, 1) if k == 'nbt']
    out['nbt_ref'] = [int(o_net[n]) for n, _, k in O.state_spec(Ct - 1, 1) if k == 'nbt']
    return out
