"""CPU oracle for the DQN Q-map hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product path
(``spatial_intention_maps_b200``) never does; it fails loudly without its CUDA library.

What it restates (reference = jimmyyhwu/spatial-intention-maps @ 336e03a):

* ``networks.FCN.forward``           networks.py:16-26   -> :func:`forward`
* ``resnet.ResNet.features``         resnet.py:93-104    -> :func:`_features`
* ``resnet.BasicBlock.forward``      resnet.py:31-47     -> :func:`_basic_block`
* ``resnet.ResNet.__init__`` init    resnet.py:70-75     -> :func:`make_state`
* ``train.train``                    train.py:108-141    -> :func:`dqn_step`
* ``policies.DQNPolicy.step`` greedy policies.py:56-64   -> :func:`greedy_action`

The reference's arithmetic lives in third-party PyTorch (pinned pytorch==1.2.0, README.md:32;
this image has torch 2.11 whose conv/BN/pool/interpolate/smooth_l1/SGD semantics are the same).
The oracle therefore calls the same ``torch.nn.functional`` CPU fp32 primitives on an explicit
parameter dictionary (reference ``state_dict`` names, without the ``module.`` prefix), and writes
out clip + momentum-SGD by hand.

Parity pinning: the reference ships NO tests or golden vectors for this path.
The oracle is pinned instead against outputs of the reference itself, imported in the build
container by ``tests/golden/make_golden.py`` and committed under ``tests/golden/`` (Q-maps, loss,
td_error, gradient / parameter / BN digests, greedy actions).  ``tests/test_oracle_golden.py``
checks the oracle against those fixtures on CPU.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5        # torch.nn.BatchNorm2d default, used by resnet.py:24,27,57 / networks.py:11,13
BN_MOMENTUM = 0.1
STAGE_PLANES = (64, 128, 256, 512)


# --------------------------------------------------------------------------------------
# parameter / buffer inventory (state_dict order of networks.FCN)
# --------------------------------------------------------------------------------------
def _bn_entries(prefix: str, ch: int):
    return [(prefix + '.weight', (ch,), 'param'), (prefix + '.bias', (ch,), 'param'),
            (prefix + '.running_mean', (ch,), 'buffer'), (prefix + '.running_var', (ch,), 'buffer'),
            (prefix + '.num_batches_tracked', (), 'nbt')]


def state_spec(C: int, A: int) -> List[Tuple[str, Tuple[int, ...], str]]:
    """All 138 state_dict entries (name, shape, kind) in reference order, no ``module.`` prefix."""
    out = [('resnet18.conv1.weight', (64, C, 7, 7), 'param')] + _bn_entries('resnet18.bn1', 64)
    inpl = 64
    for li, planes in enumerate(STAGE_PLANES, start=1):
        for blk in range(2):
            p = f'resnet18.layer{li}.{blk}'
            out.append((p + '.conv1.weight', (planes, inpl if blk == 0 else planes, 3, 3), 'param'))
            out += _bn_entries(p + '.bn1', planes)
            out.append((p + '.conv2.weight', (planes, planes, 3, 3), 'param'))
            out += _bn_entries(p + '.bn2', planes)
            if blk == 0 and inpl != planes:      # resnet.py:79-83
                out.append((p + '.downsample.0.weight', (planes, inpl, 1, 1), 'param'))
                out += _bn_entries(p + '.downsample.1', planes)
        inpl = planes
    out += [('resnet18.fc.weight', (1000, 512), 'param'), ('resnet18.fc.bias', (1000,), 'param')]
    out += [('conv1.weight', (128, 512, 1, 1), 'param'), ('conv1.bias', (128,), 'param')]
    out += _bn_entries('bn1', 128)
    out += [('conv2.weight', (32, 128, 1, 1), 'param'), ('conv2.bias', (32,), 'param')]
    out += _bn_entries('bn2', 32)
    out += [('conv3.weight', (A, 32, 1, 1), 'param'), ('conv3.bias', (A,), 'param')]
    return out


def trainable_names(C: int, A: int) -> List[str]:
    """Parameters that receive a gradient (everything but the never-executed resnet18.fc.*)."""
    return [n for n, _, k in state_spec(C, A) if k == 'param' and not n.startswith('resnet18.fc.')]


def make_state(C: int, A: int, seed: int = 0, perturb: bool = True) -> "OrderedDict[str, torch.Tensor]":
    """Deterministic (numpy RandomState) state following the reference's init DISTRIBUTIONS.

    resnet.py:70-75: kaiming_normal_(fan_out, relu) for every conv inside ResNet, BN gamma=1 beta=0.
    Head convs keep torch's Conv2d default (kaiming_uniform(a=sqrt(5)) -> U(+-1/sqrt(fan_in)) for
    weight and bias).  ``perturb`` additionally randomises BN affine params and running stats so
    that tests exercise them (a freshly initialised net has gamma=1, beta=0, mean=0, var=1).
    """
    rs = np.random.RandomState(seed)
    st: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, shape, kind in state_spec(C, A):
        if kind == 'nbt':
            st[name] = torch.tensor(3 if perturb else 0, dtype=torch.int64)
            continue
        if name.endswith('running_mean'):
            v = rs.normal(0, 0.2, shape) if perturb else np.zeros(shape)
        elif name.endswith('running_var'):
            v = rs.uniform(0.5, 1.5, shape) if perturb else np.ones(shape)
        elif len(shape) == 1 and '.bn' in '.' + name or 'downsample.1' in name:
            if name.endswith('weight'):
                v = rs.uniform(0.7, 1.3, shape) if perturb else np.ones(shape)
            else:
                v = rs.normal(0, 0.1, shape) if perturb else np.zeros(shape)
        elif name.startswith('resnet18.') and len(shape) == 4:
            fan_out = shape[0] * shape[2] * shape[3]
            v = rs.normal(0, math.sqrt(2.0 / fan_out), shape)
        else:  # head convs, fc: U(+-1/sqrt(fan_in))
            fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else {
                'conv1.bias': 512, 'conv2.bias': 128, 'conv3.bias': 32, 'resnet18.fc.bias': 512}[name]
            b = 1.0 / math.sqrt(fan_in)
            v = rs.uniform(-b, b, shape)
        st[name] = torch.from_numpy(np.asarray(v, dtype=np.float32).reshape(shape).copy())
    return st


def clone_state(st, dtype=None):
    """Deep copy; ``dtype=torch.float64`` gives the double-precision twin used to measure how far
    the fp32 reference itself is from the exact result (gradients are ill-conditioned: ReLU / max-pool
    mask flips under rounding noise dominate their error)."""
    return OrderedDict((k, (v.to(dtype) if (dtype is not None and v.is_floating_point()) else v.clone()))
                       for k, v in st.items())


# --------------------------------------------------------------------------------------
# forward (networks.py:16-26, resnet.py:31-47, 93-104)
# --------------------------------------------------------------------------------------
def _bn(st, prefix, x, training):
    """BatchNorm2d forward; mutates running stats / num_batches_tracked in train mode exactly like
    nn.BatchNorm2d (momentum 0.1, unbiased var into running_var, +1 even under no_grad)."""
    if training:
        st[prefix + '.num_batches_tracked'] += 1
    return F.batch_norm(x, st[prefix + '.running_mean'], st[prefix + '.running_var'],
                        st[prefix + '.weight'], st[prefix + '.bias'], training, BN_MOMENTUM, BN_EPS)


def _basic_block(st, p, x, training, has_ds):
    out = F.conv2d(x, st[p + '.conv1.weight'], None, 1, 1)            # resnet.py:34
    out = F.relu(_bn(st, p + '.bn1', out, training))                 # :35-36
    out = F.conv2d(out, st[p + '.conv2.weight'], None, 1, 1)          # :38
    out = _bn(st, p + '.bn2', out, training)                          # :39
    identity = x
    if has_ds:                                                        # :41-42
        identity = _bn(st, p + '.downsample.1', F.conv2d(x, st[p + '.downsample.0.weight']), training)
    return F.relu(out + identity)                                     # :44-45


def _features(st, x, training):
    x = F.conv2d(x, st['resnet18.conv1.weight'], None, 2, 3)          # resnet.py:94  7x7/2 p3
    x = F.relu(_bn(st, 'resnet18.bn1', x, training))                  # :95-96
    x = F.max_pool2d(x, 3, 2, 1)                                      # :97
    for li in range(1, 5):                                            # :99-102
        for blk in range(2):
            p = f'resnet18.layer{li}.{blk}'
            x = _basic_block(st, p, x, training, (p + '.downsample.0.weight') in st)
    return x


def forward(st, x: torch.Tensor, training: bool) -> torch.Tensor:
    """FCN forward.  x: (N,C,96,96) f32 -> (N,A,96,96) f32.  Mutates BN buffers when training."""
    x = _features(st, x, training)                                    # networks.py:17
    x = F.conv2d(x, st['conv1.weight'], st['conv1.bias'])             # :18
    x = F.relu(_bn(st, 'bn1', x, training))                           # :19-20
    x = F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=True)   # :21
    x = F.conv2d(x, st['conv2.weight'], st['conv2.bias'])             # :22
    x = F.relu(_bn(st, 'bn2', x, training))                           # :23-24
    x = F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=True)   # :25
    return F.conv2d(x, st['conv3.weight'], st['conv3.bias'])          # :26


def forward_trace(st, x: torch.Tensor, training: bool) -> Dict[str, torch.Tensor]:
    """Same arithmetic as :func:`forward`, returning the named intermediates the CUDA library can
    export through ``simq_debug_get`` (tests localise a mismatch to one layer with it).  Conv outputs
    of the head are recorded WITHOUT their bias (the library folds the bias into the BN shift)."""
    tr: Dict[str, torch.Tensor] = {}
    h = F.conv2d(x, st['resnet18.conv1.weight'], None, 2, 3)
    tr['raw0'] = h
    h = F.max_pool2d(F.relu(_bn(st, 'resnet18.bn1', h, training)), 3, 2, 1)
    tr['a0'] = h
    b = 0
    for li in range(1, 5):
        for blk in range(2):
            p = f'resnet18.layer{li}.{blk}'
            r1 = F.conv2d(h, st[p + '.conv1.weight'], None, 1, 1)
            b1 = F.relu(_bn(st, p + '.bn1', r1, training))
            r2 = F.conv2d(b1, st[p + '.conv2.weight'], None, 1, 1)
            o = _bn(st, p + '.bn2', r2, training)
            ident = h
            if (p + '.downsample.0.weight') in st:
                rd = F.conv2d(h, st[p + '.downsample.0.weight'])
                tr[f'blk{b}.rawd'] = rd
                ident = _bn(st, p + '.downsample.1', rd, training)
            h = F.relu(o + ident)
            tr[f'blk{b}.raw1'], tr[f'blk{b}.b1'], tr[f'blk{b}.raw2'], tr[f'blk{b}.out'] = r1, b1, r2, h
            b += 1
    r = F.conv2d(h, st['conv1.weight'], None)
    tr['raw_h1'] = r
    h = F.relu(_bn(st, 'bn1', r + st['conv1.bias'].view(1, -1, 1, 1), training))
    h = F.interpolate(h, scale_factor=2, mode='bilinear', align_corners=True)
    tr['u1'] = h
    r = F.conv2d(h, st['conv2.weight'], None)
    tr['raw_h2'] = r
    h = F.relu(_bn(st, 'bn2', r + st['conv2.bias'].view(1, -1, 1, 1), training))
    tr['t'] = F.conv2d(h, st['conv3.weight'], None)
    h = F.interpolate(h, scale_factor=2, mode='bilinear', align_corners=True)
    tr['q'] = F.conv2d(h, st['conv3.weight'], st['conv3.bias'])
    return tr


def hwc_to_nchw(states: Sequence[np.ndarray]) -> torch.Tensor:
    """policies.py:44-45 (ToTensor on float32 HWC ndarray = transpose only, no /255) + train.py:109 cat."""
    return torch.from_numpy(np.ascontiguousarray(np.stack(states).transpose(0, 3, 1, 2)))


def greedy_action(st, state_hwc: np.ndarray) -> Tuple[int, np.ndarray]:
    """policies.py:56-64 with exploration off: eval-mode forward at batch 1, flat first-max argmax."""
    with torch.no_grad():
        o = forward(st, hwc_to_nchw([state_hwc]), False)[0]
    return int(o.view(1, -1).max(1)[1].item()), o.numpy()


# --------------------------------------------------------------------------------------
# the DQN update (train.py:108-141)
# --------------------------------------------------------------------------------------
def dqn_step(policy, target, momentum: Optional[Dict[str, torch.Tensor]],
             state: torch.Tensor, action: torch.Tensor, reward: torch.Tensor,
             next_state: torch.Tensor, non_final_mask: torch.Tensor, *,
             discount: float, lr: float = 0.01, mom: float = 0.9, weight_decay: float = 1e-4,
             grad_clip: Optional[float] = 100.0, double_dqn: bool = True, apply_update: bool = True):
    """One update.  ``policy``/``target``: state dicts (policy mutated in place: BN buffers,
    parameters).  ``state`` (B,C,96,96); ``next_state`` (Bn,C,96,96) holds only the non-terminal
    rows in order (train.py:112); ``momentum``: dict of SGD momentum buffers or None for the first
    step (torch.optim.SGD: buf = g on first step).  Returns dict with loss, td_error, grads,
    grad_norm, q (online Q-map on ``state``), best_action, momentum."""
    B = state.shape[0]
    names = [n for n in policy if policy[n].is_floating_point() and policy[n].dim() >= 1
             and not n.endswith(('running_mean', 'running_var')) and not n.startswith('resnet18.fc.')]
    leaves = {}
    work = OrderedDict(policy)            # shallow: buffers shared (mutated), params replaced by leaves
    for n in names:
        leaves[n] = policy[n].detach().clone().requires_grad_(True)
        work[n] = leaves[n]

    output = forward(work, state, True)                                            # :114
    q_sa = output.view(B, -1).gather(1, action.view(B, 1)).squeeze(1)              # :115
    next_v = torch.zeros(B, dtype=state.dtype, device=state.device)                # :116
    best = torch.zeros(0, dtype=torch.long, device=state.device)
    with torch.no_grad():
        if next_state.shape[0] > 0:
            if double_dqn:                                                         # :119-122
                best = forward(work, next_state, True).view(next_state.shape[0], -1).max(1)[1]
                next_v[non_final_mask] = forward(target, next_state, False).view(
                    next_state.shape[0], -1).gather(1, best.view(-1, 1)).view(-1)
            else:                                                                  # :124
                next_v[non_final_mask] = forward(target, next_state, False).view(
                    next_state.shape[0], -1).max(1)[0]
    y = reward + discount * next_v                                                 # :126
    td = (q_sa - y).abs().detach()                                                 # :127
    loss = F.smooth_l1_loss(q_sa, y)                                               # :129
    grads_t = torch.autograd.grad(loss, [leaves[n] for n in names])                # :131-132
    grads = OrderedDict((n, g.detach().clone()) for n, g in zip(names, grads_t))

    total_norm = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).to(state.dtype)
    if grad_clip is not None:                                                      # :133-134
        coef = min(1.0, float(grad_clip) / (float(total_norm) + 1e-6))
        if coef < 1.0:
            for g in grads.values():
                g.mul_(coef)
    new_mom = OrderedDict()
    if apply_update:                                                               # :135, ctor :186
        for n in names:
            g = grads[n] + weight_decay * policy[n]
            buf = g.clone() if momentum is None else momentum[n] * mom + g
            new_mom[n] = buf
            policy[n] = policy[n] - lr * buf
    # BN buffers were mutated through the shared tensors in ``work``; nbt are 0-d tensors (in place)
    return {'loss': float(loss.item()), 'td_error': float(td.mean().item()), 'grads': grads,
            'grad_norm': float(total_norm), 'q': output.detach(), 'best_action': best,
            'momentum': new_mom, 'q_sa': q_sa.detach(), 'target_y': y}


def intention_step(net, momentum: Optional[Dict[str, torch.Tensor]], states_hwc: Sequence[np.ndarray], *,
                   lr: float = 0.01, mom: float = 0.9, weight_decay: float = 1e-4, apply_update: bool = True):
    """train.train_intention (train.py:143-158): ``net`` = state dict of FCN(C-1, 1) (mutated); ``states_hwc``:
    (96,96,C) float32 arrays whose LAST channel is the ground-truth intention map.  BCEWithLogitsLoss
    (mean), no gradient clipping, SGD(momentum 0.9, wd) as constructed at train.py:190."""
    x = hwc_to_nchw([s[:, :, :-1] for s in states_hwc])                            # :145
    target = hwc_to_nchw([s[:, :, -1:] for s in states_hwc])                       # :146
    names = [n for n in net if net[n].is_floating_point() and net[n].dim() >= 1
             and not n.endswith(('running_mean', 'running_var')) and not n.startswith('resnet18.fc.')]
    leaves = {n: net[n].detach().clone().requires_grad_(True) for n in names}
    work = OrderedDict(net)
    work.update(leaves)
    output = forward(work, x.to(net[names[0]].dtype), True)                         # :148
    loss = F.binary_cross_entropy_with_logits(output, target.to(output.dtype))     # :149-150
    grads_t = torch.autograd.grad(loss, [leaves[n] for n in names])                # :151-152
    grads = OrderedDict((n, g.detach().clone()) for n, g in zip(names, grads_t))
    new_mom = OrderedDict()
    if apply_update:                                                               # :153
        for n in names:
            g = grads[n] + weight_decay * net[n]
            buf = g.clone() if momentum is None else momentum[n] * mom + g
            new_mom[n] = buf
            net[n] = net[n] - lr * buf
    return {'loss_intention': float(loss.item()), 'grads': grads, 'momentum': new_mom, 'output': output.detach()}


# --------------------------------------------------------------------------------------
# digests used by the golden fixtures (small, order-robust summaries of big tensors)
# --------------------------------------------------------------------------------------
def digest(t: torch.Tensor, k: int = 8) -> np.ndarray:
    """[sum, sum|x|, l2, then k samples at fixed pseudo-random flat indices] in float64."""
    a = t.detach().double().reshape(-1).numpy()
    idx = (np.arange(k, dtype=np.int64) * 2654435761 + 12345) % max(a.size, 1)
    return np.concatenate([[a.sum(), np.abs(a).sum(), math.sqrt(float((a * a).sum()))], a[idx]])
