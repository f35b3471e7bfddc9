"""Drop-in for the reference's ``train.train`` (train.py:108-141): one Double-DQN update of a
``networks.FCN`` -- same signature, same returned ``{'td_error', 'loss'}`` Python floats -- executed as
ONE call into the simq CUDA library (``simq_train_step``: online forward on s, train-mode online
forward on s', eval-mode target forward, gather / arg-max / TD target / SmoothL1, backward, global
grad-norm clip, momentum-SGD), instead of ~450 eager kernels.

The caller keeps passing its stock ``torch.optim.SGD``: learning rate, momentum and weight decay are
read from ``optimizer.param_groups[0]`` and the momentum buffers live in the optimizer's state as
views of the network's flat momentum vector, so ``optimizer.state_dict()`` checkpoints (train.py:324)
stay valid.

Data parallel: when ``torch.distributed`` is initialised with world_size > 1 every rank runs the step
on its shard of the minibatch up to the gradients, the flat fp32 gradient vector is all-reduced once
(NCCL over NVLink; mean), and every rank applies the identical clip + SGD.  BatchNorm statistics stay
per rank, as under the reference's ``DataParallel`` (policies.py:39).
"""
from __future__ import annotations

from collections import namedtuple

import ctypes as C
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .networks import FCN

Transition = namedtuple('Transition', ('state', 'action', 'reward', 'next_state'))     # train.py:26
OVERLAP_ALLREDUCE = os.environ.get('SIMQ_OVERLAP_ALLREDUCE', '1') != '0'      # data parallel: all-reduce layer 4 + head under the rest of the backward


def _unwrap(net) -> FCN:
    m = getattr(net, 'module', net)
    if not isinstance(m, FCN):
        raise TypeError('spatial_intention_maps_b200.train.train needs spatial_intention_maps_b200.networks.FCN nets')
    return m


class HostBatch:
    """Pinned host staging for one replay minibatch in NHWC (the layout ``Mapper.get_state`` already
    produces, envs.py:2067-2184), replacing the per-sample ``transform_fn`` + ``torch.cat`` of
    train.py:109-112.  Reused across steps to avoid re-pinning."""

    def __init__(self, B: int, C: int):
        self.B, self.C = B, C
        pin = torch.cuda.is_available()
        self.s = torch.empty((B, 96, 96, C), dtype=torch.float32, pin_memory=pin)
        self.ns = torch.empty((B, 96, 96, C), dtype=torch.float32, pin_memory=pin)
        self.action = torch.empty(B, dtype=torch.int64, pin_memory=pin)
        self.reward = torch.empty(B, dtype=torch.float32, pin_memory=pin)
        self.nonfinal = torch.empty(B, dtype=torch.uint8, pin_memory=pin)
        self.Bn = 0

    _pool = None

    def fill(self, batch, after_states=None):
        """Stage one ``Transition`` of tuples.  The 2*B row copies (184 KB each at C=5) are spread over a few threads
        (numpy releases the GIL while copying): ~4x faster than the serial loop at B=128.  States and the small vectors
        are staged first; ``after_states()`` (if given) runs before the next states are staged, so the caller can start
        the host->device copy of s while this thread is still copying s' (``_enqueue_train``)."""
        B = self.B
        if len(batch.state) != B:
            raise ValueError(f'batch has {len(batch.state)} transitions, expected {B}')
        s_np, ns_np = self.s.numpy(), self.ns.numpy()
        nf = np.fromiter((n is not None for n in batch.next_state), dtype=bool, count=B)
        slot = np.cumsum(nf) - 1                     # row of sample i in the compacted next-state batch (train.py:112)

        def copy_states(lo, hi):
            for i in range(lo, hi):
                s_np[i] = batch.state[i]

        def copy_next(lo, hi):
            for i in range(lo, hi):
                if nf[i]:
                    ns_np[slot[i]] = batch.next_state[i]

        workers = min(8, max(1, B // 16))
        if workers > 1 and HostBatch._pool is None:
            from concurrent.futures import ThreadPoolExecutor
            HostBatch._pool = ThreadPoolExecutor(max_workers=8, thread_name_prefix='simq-stage')
        step = (B + workers - 1) // workers

        def run(fn):
            if workers == 1:
                fn(0, B)
            else:
                list(HostBatch._pool.map(lambda lo: fn(lo, min(B, lo + step)), range(0, B, step)))

        run(copy_states)
        self.nonfinal.numpy()[:] = nf
        self.Bn = int(nf.sum())
        self.action.numpy()[:] = np.asarray(batch.action, dtype=np.int64)
        self.reward.numpy()[:] = np.asarray(batch.reward, dtype=np.float32)
        if after_states is not None:
            after_states()
        run(copy_next)
        return self

    def h2d_bytes(self) -> int:
        return (self.B + self.Bn) * 96 * 96 * self.C * 4 + self.B * (8 + 4 + 1)


class DeviceBatch:
    def __init__(self, B: int, C: int, device):
        self.s = torch.empty((B, 96, 96, C), dtype=torch.float32, device=device)
        self.ns = torch.empty((B, 96, 96, C), dtype=torch.float32, device=device)
        self.action = torch.empty(B, dtype=torch.int64, device=device)
        self.reward = torch.empty(B, dtype=torch.float32, device=device)
        self.nonfinal = torch.empty(B, dtype=torch.uint8, device=device)
        self.out2 = torch.zeros(2, dtype=torch.float32, device=device)
        self.out2_host = torch.zeros(2, dtype=torch.float32, pin_memory=torch.cuda.is_available())
        self.Bn = 0
        self.ns_ready = None                     # event of a pending s' upload on the copy stream (consumed by the next step)
        self._copy_stream = self._ev_s = self._ev_ns = None

    def upload(self, hb: HostBatch):
        """Host -> device on the current stream, except the next states: they go second over the same link on a copy
        stream, and only the s' lane of the step waits for them (``simq_set_next_state_event``), so their half of the
        transfer runs under the forward on s."""
        self.upload_states(hb)
        return self.upload_next(hb)

    def upload_states(self, hb: HostBatch):
        """s and the per-sample vectors, on the current stream (``hb.s`` / action / reward / nonfinal must be staged)."""
        self.s.copy_(hb.s, non_blocking=True)
        self.action.copy_(hb.action, non_blocking=True)
        self.reward.copy_(hb.reward, non_blocking=True)
        self.nonfinal.copy_(hb.nonfinal, non_blocking=True)
        self.ns_ready = None
        self.Bn = hb.Bn
        if hb.Bn:
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(self.s.device)
                self._ev_s, self._ev_ns = torch.cuda.Event(), torch.cuda.Event()
            self._ev_s.record(torch.cuda.current_stream(self.s.device))     # after s (the link is shared: s first) and after
        return self                                                        # every earlier reader of self.ns

    def upload_next(self, hb: HostBatch):
        """The compacted next states, on the copy stream, ordered after ``upload_states``."""
        if hb.Bn:
            self._copy_stream.wait_event(self._ev_s)
            with torch.cuda.stream(self._copy_stream):
                self.ns[:hb.Bn].copy_(hb.ns[:hb.Bn], non_blocking=True)
                self._ev_ns.record(self._copy_stream)
            self.ns_ready = self._ev_ns
        return self


def shard_batch(batch, rank: int, world: int):
    """Rank ``rank``'s contiguous slice of a global replay minibatch (``Transition`` of tuples): the
    replay sample is drawn with a shared seed on every rank and split along B, no data-path collective."""
    B = len(batch.state)
    if B % world:
        raise ValueError(f'global batch {B} is not divisible by world size {world}')
    lo, hi = rank * (B // world), (rank + 1) * (B // world)
    return Transition(*(tuple(f[lo:hi]) for f in batch))


def allreduce_mean_(grads_ext: torch.Tensor, world: int):
    """The path's one exchange: average the flat fp32 gradient vector over ranks -> gradient of the global-batch mean
    loss.  ``grads_ext`` = the gradient vector followed by the step's (loss, td_error) report (``FCN.flat_grad_ext``), so the
    report rides in the same collective.  NCCL averages inside the collective (no extra pass over the 45 MB vector);
    back-ends without AVG (gloo, used by the CPU tests) sum and scale."""
    if dist.get_backend() == 'nccl':
        dist.all_reduce(grads_ext, op=dist.ReduceOp.AVG)
        return
    dist.all_reduce(grads_ext)
    grads_ext.mul_(1.0 / world)


def sync_replicas(*nets, src: int = 0):
    """Make every rank's copy of ``nets`` identical to rank ``src``'s: parameters, BatchNorm running statistics and
    counters, momentum.  The data-parallel step only exchanges GRADIENTS, so replicas must start identical (ranks built from
    different RNG states, or one rank loading a checkpoint, would silently train diverged replicas); ``train`` / ``train_intention``
    call this once per network on their first distributed step.  BatchNorm running statistics then evolve per rank (each rank
    sees its own shard, exactly like the replicas of the reference's DataParallel, policies.py:39): call ``sync_replicas`` again
    before checkpointing from, or syncing a target network on, anything but rank ``src`` to make rank ``src``'s statistics the
    job's, as DataParallel does."""
    for net in nets:
        m = _unwrap(net)
        for t in (m.flat_params, m.flat_bn, m.flat_nbt):
            dist.broadcast(t, src)
        has = torch.tensor([0 if m.flat_momentum is None else 1, 1 if m.momentum_initialized else 0], device=m.flat_params.device)
        dist.broadcast(has, src)
        if int(has[0]):
            if m.flat_momentum is None or m.flat_momentum.device != m.flat_params.device:
                m.flat_momentum = torch.zeros_like(m.flat_params)
                m._momentum_bound_to = None
            dist.broadcast(m.flat_momentum, src)
            m.momentum_initialized = bool(int(has[1]))
        m.mark_params_changed()
        m._dp_synced = True


def _momentum_views(net: FCN, optimizer):
    """Make the optimizer's momentum buffers views of the net's flat momentum vector (once)."""
    if net.flat_momentum is None or net.flat_momentum.device != net.flat_params.device:
        net.flat_momentum = torch.zeros_like(net.flat_params)
        net.momentum_initialized = False
    po = net._layout[2]
    if getattr(net, '_momentum_bound_to', None) is optimizer:
        p0 = net._tr_cache[0]                       # still bound?  optimizer.load_state_dict() (resume, train.py:200-210)
        buf0 = optimizer.state[p0].get('momentum_buffer') if p0 in optimizer.state else None   # replaces the buffers
        if buf0 is not None and buf0.data_ptr() == net.flat_momentum.data_ptr() + 4 * po[0]:
            return
    have = 0
    for i, (_, p) in enumerate(net.trainable()):
        view = net.flat_momentum[po[i]:po[i + 1]].view(p.shape)
        st = optimizer.state[p]
        buf = st.get('momentum_buffer')
        if buf is not None and buf.data_ptr() != view.data_ptr():     # resumed checkpoint (train.py:200-210)
            view.copy_(buf)
            have += 1
        st['momentum_buffer'] = view
    if have:
        net.momentum_initialized = True
    net._momentum_bound_to = optimizer


def train_step_device(policy: FCN, target: FCN, optimizer, db: DeviceBatch, B: int, discount_factor: float,
                      grad_norm_clipping, use_double_dqn: bool = True):
    """The device part of one update on an uploaded batch; leaves (loss, td_error) in ``db.out2``."""
    g = optimizer.param_groups[0]
    if g.get('nesterov') or g.get('dampening', 0) != 0:
        raise _lib.SimqError('the fused step implements SGD(momentum, weight_decay) without nesterov/dampening (train.py:186)')
    _momentum_views(policy, optimizer)
    ctx = policy.ctx(B)
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    clip = float(grad_norm_clipping) if grad_norm_clipping is not None else 0.0
    first = 0 if policy.momentum_initialized else 1
    lr, mom, wd = float(g['lr']), float(g.get('momentum', 0.0)), float(g.get('weight_decay', 0.0))
    if world > 1 and not (getattr(policy, '_dp_synced', False) and getattr(target, '_dp_synced', False)):
        sync_replicas(policy, target)
        first = 0 if policy.momentum_initialized else 1
    grads_ext = policy.flat_grad_ext()
    grads = policy.flat_grad()
    out2 = db.out2 if world == 1 else grads_ext[grads.numel():grads.numel() + 2]     # DP: the report rides behind the gradients
    L = _lib.lib()
    ns_ready, db.ns_ready = getattr(db, 'ns_ready', None), None
    if ns_ready is not None:                 # s' is still on its way on the copy stream (DeviceBatch.upload)
        _lib.check(L.simq_set_next_state_event(ctx.handle, C.c_void_p(ns_ready.cuda_event)), 'simq_set_next_state_event')
    def step(phase):
        _lib.check(L.simq_train_step_phase(
            ctx.handle, _lib.ptr(policy.flat_params), _lib.ptr(policy.flat_bn), _lib.ptr(policy.flat_nbt),
            _lib.ptr(target.flat_params), _lib.ptr(target.flat_bn), target.params_version, _lib.ptr(grads),
            _lib.ptr(policy.flat_momentum), _lib.ptr(db.s), _lib.ptr(db.ns), _lib.X_NHWC, _lib.ptr(db.action),
            _lib.ptr(db.reward), _lib.ptr(db.nonfinal), B, db.Bn, float(discount_factor), lr, mom, wd, clip, first,
            1 if use_double_dqn else 0, 1 if world == 1 else 0, _lib.ptr(out2), phase, _lib.stream_ptr()), 'simq_train_step')

    if world == 1:
        step(0)
    else:
        if OVERLAP_ALLREDUCE and dist.get_backend() == 'nccl':
            # the backward retires layer 4 + head first: their gradients (75 % of the vector, and the report tail behind it) are
            # all-reduced on a communication stream while phase 2 still computes the gradients of layers 3..1 and the stem
            split = policy.grad_bucket_split()
            comm = policy.__dict__.get('_comm_stream')
            if comm is None:
                comm = policy.__dict__['_comm_stream'] = torch.cuda.Stream(policy.flat_params.device)
                policy.__dict__['_comm_event'] = torch.cuda.Event()
            step(1)
            ev = policy.__dict__['_comm_event']
            ev.record()
            comm.wait_event(ev)
            with torch.cuda.stream(comm):
                dist.all_reduce(grads_ext[split:], op=dist.ReduceOp.AVG)
            step(2)
            dist.all_reduce(grads_ext[:split], op=dist.ReduceOp.AVG)
            torch.cuda.current_stream().wait_stream(comm)
        else:
            step(0)
            allreduce_mean_(grads_ext, world)
        _lib.check(L.simq_sgd_step(ctx.handle, _lib.ptr(policy.flat_params), _lib.ptr(grads), _lib.ptr(policy.flat_momentum),
                                   lr, mom, wd, clip, first, None, _lib.stream_ptr()), 'simq_sgd_step')
        db.out2.copy_(out2, non_blocking=True)
    policy.momentum_initialized = True
    policy._manual_version += 1


def _enqueue_train(cfg, policy_net, target_net, optimizer, batch, discount_factor) -> DeviceBatch:
    """Stage ``batch`` and enqueue one update on the current stream; (loss, td_error) land in the returned
    DeviceBatch's pinned ``out2_host`` once the stream has run (no synchronisation here)."""
    policy, target = _unwrap(policy_net), _unwrap(target_net)
    B = int(cfg.batch_size)
    C = policy.num_input_channels
    dev = policy.flat_params.device
    policy.ctx(B)                                # fails loudly (SimqError) off a B200: there is no CPU path
    cache = policy.__dict__.setdefault('_batch_cache', {})
    if cache.get('key') != (B, C, dev):
        cache.clear()
        cache.update(key=(B, C, dev), host=HostBatch(B, C), dev=DeviceBatch(B, C, dev))
    hb, db = cache['host'], cache['dev']
    from .replay import DeviceSample
    if isinstance(batch, DeviceSample):          # device-resident replay buffer: gather, no host staging of the states
        if len(batch) != B:
            raise ValueError(f'batch has {len(batch)} transitions, expected {B}')
        action, reward, nonfinal, Bn = batch.buffer.gather(batch, db.s, db.ns)
        hb.action.numpy()[:] = action; hb.reward.numpy()[:] = reward; hb.nonfinal.numpy()[:] = nonfinal
        db.action.copy_(hb.action, non_blocking=True)
        db.reward.copy_(hb.reward, non_blocking=True)
        db.nonfinal.copy_(hb.nonfinal, non_blocking=True)
        db.Bn = hb.Bn = Bn
    else:
        hb.fill(batch, after_states=lambda: db.upload_states(hb))    # s is on its way while s' is still being staged
        db.upload_next(hb)
    train_step_device(policy, target, optimizer, db, B, discount_factor, getattr(cfg, 'grad_norm_clipping', None),
                      bool(getattr(cfg, 'use_double_dqn', True)))
    db.out2_host.copy_(db.out2, non_blocking=True)
    return db


def train(cfg, policy_net, target_net, optimizer, batch, transform_fn, discount_factor):
    """train.py:108-141 with the same arguments.  ``transform_fn`` is accepted for signature
    compatibility; states are staged NHWC (``ToTensor`` on float32 input is only a transpose,
    policies.py:44-45, which the stem kernel absorbs)."""
    db = _enqueue_train(cfg, policy_net, target_net, optimizer, batch, discount_factor)
    torch.cuda.current_stream().synchronize()       # the reference syncs here too: two .item() calls (train.py:138-139)
    loss = float(db.out2_host[0])
    if loss != loss:                                # NaN: an input error only the device could see (e.g. an action out of range)?
        _lib.check(_lib.lib().simq_check_device_errors(_unwrap(policy_net).ctx().handle, _lib.stream_ptr()), 'simq_train_step')
    return {'td_error': float(db.out2_host[1]), 'loss': loss}


def train_groups(cfg, policy, target_nets, optimizers, batches, optimizers_intention=None):
    """The per-timestep training block of the reference's main loop (train.py:253-263; the same loop is
    ``Trainer.step``, train_multiprocess.py:366-377): one update per robot group -- ``policy.policy_nets[i]`` against
    ``target_nets[i]`` on ``batches[i]`` with ``cfg.discount_factors[i]``, plus the group's intention net when
    ``cfg.use_predicted_intention`` -- returning the reference's ``all_train_info`` dict
    (``'{name}/robot_group_{i+1:02}'`` -> float).  The updates of all groups are enqueued back to back (each network has
    its own context and workspace) and the host waits ONCE, instead of two ``.item()`` round trips per group."""
    n = policy.num_robot_groups
    if not (len(target_nets) == len(optimizers) == len(batches) == n):
        raise ValueError(f'expected {n} target nets / optimizers / batches')
    use_int = bool(getattr(cfg, 'use_predicted_intention', False))
    if use_int and (optimizers_intention is None or len(optimizers_intention) != n):
        raise ValueError('cfg.use_predicted_intention needs one intention optimizer per robot group')
    pending = []
    for i in range(n):
        db = _enqueue_train(cfg, policy.policy_nets[i], target_nets[i], optimizers[i], batches[i], cfg.discount_factors[i])
        ic = _enqueue_intention(policy.intention_nets[i], optimizers_intention[i], _as_transition(batches[i])) if use_int else None
        pending.append((db, ic))
    torch.cuda.current_stream().synchronize()
    info = {}
    for i, (db, ic) in enumerate(pending):
        tag = 'robot_group_{:02}'.format(i + 1)
        info[f'td_error/{tag}'] = float(db.out2_host[1])
        info[f'loss/{tag}'] = float(db.out2_host[0])
        if ic is not None:
            info[f'loss_intention/{tag}'] = float(ic['out_host'][0])
    return info


def _as_transition(batch):
    from .replay import DeviceSample
    return batch.to_transition() if isinstance(batch, DeviceSample) else batch


def _enqueue_intention(intention_net, optimizer, batch):
    net = _unwrap(intention_net)
    B = len(batch.state)
    Ct = net.num_input_channels + 1
    dev = net.flat_params.device
    cache = net.__dict__.setdefault('_intention_cache', {})
    if cache.get('key') != (B, Ct, dev):
        pin = torch.cuda.is_available()
        cache.clear()
        cache.update(key=(B, Ct, dev), host=torch.empty((B, 96, 96, Ct), dtype=torch.float32, pin_memory=pin),
                     dev=torch.empty((B, 96, 96, Ct), dtype=torch.float32, device=dev),
                     out=torch.zeros(1, dtype=torch.float32, device=dev),
                     out_host=torch.zeros(1, dtype=torch.float32, pin_memory=pin))
    h = cache['host'].numpy()
    for i in range(B):
        if batch.state[i].shape != (96, 96, Ct):
            raise ValueError(f'state {i} has shape {batch.state[i].shape}, expected (96, 96, {Ct})')
        h[i] = batch.state[i]
    cache['dev'].copy_(cache['host'], non_blocking=True)
    intention_step_device(net, optimizer, cache['dev'], B, cache['out'])
    cache['out_host'].copy_(cache['out'], non_blocking=True)
    return cache


def train_intention(intention_net, optimizer, batch, transform_fn):
    """train.py:143-158 with the same arguments: one supervised update of the intention-prediction net
    ``FCN(C-1, 1)`` -- input = all channels of ``batch.state`` but the last, target = the last channel,
    ``BCEWithLogitsLoss`` (mean), backward, plain momentum-SGD (no clipping) -- as ONE call into the library
    (``simq_intention_step``).  Returns ``{'loss_intention': float}``."""
    cache = _enqueue_intention(intention_net, optimizer, batch)
    torch.cuda.current_stream().synchronize()
    return {'loss_intention': float(cache['out_host'][0])}


def intention_step_device(net: FCN, optimizer, state_dev: torch.Tensor, B: int, out1: torch.Tensor):
    g = optimizer.param_groups[0]
    if g.get('nesterov') or g.get('dampening', 0) != 0:
        raise _lib.SimqError('the fused step implements SGD(momentum, weight_decay) without nesterov/dampening (train.py:190)')
    _momentum_views(net, optimizer)
    ctx = net.ctx(B)
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    first = 0 if net.momentum_initialized else 1
    lr, mom, wd = float(g['lr']), float(g.get('momentum', 0.0)), float(g.get('weight_decay', 0.0))
    if world > 1 and not getattr(net, '_dp_synced', False):
        sync_replicas(net)
        first = 0 if net.momentum_initialized else 1
    grads_ext = net.flat_grad_ext()
    grads = net.flat_grad()
    rep = out1 if world == 1 else grads_ext[grads.numel():grads.numel() + 1]
    L = _lib.lib()
    _lib.check(L.simq_intention_step(ctx.handle, _lib.ptr(net.flat_params), _lib.ptr(net.flat_bn), _lib.ptr(net.flat_nbt),
                                     _lib.ptr(grads), _lib.ptr(net.flat_momentum), _lib.ptr(state_dev), B, lr, mom, wd, 0.0, first,
                                     1 if world == 1 else 0, _lib.ptr(rep), _lib.stream_ptr()), 'simq_intention_step')
    if world > 1:
        allreduce_mean_(grads_ext, world)
        _lib.check(L.simq_sgd_step(ctx.handle, _lib.ptr(net.flat_params), _lib.ptr(grads), _lib.ptr(net.flat_momentum),
                                   lr, mom, wd, 0.0, first, None, _lib.stream_ptr()), 'simq_sgd_step')
        out1.copy_(rep, non_blocking=True)
    net.momentum_initialized = True
    net._manual_version += 1
