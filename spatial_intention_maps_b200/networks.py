"""Drop-in for the reference's ``networks.FCN`` (networks.py:6-26 over resnet.py:50-120), computed by
the simq CUDA library (``include/simq.h``) instead of ``torch.nn`` layers.

Same constructor, same parameter / buffer names and shapes (so ``state_dict()`` /
``load_state_dict()`` round-trip reference checkpoints, incl. the never-executed
``resnet18.fc.*``), same ``forward(x: (N,C,96,96) f32) -> (N,A,96,96) f32`` with BatchNorm behaviour
following ``.training``, differentiable through a ``torch.autograd.Function`` so a stock
``torch.optim.SGD`` over ``net.parameters()`` works.

All 70 trainable tensors are views into ONE flat fp32 device buffer (``flat_params``), BN running
statistics into ``flat_bn`` and the 22 ``num_batches_tracked`` counters into ``flat_nbt`` -- the
layout ``simq_layout`` reports -- so the fused step (``train.train``) hands the library three
pointers.  There is no PyTorch fallback: without the CUDA library / a B200 ``forward`` raises.
"""
from __future__ import annotations

import copy
import math
from typing import List, Tuple

import torch
import torch.nn as nn

from . import _lib

STAGE_PLANES = (64, 128, 256, 512)
DEFAULT_MAX_BATCH = 32


class _P(nn.Module):
    """A named bag of parameters / buffers (stands in for Conv2d / BatchNorm2d / Linear)."""


def _conv(shape, bias=False):
    m = _P()
    m.weight = nn.Parameter(torch.empty(shape))
    if bias:
        m.bias = nn.Parameter(torch.empty(shape[0]))
    return m


def _bn(ch):
    m = _P()
    m.weight = nn.Parameter(torch.ones(ch))
    m.bias = nn.Parameter(torch.zeros(ch))
    m.register_buffer('running_mean', torch.zeros(ch))
    m.register_buffer('running_var', torch.ones(ch))
    m.register_buffer('num_batches_tracked', torch.tensor(0, dtype=torch.long))
    return m


def _block(inpl, planes, downsample):          # resnet.py:19-29
    m = _P()
    m.conv1 = _conv((planes, inpl, 3, 3)); m.bn1 = _bn(planes)
    m.conv2 = _conv((planes, planes, 3, 3)); m.bn2 = _bn(planes)
    if downsample:                              # resnet.py:79-83
        m.downsample = nn.Sequential(_conv((planes, inpl, 1, 1)), _bn(planes))
    return m


def _resnet18(num_input_channels):             # resnet.py:52-68
    m = _P()
    m.conv1 = _conv((64, num_input_channels, 7, 7)); m.bn1 = _bn(64)
    inpl = 64
    for li, planes in enumerate(STAGE_PLANES, start=1):
        setattr(m, f'layer{li}', nn.Sequential(_block(inpl, planes, inpl != planes), _block(planes, planes, False)))
        inpl = planes
    m.fc = _P()
    m.fc.weight = nn.Parameter(torch.empty(1000, 512))
    m.fc.bias = nn.Parameter(torch.empty(1000))
    return m


class _FCNFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, net, x, *params):
        q, token = net._run_forward(x, training=net.training, save=True)
        ctx.net, ctx.token, ctx.x = net, token, x
        return q

    @staticmethod
    def backward(ctx, dq):
        views = ctx.net._run_backward(ctx.x, dq.contiguous(), ctx.token)
        return (None, None) + tuple(views)


class FCN(nn.Module):
    def __init__(self, num_input_channels=3, num_output_channels=1, max_batch: int = DEFAULT_MAX_BATCH):
        super().__init__()
        self.num_input_channels, self.num_output_channels = num_input_channels, num_output_channels
        self.resnet18 = _resnet18(num_input_channels)
        self.conv1 = _conv((128, 512, 1, 1), bias=True); self.bn1 = _bn(128)
        self.conv2 = _conv((32, 128, 1, 1), bias=True); self.bn2 = _bn(32)
        self.conv3 = _conv((num_output_channels, 32, 1, 1), bias=True)
        self._init_like_reference()
        self.max_batch = max_batch
        self._ctx = None
        self._token = 0
        self._manual_version = 0
        self._flat_grad = None
        self.flat_momentum = None
        self.momentum_initialized = False
        self._layout = None
        self._flatten()

    # ---- initialisation: resnet.py:70-75 + torch defaults for the head convs / fc ----
    def _init_like_reference(self):
        for name, p in self.named_parameters():
            if p.dim() == 4 and name.startswith('resnet18.'):
                nn.init.kaiming_normal_(p, mode='fan_out', nonlinearity='relu')
            elif p.dim() >= 2:                  # head convs, fc: Conv2d / Linear default
                nn.init.kaiming_uniform_(p, a=math.sqrt(5))
        for conv in (self.conv1, self.conv2, self.conv3, self.resnet18.fc):
            fan_in = conv.weight[0].numel()
            nn.init.uniform_(conv.bias, -1 / math.sqrt(fan_in), 1 / math.sqrt(fan_in))

    # ---- flat storage ----
    def trainable(self) -> List[Tuple[str, nn.Parameter]]:
        return [(n, p) for n, p in self.named_parameters() if not n.startswith('resnet18.fc.')]

    def _flatten(self):
        """(Re)build the flat buffers on the parameters' current device and re-point every
        parameter / buffer at its slice.  Called at construction and after ``.to()/.cuda()``."""
        tr = self.trainable()
        dev = tr[0][1].device
        if self._layout is None:
            n_p, n_b, po, bo = (sum(p.numel() for _, p in tr), None, None, None)
            try:
                n_p, n_b, po, bo = _lib.layout(self.num_input_channels, self.num_output_channels)
            except _lib.SimqError:
                po = [0]
                for _, p in tr:
                    po.append(po[-1] + p.numel())
                bo = [0]
                for _, m in self._bns():
                    bo.append(bo[-1] + 2 * m.weight.numel())
                n_b = bo[-1]
            assert len(po) == len(tr) + 1 and all(po[i + 1] - po[i] == p.numel() for i, (_, p) in enumerate(tr)), \
                'simq_layout disagrees with the module definition'
            self._layout = (n_p, n_b, po, bo)
        n_p, n_b, po, bo = self._layout
        flat = torch.empty(n_p, dtype=torch.float32, device=dev)
        for i, (_, p) in enumerate(tr):
            flat[po[i]:po[i + 1]].copy_(p.data.reshape(-1))
            p.data = flat[po[i]:po[i + 1]].view(p.shape)
        bns = self._bns()
        fbn = torch.empty(n_b, dtype=torch.float32, device=dev)
        nbt = torch.empty(len(bns), dtype=torch.int64, device=dev)
        for i, (_, m) in enumerate(bns):
            ch = m.weight.numel()
            fbn[bo[i]:bo[i] + ch].copy_(m.running_mean)
            fbn[bo[i] + ch:bo[i] + 2 * ch].copy_(m.running_var)
            nbt[i] = m.num_batches_tracked
            m._buffers['running_mean'] = fbn[bo[i]:bo[i] + ch]
            m._buffers['running_var'] = fbn[bo[i] + ch:bo[i] + 2 * ch]
            m._buffers['num_batches_tracked'] = nbt[i]
        self.flat_params, self.flat_bn, self.flat_nbt = flat, fbn, nbt
        self._tr_cache = [p for _, p in tr]
        self._manual_version += 1
        self._flat_grad = None
        if self.flat_momentum is not None:
            self.flat_momentum = self.flat_momentum.to(dev)
        if self._ctx is not None:
            self._ctx.close()
            self._ctx = None

    def _bns(self):
        return [(n, m) for n, m in self.named_modules() if isinstance(m, _P) and 'running_mean' in m._buffers]

    # per-object device state that a copy must NOT share: the library context (one ctypes handle = one workspace), gradient /
    # staging buffers, and the flat vectors themselves (rebuilt by _flatten from the copied parameters)
    _NO_COPY = ('_ctx', '_flat_grad', '_flat_grad_ext', '_batch_cache', '_intention_cache', '_ga', '_saved_x', '_tr_cache', 'flat_params',
                'flat_bn', 'flat_nbt', 'flat_momentum', '_momentum_bound_to', '_dp_synced')

    def __deepcopy__(self, memo):
        """``copy.deepcopy(net)`` -- the common target-network idiom.  ``Parameter.__deepcopy__`` clones every tensor, which would
        leave the copy's parameters detached from any flat vector (the library would read stale buffers) and share this
        object's context handle; so: copy the module tree, then re-alias the copy into flat vectors and a context of its own."""
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k not in FCN._NO_COPY:
                new.__dict__[k] = copy.deepcopy(v, memo)
        new._ctx, new._flat_grad, new._layout = None, None, self._layout
        new.flat_momentum = self.flat_momentum.clone() if self.flat_momentum is not None else None
        new._flatten()
        return new

    def mark_params_changed(self):
        """Call after writing parameters through ``p.data`` / ``flat_params`` views in ways autograd's version counters do not
        see (e.g. a Polyak target update ``target_p.data.copy_(...)``, which bumps no ``_version``): the next forward re-packs the
        tensor-core weight shadows.  ``load_state_dict``, optimizer steps and the fused ``train.train`` need no such call."""
        self._manual_version += 1

    def _apply(self, fn, *a, **k):
        super()._apply(fn, *a, **k)
        self._flatten()
        return self

    @property
    def params_version(self) -> int:
        """Changes whenever any parameter may have changed (in-place torch updates bump the tensors'
        version counters; the fused step bumps ``_manual_version``) -> the library re-packs weights."""
        return 1 + self._manual_version + int(self.flat_params._version) + sum(int(p._version) for p in self._tr_cache)

    # ---- device context ----
    def ctx(self, batch: int = 1) -> _lib.Ctx:
        dev = self.flat_params.device
        if dev.type != 'cuda':
            raise _lib.SimqError('spatial_intention_maps_b200.networks.FCN runs only on a CUDA (B200) device; '
                                 f'parameters are on {dev}. There is no CPU fallback.')
        if self._ctx is None or self._ctx.max_batch < batch:
            if self._ctx is not None:
                torch.cuda.synchronize(dev)
                self._ctx.close()
            self.max_batch = max(self.max_batch, batch)
            self._ctx = _lib.Ctx(dev.index if dev.index is not None else torch.cuda.current_device(),
                                 self.num_input_channels, self.num_output_channels, self.max_batch)
            if getattr(self, '_backend', None) is not None:          # settings survive a workspace re-allocation
                _lib.check(_lib.lib().simq_set_backend(self._ctx.handle, self._backend), 'simq_set_backend')
            if getattr(self, '_precision', None) is not None:
                _lib.check(_lib.lib().simq_set_precision(self._ctx.handle, self._precision), 'simq_set_precision')
            if getattr(self, '_schedule', None) is not None:
                _lib.check(_lib.lib().simq_set_schedule(self._ctx.handle, self._schedule), 'simq_set_schedule')
            if getattr(self, '_bwd_terms', None) is not None:
                _lib.check(_lib.lib().simq_set_backward_terms(self._ctx.handle, *self._bwd_terms), 'simq_set_backward_terms')
        return self._ctx

    def set_backend(self, backend: int):
        self._backend = backend
        _lib.check(_lib.lib().simq_set_backend(self.ctx().handle, backend), 'simq_set_backend')

    def set_precision(self, mode: str):
        """'parity' (default: split-bf16 operands, 3 MMAs per product, meets the reference-parity bar) or 'bf16'
        (one MMA per product: ~2.5x faster conv kernels, Q-map error ~1e-2 -- does NOT meet the parity bar)."""
        m = {'parity': _lib.PRECISION_PARITY, 'bf16': _lib.PRECISION_BF16}[mode]
        self._precision = m
        _lib.check(_lib.lib().simq_set_precision(self.ctx().handle, m), 'simq_set_precision')

    def set_backward_terms(self, dgrad: int = 3, wgrad: int = 3, dgrad2_min_planes: int = 0):
        """Operand terms of the backward GEMMs (3 = the forward's split-bf16 scheme; 2 = the output gradient contributes its bf16
        hi plane only: 2 MMAs per product).  ``dgrad`` applies to the residual blocks with at least ``dgrad2_min_planes`` planes.
        The forward passes -- Q-map and arg-max parity -- are unaffected."""
        self._bwd_terms = (int(dgrad), int(wgrad), int(dgrad2_min_planes))
        _lib.check(_lib.lib().simq_set_backward_terms(self.ctx().handle, *self._bwd_terms), 'simq_set_backward_terms')

    def set_schedule(self, mode: str):
        """'lanes' (default: independent pieces of a step on two streams / graph branches) or 'serial' (one stream, the
        reference's order).  Bit-identical results; 'serial' exists for A/B timing and debugging."""
        m = {'serial': _lib.SCHEDULE_SERIAL, 'lanes': _lib.SCHEDULE_LANES}[mode]
        self._schedule = m
        _lib.check(_lib.lib().simq_set_schedule(self.ctx().handle, m), 'simq_set_schedule')

    @staticmethod
    def _x_layout(x):
        if x.dim() != 4 or x.dtype != torch.float32:
            raise ValueError(f'expected a float32 (N,C,96,96) tensor, got {tuple(x.shape)} {x.dtype}')
        if x.is_contiguous():
            return x, _lib.X_NCHW
        if x.is_contiguous(memory_format=torch.channels_last):
            return x, _lib.X_NHWC          # same bytes as an (N,96,96,C) array: no transpose pass
        return x.contiguous(), _lib.X_NCHW

    def _run_forward(self, x, training: bool, save: bool):
        if x.shape[1] != self.num_input_channels or x.shape[2] != 96 or x.shape[3] != 96:
            raise ValueError(f'expected (N,{self.num_input_channels},96,96), got {tuple(x.shape)}')
        B = x.shape[0]
        c = self.ctx(B)
        x, lay = self._x_layout(x)
        q = torch.empty((B, self.num_output_channels, 96, 96), dtype=torch.float32, device=x.device)
        if save:
            self._token += 1
        _lib.check(_lib.lib().simq_fcn_forward(
            c.handle, _lib.ptr(self.flat_params), _lib.ptr(self.flat_bn), _lib.ptr(self.flat_nbt), _lib.ptr(x), B, lay,
            1 if training else 0, 1 if save else 0, _lib.ptr(q), self.params_version, _lib.stream_ptr()), 'simq_fcn_forward')
        self._saved_x = x if save else getattr(self, '_saved_x', None)
        return q, self._token

    def _run_backward(self, x, dq, token):
        if token != self._token:
            raise _lib.SimqError('only the most recent grad-enabled forward of this FCN can be differentiated '
                                 '(one saved activation set per network)')
        x, lay = self._x_layout(x)
        tr = self.trainable()
        alias = self._flat_grad is not None and any(
            p.grad is not None and p.grad.untyped_storage().data_ptr() == self._flat_grad.untyped_storage().data_ptr()
            for _, p in tr[:1])
        if self._flat_grad is None or alias:
            self._flat_grad = torch.empty_like(self.flat_params)
        g = self._flat_grad
        _lib.check(_lib.lib().simq_fcn_backward(self.ctx().handle, _lib.ptr(self.flat_params), _lib.ptr(x), lay, _lib.ptr(dq),
                                                x.shape[0], _lib.ptr(g), _lib.stream_ptr()), 'simq_fcn_backward')
        po = self._layout[2]
        return [g[po[i]:po[i + 1]].view(p.shape) for i, (_, p) in enumerate(tr)]

    def flat_grad(self) -> torch.Tensor:
        """The flat fp32 gradient vector (layout of ``flat_params``).  It is the head of a slightly longer buffer whose 4-float
        tail carries the step's (loss, td_error) report: data-parallel steps all-reduce head and tail as ONE collective."""
        if self._flat_grad is None:
            n = self.flat_params.numel()
            self._flat_grad_ext = torch.zeros(n + 4, dtype=torch.float32, device=self.flat_params.device)
            self._flat_grad = self._flat_grad_ext[:n]
        return self._flat_grad

    def grad_bucket_split(self) -> int:
        """Offset of ``resnet18.layer4.0.conv1.weight`` in the flat vectors: everything from there on is final after phase 1
        of ``simq_train_step_phase`` (the backward runs head -> layer 4 -> ... -> stem)."""
        names = [n for n, _ in self.trainable()]
        return int(self._layout[2][names.index('resnet18.layer4.0.conv1.weight')])

    def flat_grad_ext(self) -> torch.Tensor:
        self.flat_grad()
        if getattr(self, '_flat_grad_ext', None) is None or self._flat_grad_ext.data_ptr() != self._flat_grad.data_ptr():
            n = self.flat_params.numel()            # _run_backward replaced the buffer (autograd aliasing): rebuild the pair
            self._flat_grad_ext = torch.zeros(n + 4, dtype=torch.float32, device=self.flat_params.device)
            self._flat_grad = self._flat_grad_ext[:n]
        return self._flat_grad_ext

    def forward(self, x):
        tr = [p for _, p in self.trainable()]
        if torch.is_grad_enabled() and any(p.requires_grad for p in tr):
            if not self.training:
                raise _lib.SimqError('differentiating an eval-mode forward is not supported (the reference never does)')
            return _FCNFunction.apply(self, x, *tr)
        return self._run_forward(x, self.training, save=False)[0]

    def greedy_action(self, x, want_q: bool = False):
        """Eval-mode forward + per-sample flat first-max argmax on the device (policies.py:56-64)."""
        B = x.shape[0]
        c = self.ctx(B)
        x, lay = self._x_layout(x)
        act = torch.empty(B, dtype=torch.int64, device=x.device)
        q = torch.empty((B, self.num_output_channels, 96, 96), dtype=torch.float32, device=x.device) if want_q else None
        _lib.check(_lib.lib().simq_greedy_action(c.handle, _lib.ptr(self.flat_params), _lib.ptr(self.flat_bn), _lib.ptr(x), B, lay,
                                                 _lib.ptr(act), _lib.ptr(q), self.params_version, _lib.stream_ptr()),
                   'simq_greedy_action')
        return act, q


    def greedy_action_hwc(self, state_hwc, want_q: bool = False):
        """One environment step's action for a single (96,96,C) float32 state (policies.py:56-64): pinned staging ->
        persistent device buffers -> eval forward + arg-max replayed as ONE CUDA graph -> the action index (8 bytes)
        back.  Returns (int action, Q-map ndarray (A,96,96) or None)."""
        import numpy as np
        dev = self.flat_params.device
        c = self.ctx(1)
        st = self.__dict__.get('_ga')
        if st is None or st['dev'] != dev:
            Cn, A = self.num_input_channels, self.num_output_channels
            st = {'dev': dev, 'pin': torch.empty((1, 96, 96, Cn), dtype=torch.float32, pin_memory=True),
                  'x': torch.empty((1, 96, 96, Cn), dtype=torch.float32, device=dev),
                  'act': torch.zeros(1, dtype=torch.int64, device=dev),
                  'act_host': torch.zeros(1, dtype=torch.int64, pin_memory=True),
                  'q': torch.empty((1, A, 96, 96), dtype=torch.float32, device=dev)}
            self.__dict__['_ga'] = st
        if state_hwc.shape != tuple(st['pin'].shape[1:]):
            raise ValueError(f'expected a {tuple(st["pin"].shape[1:])} state, got {state_hwc.shape}')
        st['pin'].numpy()[0] = np.asarray(state_hwc, dtype=np.float32)
        st['x'].copy_(st['pin'], non_blocking=True)
        _lib.check(_lib.lib().simq_greedy_action(c.handle, _lib.ptr(self.flat_params), _lib.ptr(self.flat_bn), _lib.ptr(st['x']), 1,
                                                 _lib.X_NHWC, _lib.ptr(st['act']), _lib.ptr(st['q']) if want_q else None,
                                                 self.params_version, _lib.stream_ptr()), 'simq_greedy_action')
        st['act_host'].copy_(st['act'], non_blocking=True)
        q = st['q'][0].cpu().numpy() if want_q else None
        torch.cuda.current_stream().synchronize()
        return int(st['act_host'][0]), q


class SingleDeviceParallel(nn.Module):
    """Stands in for the ``torch.nn.DataParallel`` wrapper of policies.py:39-41: keeps the ``module.``
    state_dict prefix of reference checkpoints, but never replicates -- this framework is one
    process per GPU (gradient all-reduce over NCCL, see train.py)."""

    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, *a, **k):
        return self.module(*a, **k)
