"""Device-resident drop-in for the reference's ``ReplayBuffer`` (train.py:28-45): same ``push`` /
``sample`` / ``__len__`` surface and the same ``random.sample`` index stream (seeded runs draw the same
minibatches), but the (96,96,C) float32 states live in HBM from the moment they are pushed -- one 184 KB
H2D copy per environment step instead of re-uploading 2*B states on every update (46.8 MB at B=128) --
and ``train.train`` assembles the minibatch with a device gather (``simq_gather_rows``).

``sample`` returns a :class:`DeviceSample`; ``train.train`` recognises it.  ``DeviceSample.to_transition()``
materialises the reference's ``Transition`` of tuples (host copies) for any other consumer.  The buffer
pickles to host arrays, so ``torch.save(replay_buffers)`` checkpoints (train.py:331) keep working.
"""
from __future__ import annotations

import random
from collections import namedtuple

import numpy as np
import torch

from . import _lib

Transition = namedtuple('Transition', ('state', 'action', 'reward', 'next_state'))     # train.py:26


class DeviceSample:
    """Indices of one sampled minibatch into a :class:`ReplayBuffer`."""

    def __init__(self, buffer: 'ReplayBuffer', idx):
        self.buffer, self.idx = buffer, list(idx)

    def __len__(self):
        return len(self.idx)

    def to_transition(self) -> Transition:
        b = self.buffer
        st = b.states[self.idx].cpu().numpy()
        ns = b.next_states[self.idx].cpu().numpy()
        return Transition(tuple(st[i] for i in range(len(self.idx))), tuple(int(b.action[i]) for i in self.idx),
                          tuple(float(b.reward[i]) for i in self.idx),
                          tuple(ns[j] if b.nonfinal[i] else None for j, i in enumerate(self.idx)))


class ReplayBuffer:
    def __init__(self, capacity, device=None):
        self.capacity = int(capacity)
        self.device = torch.device(device) if device is not None else torch.device(
            'cuda', torch.cuda.current_device()) if torch.cuda.is_available() else torch.device('cpu')
        self.position = 0
        self.size = 0
        self.states = self.next_states = None            # (capacity,96,96,C) f32 on the device, allocated at first push
        self.action = np.zeros(self.capacity, dtype=np.int64)       # host mirrors of the scalars (tiny)
        self.reward = np.zeros(self.capacity, dtype=np.float32)
        self.nonfinal = np.zeros(self.capacity, dtype=np.uint8)
        self._stage = None

    def _alloc(self, shape):
        self.states = torch.zeros((self.capacity,) + tuple(shape), dtype=torch.float32, device=self.device)
        self.next_states = torch.zeros_like(self.states)
        pin = self.device.type == 'cuda'
        self._stage = [torch.empty((2,) + tuple(shape), dtype=torch.float32, pin_memory=pin) for _ in range(4)]
        self._stage_ev = [None] * 4
        self._stage_i = 0

    def push(self, *args):
        state, action, reward, next_state = Transition(*args)
        if self.states is None:
            self._alloc(np.asarray(state).shape)
        i = self.position
        k = self._stage_i
        self._stage_i = (k + 1) % len(self._stage)
        if self._stage_ev[k] is not None:
            self._stage_ev[k].synchronize()              # the async copy out of this staging slot has finished
        st = self._stage[k]
        st[0].numpy()[...] = state
        if next_state is not None:
            st[1].numpy()[...] = next_state
        self.states[i].copy_(st[0], non_blocking=True)
        if next_state is not None:
            self.next_states[i].copy_(st[1], non_blocking=True)
        if self.device.type == 'cuda':
            self._stage_ev[k] = torch.cuda.Event()
            self._stage_ev[k].record()
        self.action[i], self.reward[i], self.nonfinal[i] = int(action), float(reward), 0 if next_state is None else 1
        self.size = min(self.size + 1, self.capacity)
        self.position = (self.position + 1) % self.capacity

    def sample(self, batch_size) -> DeviceSample:
        # random.sample(list_of_n, k) and random.sample(range(n), k) consume the RNG identically and select the
        # same positions: seeded runs draw the minibatches the reference's buffer would (train.py:41-42)
        return DeviceSample(self, random.sample(range(self.size), batch_size))

    def __len__(self):
        return self.size

    # ---- minibatch assembly on the device (used by train.train) ----
    def gather(self, sample: DeviceSample, out_s: torch.Tensor, out_ns: torch.Tensor):
        """Writes states into ``out_s[:B]`` and the non-terminal next states, compacted in order (train.py:112),
        into ``out_ns[:Bn]``; returns (action int64[B], reward f32[B], nonfinal u8[B], Bn) host arrays."""
        idx = np.asarray(sample.idx, dtype=np.int64)
        nf = self.nonfinal[idx]
        idx_ns = idx[nf != 0]
        row = int(np.prod(self.states.shape[1:]))
        if self.device.type == 'cuda':
            di = torch.from_numpy(np.concatenate([idx, idx_ns])).to(self.device, non_blocking=True)
            L = _lib.lib()
            _lib.check(L.simq_gather_rows(_lib.ptr(self.states), _lib.ptr(di), len(idx), row, _lib.ptr(out_s), _lib.stream_ptr()),
                       'simq_gather_rows')
            if len(idx_ns):
                _lib.check(L.simq_gather_rows(_lib.ptr(self.next_states), _lib.ptr(di[len(idx):]), len(idx_ns), row, _lib.ptr(out_ns),
                                              _lib.stream_ptr()), 'simq_gather_rows')
        else:                                            # host bookkeeping path (CPU tests of the buffer logic)
            out_s[:len(idx)] = self.states[torch.from_numpy(idx)]
            if len(idx_ns):
                out_ns[:len(idx_ns)] = self.next_states[torch.from_numpy(idx_ns)]
        return self.action[idx], self.reward[idx], nf, int(len(idx_ns))

    # ---- checkpoints: pickle as host arrays ----
    def __getstate__(self):
        n = self.size
        return {'capacity': self.capacity, 'position': self.position, 'size': n, 'device': str(self.device),
                'states': None if self.states is None else self.states[:n].cpu().numpy(),
                'next_states': None if self.states is None else self.next_states[:n].cpu().numpy(),
                'action': self.action[:n].copy(), 'reward': self.reward[:n].copy(), 'nonfinal': self.nonfinal[:n].copy()}

    def __setstate__(self, st):
        dev = torch.device(st['device'])
        if dev.type == 'cuda' and not torch.cuda.is_available():
            dev = torch.device('cpu')
        self.__init__(st['capacity'], dev)
        self.position, self.size = st['position'], st['size']
        n = self.size
        if st['states'] is not None:
            self._alloc(st['states'].shape[1:])
            self.states[:n].copy_(torch.from_numpy(st['states']))
            self.next_states[:n].copy_(torch.from_numpy(st['next_states']))
        self.action[:n], self.reward[:n], self.nonfinal[:n] = st['action'], st['reward'], st['nonfinal']
