"""Seeded synthetic replay batches shaped like ``Mapper.get_state`` output (reference
envs.py:2067-2184): (96,96,C) float32 HWC per state.

Channel model: ch0 overhead segmentation map in {0..8}/8 drawn per 4x4 block
(envs.py:1880-1889), ch1 robot map in {0,0.5,1} at ~1 % coverage (envs.py:2258-2262), ch2/ch3
min-subtracted distance ramps scaled 0.25 (envs.py:2212-2215, 2287-2299), ch4 sparse [0,1]
intention ramp (envs.py:2333-2344), further channels sparse point/constant maps
(envs.py:2348-2377).  Actions uniform over A*96*96, rewards clip(N(0,0.5),-1.5,2), every
``terminal_every``-th transition terminal (next_state=None) so shapes are static.
"""
from __future__ import annotations

from collections import namedtuple
from typing import List, Optional

import numpy as np

Transition = namedtuple('Transition', ('state', 'action', 'reward', 'next_state'))  # train.py:26
W = 96


def _ramp(rs, n):
    yy, xx = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(W, dtype=np.float32), indexing='ij')
    cy = rs.uniform(0, W, (n, 1, 1)).astype(np.float32)
    cx = rs.uniform(0, W, (n, 1, 1)).astype(np.float32)
    d = np.sqrt((yy[None] - cy) ** 2 + (xx[None] - cx) ** 2) * (1.0 / 96.0)   # ~metres over a 1 m crop
    d = d * rs.uniform(0.5, 2.4, (n, 1, 1)).astype(np.float32) * 0.25
    return d - d.min(axis=(1, 2), keepdims=True)


def synth_states(n: int, C: int, seed: int, uniform: bool = False) -> np.ndarray:
    """(n,96,96,C) float32 NHWC."""
    rs = np.random.RandomState(seed)
    if uniform:
        return rs.uniform(0, 1, (n, W, W, C)).astype(np.float32)
    x = np.zeros((n, W, W, C), dtype=np.float32)
    seg = rs.randint(0, 9, (n, W // 4, W // 4)).astype(np.float32) / 8.0
    x[..., 0] = np.repeat(np.repeat(seg, 4, axis=1), 4, axis=2)
    if C > 1:
        x[..., 1] = rs.choice(np.array([0, 0.5, 1.0], dtype=np.float32), size=(n, W, W), p=[0.99, 0.005, 0.005])
    for c in (2, 3):
        if C > c:
            x[..., c] = _ramp(rs, n)
    if C > 4:
        m = rs.uniform(0, 1, (n, W, W)) < 0.02
        x[..., 4] = np.where(m, rs.uniform(0, 1, (n, W, W)), 0).astype(np.float32)
    for c in range(5, C):
        m = rs.uniform(0, 1, (n, W, W)) < 0.01
        x[..., c] = np.where(m, 1.0, 0.0).astype(np.float32) * rs.uniform(0.2, 1.0, (n, 1, 1)).astype(np.float32)
    return x


def synth_batch(B: int, C: int, A: int, seed: int, terminal_every: Optional[int] = 64,
                uniform: bool = False) -> Transition:
    """A replay minibatch in the reference's ``Transition(*zip(*transitions))`` form
    (train.py:43-44): tuples of length B; terminal rows have ``next_state is None``."""
    rs = np.random.RandomState(seed + 7919)
    s = synth_states(B, C, seed, uniform)
    ns = synth_states(B, C, seed + 104729, uniform)
    action = rs.randint(0, A * W * W, B)
    reward = np.clip(rs.normal(0, 0.5, B), -1.5, 2.0).astype(np.float32)
    nxt: List[Optional[np.ndarray]] = []
    for i in range(B):
        term = terminal_every is not None and (i % terminal_every) == terminal_every - 1
        nxt.append(None if term else ns[i])
    return Transition(tuple(s[i] for i in range(B)), tuple(int(a) for a in action),
                      tuple(float(r) for r in reward), tuple(nxt))
