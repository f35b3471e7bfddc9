"""Drives the staged reference (``oracle/_ref/*.bc``, see ``oracle/build_ref.py``) for bench.py's reference arm and same-GPU
comparator.  TEST / MEASUREMENT INFRASTRUCTURE ONLY.

``train.train`` (train.py:108-141), ``policies.DQNPolicy`` (policies.py:11-74) and ``networks.FCN`` (networks.py:6-26) run
UNMODIFIED; only their unavailable imports are stubbed: ``envs.VectorEnv`` (static surface of envs.py:366-376, :810, :1090 ->
pybullet is not installed) and ``utils`` (unused by ``train.train``).  The reference picks its device itself
(``torch.device('cuda' if torch.cuda.is_available() else 'cpu')``, train.py:24 / policies.py:20): the CPU arm hides the GPUs with
CUDA_VISIBLE_DEVICES before torch is imported, the same-GPU comparator leaves them visible.
"""
from __future__ import annotations

import os
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, '_ref')


def available() -> bool:
    return all(os.path.exists(os.path.join(REF_DIR, m + '.bc')) for m in ('networks', 'resnet', 'policies', 'train'))


_mods = None


def load():
    """(networks, policies, train) of the reference, imported once."""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError('oracle/_ref is empty: run oracle/build_ref.py where /root/reference exists')
    envs = types.ModuleType('envs')

    class VectorEnv:
        @staticmethod
        def get_num_output_channels(robot_type):
            return 1 if robot_type == 'pushing_robot' else 2

        @staticmethod
        def get_action_space(robot_type):
            return VectorEnv.get_num_output_channels(robot_type) * 96 * 96

        @staticmethod
        def get_state_width():
            return 96

    envs.VectorEnv = VectorEnv
    saved = {k: sys.modules.get(k) for k in ('envs', 'utils', 'networks', 'resnet', 'policies', 'train')}
    sys.modules['envs'] = envs
    sys.modules['utils'] = types.ModuleType('utils')
    for k in ('networks', 'resnet', 'policies', 'train'):
        sys.modules.pop(k, None)
    import importlib.machinery
    import importlib.util
    mods = {}
    try:
        for name in ('resnet', 'networks', 'policies', 'train'):          # import order = dependency order (networks.py:4, policies.py:7)
            loader = importlib.machinery.SourcelessFileLoader(name, os.path.join(REF_DIR, name + '.bc'))
            mod = importlib.util.module_from_spec(importlib.util.spec_from_loader(name, loader))
            sys.modules[name] = mod
            loader.exec_module(mod)
            mods[name] = mod
    finally:
        for k, v in saved.items():                # do not leave the reference's top-level names in sys.modules
            if v is not None:
                sys.modules[k] = v
            else:
                sys.modules.pop(k, None)
    _mods = (mods['networks'], mods['policies'], mods['train'])
    return _mods


def make_cfg(C, robot_type, B, clip=100):
    return types.SimpleNamespace(robot_config=[{robot_type: 1}], num_input_channels=C, checkpoint_path=None, policy_path=None,
                                 final_exploration=0.01, batch_size=B, use_double_dqn=True, grad_norm_clipping=clip)


def time_train(C, A, B, gamma, steps, warmup, terminal_every, allow_tf32=None, n_batches=4, seed=1234):
    """Wall-clock of the reference's own ``train.train`` calls (they end in two ``.item()`` syncs), rotating ``n_batches`` distinct
    synthetic minibatches.  Returns dict(value samples/s, ms_per_step, device, loss)."""
    import torch
    ROOT = os.path.dirname(HERE)
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from spatial_intention_maps_b200 import synth
    networks, policies, train = load()
    old = None
    if allow_tf32 is not None:
        old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = bool(allow_tf32)
    try:
        robot = 'pushing_robot' if A == 1 else 'lifting_robot'
        cfg = make_cfg(C, robot, B)
        torch.manual_seed(0)
        policy = policies.DQNPolicy(cfg, train=True)
        net = policy.policy_nets[0]
        target = policy.build_policy_nets()[0]
        target.load_state_dict(net.state_dict())                                                   # train.py:213-216
        target.eval()
        net.train()
        opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)          # train.py:186
        batches = [train.Transition(*synth.synth_batch(B, C, A, seed + i, terminal_every=terminal_every)) for i in range(n_batches)]
        info, t0 = None, 0.0
        for i in range(warmup + steps):
            if i == warmup:
                if policy.device.type == 'cuda':
                    torch.cuda.synchronize()
                t0 = time.perf_counter()
            info = train.train(cfg, net, target, opt, batches[i % n_batches], policy.apply_transform, gamma)
        if policy.device.type == 'cuda':
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return {'value': B * steps / dt, 'unit': 'samples/s', 'ms_per_step': dt / steps * 1e3, 'device': str(policy.device),
                'batch': B, 'steps': steps, 'loss': float(info['loss'])}
    finally:
        if old is not None:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
