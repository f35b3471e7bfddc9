"""Host-side mirror of the reference's policy objects (``policies.DQNPolicy`` policies.py:11-74,
``policies.DQNIntentionPolicy`` policies.py:76-146) on top of the simq CUDA library.

What the callers of the reference use -- and what is kept: ``DQNPolicy(cfg, train=False, random_seed=None)``;
attributes ``policy_nets`` / ``intention_nets`` (one network per robot group, ``state_dict`` keys prefixed
``module.``), ``device``, ``num_robot_groups``, ``robot_group_types``, ``train``; methods
``build_policy_nets()`` (also used by ``train.py:213`` to create the target nets), ``apply_transform(s)``,
``step(state, exploration_eps=None, debug=False)``, ``build_intention_nets()``, ``step_intention(state, debug)``.
``state`` is the environment's nested list ``[[ndarray | None per robot] per group]``; results have the
same nesting.  The Python ``random`` stream is consumed exactly as by the reference (one ``random()`` per
pending robot, plus one ``randrange`` when exploring), so seeded rollouts pick the same exploratory actions.

Differences by design: one process drives one GPU (no ``DataParallel`` replication; the wrapper only keeps
the checkpoint prefix); the greedy action is computed on the device and only its index comes back -- the
Q-map is copied to the host on ``debug=True`` only (the reference copies it on every step, policies.py:66).
"""
from __future__ import annotations

import random

import numpy as np
import torch

from . import networks

_ACTION_CHANNELS = {'pushing_robot': 1}          # envs.py:810; every other robot type has 2 (envs.py:1090)
_MAP = 96                                        # envs.py:2010


class _StaticEnv:
    """Static facts about the environment the policy needs (envs.py:366-376) without importing pybullet."""

    @staticmethod
    def get_num_output_channels(robot_type):
        return _ACTION_CHANNELS.get(robot_type, 2)

    @staticmethod
    def get_action_space(robot_type):
        return _StaticEnv.get_num_output_channels(robot_type) * _MAP * _MAP

    @staticmethod
    def get_state_width():
        return _MAP


try:                                             # inside the reference tree the real class is importable
    from envs import VectorEnv                   # type: ignore
except Exception:                                # pybullet & co. absent
    VectorEnv = _StaticEnv


def _pending(state):
    """(group index, robot index, state array) for every robot that awaits an action."""
    for gi, group in enumerate(state):
        for ri, s in enumerate(group):
            if s is not None:
                yield gi, ri, s


def _like(state):
    return [[None] * len(group) for group in state]


class DQNPolicy:
    def __init__(self, cfg, train=False, random_seed=None, device=None, max_batch=None):
        self.cfg, self.train = cfg, train
        if random_seed is not None:
            random.seed(random_seed)
        self.robot_group_types = [next(iter(group)) for group in cfg.robot_config]
        self.num_robot_groups = len(self.robot_group_types)
        self.device = self._pick_device(device)
        self.max_batch = int(max_batch if max_batch is not None else getattr(cfg, 'batch_size', networks.DEFAULT_MAX_BATCH))
        self.policy_nets = self.build_policy_nets()
        self.policy_checkpoint = None
        if getattr(cfg, 'checkpoint_path', None) is not None:       # resume / evaluate a trained policy
            self.policy_checkpoint = torch.load(cfg.policy_path, map_location=self.device)
            self._restore(self.policy_nets, 'state_dicts')
            print("=> loaded policy '{}'".format(cfg.policy_path))

    @staticmethod
    def _pick_device(device):
        if device is not None:
            return torch.device(device)
        if torch.cuda.is_available():
            return torch.device('cuda', torch.cuda.current_device())
        return torch.device('cpu')

    def _restore(self, nets, key):
        for net, sd in zip(nets, self.policy_checkpoint[key]):
            net.load_state_dict(sd)
            net.train(self.train)

    def _new_net(self, c_in, c_out):
        net = networks.FCN(num_input_channels=c_in, num_output_channels=c_out, max_batch=self.max_batch)
        return networks.SingleDeviceParallel(net).to(self.device)

    def build_policy_nets(self):
        return [self._new_net(self.cfg.num_input_channels, VectorEnv.get_num_output_channels(t)) for t in self.robot_group_types]

    def apply_transform(self, s):
        """What torchvision's ``ToTensor`` does to a float32 HWC array (policies.py:20, 44-45): HWC -> CHW without
        rescaling, plus a leading batch axis."""
        a = np.asarray(s)
        if a.ndim == 2:
            a = a[:, :, None]
        return torch.from_numpy(np.ascontiguousarray(np.moveaxis(a, 2, 0))).unsqueeze(0)

    def _act(self, gi, s, eps, want_q):
        """One robot: epsilon-greedy over the flat (A*96*96) action space; returns (action, Q-map or None)."""
        fcn = self.policy_nets[gi].module
        if random.random() < eps:
            action = random.randrange(VectorEnv.get_action_space(self.robot_group_types[gi]))
            return action, (fcn.greedy_action_hwc(s, want_q=True)[1] if want_q else None)
        return fcn.greedy_action_hwc(s, want_q=want_q)

    def step(self, state, exploration_eps=None, debug=False):
        eps = self.cfg.final_exploration if exploration_eps is None else exploration_eps
        action, output = _like(state), _like(state)
        touched = set()
        with torch.no_grad():
            for gi, ri, s in _pending(state):
                if gi not in touched:
                    self.policy_nets[gi].eval()                      # inference uses the running BN statistics
                    touched.add(gi)
                action[gi][ri], output[gi][ri] = self._act(gi, s, eps, debug)
        if self.train:
            for gi in touched:
                self.policy_nets[gi].train()
        return (action, {'output': output}) if debug else action


class DQNIntentionPolicy(DQNPolicy):
    """Adds the intention-prediction nets ``FCN(C-1 -> 1)``: their sigmoid output replaces the ground-truth
    intention channel outside of training rollouts that are allowed to use it."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.intention_nets = self.build_intention_nets()
        if self.policy_checkpoint is not None:
            self._restore(self.intention_nets, 'state_dicts_intention')
            print("=> loaded intention network '{}'".format(self.cfg.policy_path))

    def build_intention_nets(self):
        return [self._new_net(self.cfg.num_input_channels - 1, 1) for _ in self.robot_group_types]

    def _predict(self, gi, s):
        x = torch.from_numpy(np.ascontiguousarray(s, dtype=np.float32)).unsqueeze(0).to(self.device).permute(0, 3, 1, 2)
        return torch.sigmoid(self.intention_nets[gi](x))[0, 0].cpu().numpy()

    def step_intention(self, state, debug=False):
        augmented, predicted = _like(state), _like(state)
        touched = set()
        with torch.no_grad():
            for gi, ri, s in _pending(state):
                if gi not in touched:
                    self.intention_nets[gi].eval()
                    touched.add(gi)
                p = self._predict(gi, s)
                predicted[gi][ri] = p
                augmented[gi][ri] = np.concatenate((s, p[:, :, None]), axis=2)
        if self.train:
            for gi in touched:
                self.intention_nets[gi].train()
        return (augmented, {'output_intention': predicted}) if debug else augmented

    def step(self, state, exploration_eps=None, debug=False, use_ground_truth_intention=False):
        if self.train and use_ground_truth_intention:                # the state already carries the true intention map
            return super().step(state, exploration_eps=exploration_eps, debug=debug)
        if self.train:                                               # drop the ground-truth channel, predict it instead
            state = [[None if s is None else s[:, :, :-1] for s in group] for group in state]
        result = self.step_intention(state, debug=debug)
        state, intention_info = result if debug else (result, None)
        result = super().step(state, exploration_eps=exploration_eps, debug=debug)
        if not debug:
            return result
        action, info = result
        info['state_intention'] = state
        info['output_intention'] = intention_info['output_intention']
        return action, info
