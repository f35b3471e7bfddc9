"""Drop-in for the reference's ``policies.DQNPolicy`` / ``DQNIntentionPolicy`` (policies.py:11-146):
same constructor, ``build_policy_nets`` / ``apply_transform`` / ``step`` surface and attributes
(``policy_nets``, ``intention_nets``, ``device``, ``num_robot_groups``), with the networks computed by
the simq CUDA library (``networks.FCN``).  One process drives one GPU; the ``DataParallel`` wrapper of
policies.py:39 is replaced by a non-replicating wrapper that keeps the ``module.`` checkpoint prefix.
"""
from __future__ import annotations

import random

import numpy as np
import torch

from . import networks


class _StaticEnv:
    """The three static methods of envs.VectorEnv the policy needs (envs.py:366-376; action-channel
    counts from envs.py:810 (pushing: 1) and :1090 (lifting / throwing / rescue: 2))."""

    @staticmethod
    def get_num_output_channels(robot_type):
        return 1 if robot_type == 'pushing_robot' else 2

    @staticmethod
    def get_action_space(robot_type):
        return _StaticEnv.get_num_output_channels(robot_type) * 96 * 96

    @staticmethod
    def get_state_width():
        return 96


try:                                    # inside the reference tree the real class is importable
    from envs import VectorEnv          # type: ignore
except Exception:                       # pybullet & co. absent: fall back to the static table
    VectorEnv = _StaticEnv


class DQNPolicy:
    def __init__(self, cfg, train=False, random_seed=None, device=None, max_batch=None):
        self.cfg = cfg
        self.robot_group_types = [next(iter(g.keys())) for g in self.cfg.robot_config]
        self.train = train
        if random_seed is not None:
            random.seed(random_seed)
        self.num_robot_groups = len(self.robot_group_types)
        if device is None:
            device = torch.device('cuda', torch.cuda.current_device()) if torch.cuda.is_available() else torch.device('cpu')
        self.device = torch.device(device)
        self.max_batch = max_batch if max_batch is not None else int(getattr(cfg, 'batch_size', networks.DEFAULT_MAX_BATCH))
        self.policy_nets = self.build_policy_nets()

        if getattr(self.cfg, 'checkpoint_path', None) is not None:          # policies.py:25-33
            self.policy_checkpoint = torch.load(self.cfg.policy_path, map_location=self.device)
            for i in range(self.num_robot_groups):
                self.policy_nets[i].load_state_dict(self.policy_checkpoint['state_dicts'][i])
                if self.train:
                    self.policy_nets[i].train()
                else:
                    self.policy_nets[i].eval()
            print("=> loaded policy '{}'".format(self.cfg.policy_path))

    def build_policy_nets(self):
        policy_nets = []
        for robot_type in self.robot_group_types:
            num_output_channels = VectorEnv.get_num_output_channels(robot_type)
            policy_nets.append(networks.SingleDeviceParallel(
                networks.FCN(num_input_channels=self.cfg.num_input_channels, num_output_channels=num_output_channels,
                             max_batch=self.max_batch)
            ).to(self.device))
        return policy_nets

    def apply_transform(self, s):
        """torchvision ``ToTensor`` on a float32 HWC ndarray (policies.py:20,44-45): HWC -> CHW, no
        rescaling, plus a leading batch dimension."""
        if s.ndim == 2:
            s = s[:, :, None]
        return torch.from_numpy(np.ascontiguousarray(s.transpose(2, 0, 1))).unsqueeze(0)

    def _to_device_nhwc(self, s):
        """(96,96,C) float32 ndarray -> (1,C,96,96) channels_last device tensor (no transpose pass)."""
        t = torch.from_numpy(np.ascontiguousarray(s, dtype=np.float32)).unsqueeze(0)     # (1,96,96,C)
        return t.to(self.device, non_blocking=True).permute(0, 3, 1, 2)

    def step(self, state, exploration_eps=None, debug=False):
        if exploration_eps is None:
            exploration_eps = self.cfg.final_exploration

        action = [[None for _ in g] for g in state]
        output = [[None for _ in g] for g in state]
        with torch.no_grad():
            for i, g in enumerate(state):
                robot_type = self.robot_group_types[i]
                net = self.policy_nets[i]
                net.eval()
                for j, s in enumerate(g):
                    if s is not None:
                        explore = random.random() < exploration_eps        # same RNG consumption order as policies.py:61-62
                        if explore:
                            a = random.randrange(VectorEnv.get_action_space(robot_type))
                            q = net.module.greedy_action_hwc(s, want_q=True)[1] if debug else None
                        else:
                            a, q = net.module.greedy_action_hwc(s, want_q=debug)
                        action[i][j] = a
                        if debug:                       # the reference copies the Q-map to the host every step
                            output[i][j] = q            # (policies.py:66); only done on request here
                if self.train:
                    net.train()

        if debug:
            info = {'output': output}
            return action, info

        return action


class DQNIntentionPolicy(DQNPolicy):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.intention_nets = self.build_intention_nets()
        if getattr(self.cfg, 'checkpoint_path', None) is not None:
            for i in range(self.num_robot_groups):
                self.intention_nets[i].load_state_dict(self.policy_checkpoint['state_dicts_intention'][i])
                if self.train:
                    self.intention_nets[i].train()
                else:
                    self.intention_nets[i].eval()
            print("=> loaded intention network '{}'".format(self.cfg.policy_path))

    def build_intention_nets(self):
        intention_nets = []
        for _ in range(self.num_robot_groups):
            intention_nets.append(networks.SingleDeviceParallel(
                networks.FCN(num_input_channels=(self.cfg.num_input_channels - 1), num_output_channels=1,
                             max_batch=self.max_batch)
            ).to(self.device))
        return intention_nets

    def step_intention(self, state, debug=False):
        state_intention = [[None for _ in g] for g in state]
        output_intention = [[None for _ in g] for g in state]
        with torch.no_grad():
            for i, g in enumerate(state):
                self.intention_nets[i].eval()
                for j, s in enumerate(g):
                    if s is not None:
                        s_copy = s.copy()
                        x = self._to_device_nhwc(s)
                        o = torch.sigmoid(self.intention_nets[i](x)).squeeze(0).squeeze(0).cpu().numpy()
                        state_intention[i][j] = np.concatenate((s_copy, np.expand_dims(o, 2)), axis=2)
                        output_intention[i][j] = o
                if self.train:
                    self.intention_nets[i].train()

        if debug:
            info = {'output_intention': output_intention}
            return state_intention, info

        return state_intention

    def step(self, state, exploration_eps=None, debug=False, use_ground_truth_intention=False):
        if self.train and use_ground_truth_intention:
            return super().step(state, exploration_eps=exploration_eps, debug=debug)

        if self.train:                              # remove the ground-truth intention map
            state_copy = [[None for _ in g] for g in state]
            for i, g in enumerate(state):
                for j, s in enumerate(g):
                    if s is not None:
                        state_copy[i][j] = s[:, :, :-1]
            state = state_copy

        state = self.step_intention(state, debug=debug)
        if debug:
            state, info_intention = state

        action = super().step(state, exploration_eps=exploration_eps, debug=debug)

        if debug:
            action, info = action
            info['state_intention'] = state
            info['output_intention'] = info_intention['output_intention']
            return action, info

        return action
