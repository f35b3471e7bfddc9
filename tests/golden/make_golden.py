"""Generate golden fixtures by running the UNMODIFIED reference (imported from /root/reference)
on CPU fp32.  Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Imports reference ``networks``, ``policies``, ``train`` with ``envs`` / ``utils`` stubbed
(pybullet, skimage, munch ... are not installed) and drives
``networks.FCN.forward``, ``train.train`` and ``policies.DQNPolicy.step`` unmodified.  Parameters
come from ``oracle.fcn_oracle.make_state`` (numpy RandomState => reproducible anywhere) and are
loaded through the reference's own ``load_state_dict``; inputs from
``spatial_intention_maps_b200.synth``.  Outputs: ``tests/golden/*.npz``.
"""
import os
import sys
import types

os.environ.setdefault('MKL_NUM_THREADS', '1')   # as train.py:10
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')

import numpy as np
import torch

envs = types.ModuleType('envs')


class VectorEnv:  # static surface of envs.py:366-376 (+ :810, :1090 for the channel counts)
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
sys.modules['envs'] = envs
sys.modules['utils'] = types.ModuleType('utils')

import networks  # noqa: E402  (reference)
import policies  # noqa: E402  (reference)
import train as ref_train  # noqa: E402  (reference)

from oracle import fcn_oracle as O  # noqa: E402
from spatial_intention_maps_b200 import synth  # noqa: E402

torch.manual_seed(0)
torch.set_num_threads(os.cpu_count())


def load_ref(net, st):
    net.load_state_dict({'module.' + k: v.clone() for k, v in st.items()})


def ref_state(net):
    return {k[len('module.'):]: v.detach().clone() for k, v in net.state_dict().items()}


def make_cfg(C, robot_type, B, clip=100):
    return types.SimpleNamespace(
        robot_config=[{robot_type: 1}], num_input_channels=C, checkpoint_path=None, policy_path=None,
        final_exploration=0.01, batch_size=B, use_double_dqn=True, grad_norm_clipping=clip)


def gen_manifest():
    net = torch.nn.DataParallel(networks.FCN(5, 2))
    names, shapes, dtypes = [], [], []
    for k, v in net.state_dict().items():
        names.append(k); shapes.append(str(tuple(v.shape))); dtypes.append(str(v.dtype))
    spec = O.state_spec(5, 2)
    assert [('module.' + n) for n, _, _ in spec] == names, 'oracle state_spec order != reference'
    assert [str(tuple(s)) for _, s, _ in spec] == shapes
    np.savez(os.path.join(HERE, 'manifest_C5_A2.npz'), names=np.array(names), shapes=np.array(shapes),
             dtypes=np.array(dtypes))


def gen_forward():
    out = {}
    for (C, A) in [(4, 2), (5, 2), (5, 1), (8, 2), (3, 2), (10, 2)]:
        seed = 100 + C * 10 + A
        st = O.make_state(C, A, seed)
        net = torch.nn.DataParallel(networks.FCN(C, A))
        load_ref(net, st)
        x = synth.synth_states(2, C, seed)
        xt = torch.cat([policies.transforms.ToTensor()(s).unsqueeze(0) for s in x])
        net.eval()
        with torch.no_grad():
            q_eval = net(xt)
        net.train()
        with torch.no_grad():
            q_train = net(xt)
        after = ref_state(net)
        key = f'C{C}_A{A}'
        out[key + '_seed'] = np.int64(seed)
        out[key + '_q_eval'] = q_eval.numpy()
        out[key + '_q_train'] = q_train.numpy()
        out[key + '_bn1_rm'] = after['bn1.running_mean'].numpy()
        out[key + '_bn1_rv'] = after['bn1.running_var'].numpy()
        out[key + '_l4_rm'] = after['resnet18.layer4.1.bn2.running_mean'].numpy()
        out[key + '_l4_rv'] = after['resnet18.layer4.1.bn2.running_var'].numpy()
        out[key + '_nbt'] = after['bn2.num_batches_tracked'].numpy()
        print('forward', key, float(q_eval.abs().max()), float(q_train.abs().max()))
    np.savez(os.path.join(HERE, 'forward.npz'), **out)


def run_ref_steps(C, robot_type, A, B, gamma, nsteps, seed, terminal_every, clip=100):
    cfg = make_cfg(C, robot_type, B, clip)
    policy = policies.DQNPolicy(cfg, train=True)
    st = O.make_state(C, A, seed)
    load_ref(policy.policy_nets[0], st)
    target = policy.build_policy_nets()[0]
    target.load_state_dict(policy.policy_nets[0].state_dict())     # train.py:213-216
    target.eval()
    policy.policy_nets[0].train()
    opt = torch.optim.SGD(policy.policy_nets[0].parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)  # train.py:186
    infos, grads = [], None
    for step in range(nsteps):
        batch = synth.synth_batch(B, C, A, seed + 1000 * step, terminal_every=terminal_every)
        batch = ref_train.Transition(*batch)
        info = ref_train.train(cfg, policy.policy_nets[0], target, opt, batch, policy.apply_transform, gamma)
        infos.append((info['loss'], info['td_error']))
        if step == 0:
            grads = {n[len('module.'):]: p.grad.detach().clone()
                     for n, p in policy.policy_nets[0].named_parameters() if p.grad is not None}
    after = ref_state(policy.policy_nets[0])
    mom = {}
    for (n, p) in policy.policy_nets[0].named_parameters():
        if p in opt.state and 'momentum_buffer' in opt.state[p]:
            mom[n[len('module.'):]] = opt.state[p]['momentum_buffer'].detach().clone()
    return infos, grads, after, mom


def pack_step(prefix, out, infos, grads, after, mom, C, A):
    names = O.trainable_names(C, A)
    assert sorted(grads.keys()) == sorted(names), 'trainable set mismatch'
    out[prefix + '_loss'] = np.array([i[0] for i in infos], dtype=np.float64)
    out[prefix + '_td'] = np.array([i[1] for i in infos], dtype=np.float64)
    out[prefix + '_grad_digest'] = np.stack([O.digest(grads[n]) for n in names])
    out[prefix + '_param_digest'] = np.stack([O.digest(after[n]) for n in names])
    out[prefix + '_mom_digest'] = np.stack([O.digest(mom[n]) for n in names])
    out[prefix + '_grad_norm'] = np.float64(np.sqrt(sum(float((g.double() ** 2).sum()) for g in grads.values())))
    bn = [n for n, _, k in O.state_spec(C, A) if k == 'buffer']
    out[prefix + '_bn'] = np.concatenate([after[n].numpy().ravel() for n in bn])
    out[prefix + '_nbt'] = np.array([int(after[n]) for n, _, k in O.state_spec(C, A) if k == 'nbt'])
    out[prefix + '_fc_w_digest'] = O.digest(after['resnet18.fc.weight'])


def gen_steps():
    out = {}
    # (key, C, robot_type, A, B, gamma, nsteps, seed, terminal_every) -- c1/c2/c3 of SURVEY.md §8
    cases = [
        ('c1', 4, 'lifting_robot', 2, 16, 0.75, 1, 11, 8),
        ('c2', 5, 'lifting_robot', 2, 32, 0.85, 1, 12, 16),
        ('c3', 5, 'pushing_robot', 1, 128, 0.85, 1, 13, 64),
        ('traj', 4, 'lifting_robot', 2, 8, 0.75, 3, 14, 4),   # 3 steps: beyond that 1-ulp differences
        # are amplified chaotically by Double-DQN argmax flips at B=8 (observed: 1e-5 at step 3 -> 30 % at step 5)
    ]
    for key, C, rt, A, B, gamma, nsteps, seed, te in cases:
        infos, grads, after, mom = run_ref_steps(C, rt, A, B, gamma, nsteps, seed, te)
        pack_step(key, out, infos, grads, after, mom, C, A)
        out[key + '_cfg'] = np.array([C, A, B, nsteps, seed, te], dtype=np.int64)
        out[key + '_gamma'] = np.float64(gamma)
        print('step', key, infos)
    np.savez(os.path.join(HERE, 'steps.npz'), **out)


def gen_steps_cstar():
    """The north_star's headline shape c* (C=8 input channels, A=2, gamma 0.85): one reference update at
    B=16, kept in its own file so that adding it does not regenerate steps.npz."""
    out = {}
    key, C, rt, A, B, gamma, nsteps, seed, te = 'cstar', 8, 'lifting_robot', 2, 16, 0.85, 1, 15, 8
    infos, grads, after, mom = run_ref_steps(C, rt, A, B, gamma, nsteps, seed, te)
    pack_step(key, out, infos, grads, after, mom, C, A)
    out[key + '_cfg'] = np.array([C, A, B, nsteps, seed, te], dtype=np.int64)
    out[key + '_gamma'] = np.float64(gamma)
    print('step', key, infos)
    np.savez(os.path.join(HERE, 'steps_cstar.npz'), **out)


def gen_steps_clip():
    """train.py:133-134 with the clip ENGAGED (max-norm 1 and 5: the gradient norm of these batches is ~10-40, so
    coef < 1) and with ``grad_norm_clipping: None`` (the branch that skips clip_grad_norm_); ``_grad_digest`` holds the
    gradients AFTER the in-place rescale, ``_grad_norm`` their norm (= the max-norm when the clip is active)."""
    out = {}
    for key, clip in (('clip1', 1.0), ('clip5', 5.0), ('clipnone', None)):
        # (seed 11 = the c1 batch: its gradient is known to be well conditioned at B=16 -- with 16 one-hot dL/dQ entries a single ReLU
        # flip near an output pixel can move the gradient norm by 0.5 %, which would mask what this case is about: the clip)
        C, rt, A, B, gamma, nsteps, seed, te = 4, 'lifting_robot', 2, 16, 0.75, 2, 11, 8
        infos, grads, after, mom = run_ref_steps(C, rt, A, B, gamma, nsteps, seed, te, clip)
        pack_step(key, out, infos, grads, after, mom, C, A)
        out[key + '_cfg'] = np.array([C, A, B, nsteps, seed, te], dtype=np.int64)
        out[key + '_gamma'] = np.float64(gamma)
        out[key + '_clip'] = np.float64(-1.0 if clip is None else clip)
        print('step', key, infos, float(out[key + '_grad_norm']))
    np.savez(os.path.join(HERE, 'steps_clip.npz'), **out)


def gen_steps_cstar128():
    """c* at the north_star's full batch: C=8, A=2, gamma 0.85, B=128, every 64th transition terminal (digests only)."""
    out = {}
    key, C, rt, A, B, gamma, nsteps, seed, te = 'cstar128', 8, 'lifting_robot', 2, 128, 0.85, 1, 17, 64
    infos, grads, after, mom = run_ref_steps(C, rt, A, B, gamma, nsteps, seed, te)
    pack_step(key, out, infos, grads, after, mom, C, A)
    out[key + '_cfg'] = np.array([C, A, B, nsteps, seed, te], dtype=np.int64)
    out[key + '_gamma'] = np.float64(gamma)
    print('step', key, infos)
    np.savez(os.path.join(HERE, 'steps_cstar128.npz'), **out)


def gen_policy():
    C, A, seed = 4, 2, 21
    cfg = make_cfg(C, 'lifting_robot', 16)
    policy = policies.DQNPolicy(cfg, train=False)
    load_ref(policy.policy_nets[0], O.make_state(C, A, seed))
    states = synth.synth_states(16, C, seed)
    acts, qmax = [], []
    for i in range(16):
        a, info = policy.step([[states[i]]], exploration_eps=0.0, debug=True)
        acts.append(a[0][0]); qmax.append(float(info['output'][0][0].max()))
    np.savez(os.path.join(HERE, 'policy_step.npz'), cfg=np.array([C, A, seed], dtype=np.int64),
             actions=np.array(acts, dtype=np.int64), qmax=np.array(qmax))
    print('policy', acts)


def gen_intention():
    """train.train_intention (train.py:143-158) on the reference's DQNIntentionPolicy nets: 2 steps, B=8, C=5
    (4 input channels + the ground-truth intention channel)."""
    C, B, seed = 5, 8, 31
    cfg = make_cfg(C, 'lifting_robot', B)
    policy = policies.DQNIntentionPolicy(cfg, train=True)
    net = policy.intention_nets[0]
    load_ref(net, O.make_state(C - 1, 1, seed))
    net.train()
    opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)      # train.py:190
    losses, grads = [], None
    for step in range(2):
        batch = ref_train.Transition(*synth.synth_batch(B, C, 2, seed + 1000 * step, terminal_every=None))
        info = ref_train.train_intention(net, opt, batch, policy.apply_transform)
        losses.append(info['loss_intention'])
        if step == 0:
            grads = {n[len('module.'):]: p.grad.detach().clone() for n, p in net.named_parameters() if p.grad is not None}
    after = ref_state(net)
    names = O.trainable_names(C - 1, 1)
    np.savez(os.path.join(HERE, 'intention.npz'), cfg=np.array([C, B, seed], dtype=np.int64), loss=np.array(losses),
             grad_digest=np.stack([O.digest(grads[n]) for n in names]),
             param_digest=np.stack([O.digest(after[n]) for n in names]),
             nbt=np.array([int(after[n]) for n, _, k in O.state_spec(C - 1, 1) if k == 'nbt']))
    print('intention', losses)


if __name__ == '__main__':
    parts = sys.argv[1:] or ['manifest', 'forward', 'policy', 'steps', 'steps_cstar', 'intention', 'steps_clip', 'steps_cstar128']
    for part in parts:
        {'manifest': gen_manifest, 'forward': gen_forward, 'policy': gen_policy, 'steps': gen_steps,
         'steps_cstar': gen_steps_cstar, 'intention': gen_intention, 'steps_clip': gen_steps_clip,
         'steps_cstar128': gen_steps_cstar128}[part]()
