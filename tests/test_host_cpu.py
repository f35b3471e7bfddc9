"""CPU-side checks: the C-ABI library loads and exports every symbol include/simq.h declares, the flat
layout matches the reference's state_dict inventory, the host mirror of networks.FCN keeps the
reference's names / checkpoint format, and the product path refuses to run without a GPU."""
import os
import re

import numpy as np
import pytest
import torch

from oracle import fcn_oracle as O
from spatial_intention_maps_b200 import _lib, networks, policies, synth, train as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, 'include', 'simq.h')).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return sorted(set(re.findall(r'\b(simq_[a-z0-9_]+)\s*\(', txt)))


def test_library_exports_every_declared_symbol():
    names = _declared()
    assert len(names) >= 15
    L = _lib.lib()
    for n in names:
        assert hasattr(L, n), f'{n} declared in include/simq.h but not exported by libsimq.so'
        assert n in _lib.EXPORTS, f'{n} has no ctypes prototype in _lib.py'
    assert L.simq_version() >= 1


@pytest.mark.parametrize('C,A', [(4, 2), (5, 1), (3, 2), (10, 2)])
def test_layout_matches_reference_inventory(C, A):
    n_p, n_b, po, bo = _lib.layout(C, A)
    spec = O.state_spec(C, A)
    tr = [(n, s) for n, s, k in spec if k == 'param' and not n.startswith('resnet18.fc.')]
    assert len(po) == len(tr) + 1 == 71
    assert [po[i + 1] - po[i] for i in range(70)] == [int(np.prod(s)) for _, s in tr]
    bns = [s[0] for n, s, k in spec if n.endswith('running_mean')]
    assert [bo[i + 1] - bo[i] for i in range(22)] == [2 * c for c in bns]
    assert n_p == po[-1] and n_b == bo[-1]


def test_layout_rejects_bad_arguments():
    assert _lib.lib().simq_layout(0, 2, None, None, None, None) != 0
    assert b'unsupported' in _lib.lib().simq_last_error()
    assert _lib.lib().simq_layout(5, 3, None, None, None, None) != 0


def test_fcn_state_dict_is_the_reference_format():
    net = networks.SingleDeviceParallel(networks.FCN(5, 2))
    sd = net.state_dict()
    spec = O.state_spec(5, 2)
    assert list(sd.keys()) == ['module.' + n for n, _, _ in spec] and len(sd) == 138
    for n, shape, kind in spec:
        assert tuple(sd['module.' + n].shape) == tuple(shape)
        assert sd['module.' + n].dtype == (torch.int64 if kind == 'nbt' else torch.float32)
    st = O.make_state(5, 2, 3)
    v0 = net.module.params_version
    net.load_state_dict({'module.' + k: v for k, v in st.items()})
    assert net.module.params_version != v0                     # the library must re-pack its weight shadows
    for k, v in st.items():
        assert torch.equal(net.state_dict()['module.' + k], v)
    # parameters are views of one flat vector in simq_layout order
    flat = net.module.flat_params
    po = net.module._layout[2]
    for i, (name, p) in enumerate(net.module.trainable()):
        assert p.data_ptr() == flat.data_ptr() + 4 * po[i]
        assert torch.equal(flat[po[i]:po[i + 1]].view(p.shape), st[name])
    assert int(net.module.flat_nbt[0]) == 3 and float(net.module.flat_bn[0]) == float(st['resnet18.bn1.running_mean'][0])
    # a stock optimizer sees 72 parameters (70 trainable + the never-executed resnet18.fc.*)
    assert len(list(net.parameters())) == 72


def test_init_follows_reference_distributions():
    torch.manual_seed(0)
    net = networks.FCN(4, 2)
    w = net.resnet18.layer3[0].conv1.weight                     # kaiming_normal(fan_out, relu): std = sqrt(2/(256*9))
    assert abs(float(w.std()) - (2.0 / (256 * 9)) ** 0.5) < 2e-3
    assert float(net.conv1.weight.abs().max()) <= 1 / 512 ** 0.5 + 1e-6      # Conv2d default: U(+-1/sqrt(fan_in))
    assert torch.equal(net.bn1.weight, torch.ones(128)) and torch.equal(net.bn1.running_var, torch.ones(128))


def test_no_cpu_fallback():
    net = networks.FCN(4, 2)
    with pytest.raises(_lib.SimqError, match='no CPU fallback'):
        net(torch.zeros(1, 4, 96, 96))
    if not torch.cuda.is_available():
        with pytest.raises(_lib.SimqError):
            _lib.Ctx(0, 4, 2, 1)                                # simq_ctx_create fails loudly without a device


def test_policy_surface():
    class Cfg:
        robot_config = [{'lifting_robot': 2}, {'pushing_robot': 2}]
        num_input_channels, batch_size, final_exploration = 5, 8, 0.01
        checkpoint_path = policy_path = None
    pol = policies.DQNPolicy(Cfg(), train=True, device='cpu')
    assert pol.num_robot_groups == 2 and [n.module.num_output_channels for n in pol.policy_nets] == [2, 1]
    assert list(pol.policy_nets[0].state_dict())[0] == 'module.resnet18.conv1.weight'
    s = synth.synth_states(1, 5, 0)[0]
    t = pol.apply_transform(s)                                  # ToTensor on float32 HWC: transpose only, no /255
    assert t.shape == (1, 5, 96, 96) and float(t.max()) == float(s.max())
    tgt = pol.build_policy_nets()
    tgt[0].load_state_dict(pol.policy_nets[0].state_dict())    # train.py:213-216
    assert torch.equal(tgt[0].module.flat_params, pol.policy_nets[0].module.flat_params)


def test_train_groups_checks_its_arguments_and_has_no_cpu_path():
    """train.train_groups (the training block of train.py:253-263): one target net / optimizer / batch per robot group,
    intention optimizers when cfg.use_predicted_intention; on a CPU-resident policy the first update raises (no fallback)."""
    class Cfg:
        robot_config = [{'lifting_robot': 2}, {'pushing_robot': 2}]
        num_input_channels, batch_size, final_exploration = 5, 4, 0.01
        checkpoint_path = policy_path = None
        discount_factors = [0.85, 0.85]
        use_predicted_intention = True
        use_double_dqn, grad_norm_clipping = True, 100
    cfg = Cfg()
    pol = policies.DQNIntentionPolicy(cfg, train=True, device='cpu')
    tgts = pol.build_policy_nets()
    opts = [torch.optim.SGD(n.parameters(), lr=0.01, momentum=0.9) for n in pol.policy_nets]
    batches = [synth.synth_batch(4, 5, A, 3 + A, terminal_every=2) for A in (2, 1)]
    with pytest.raises(ValueError):
        T.train_groups(cfg, pol, tgts[:1], opts, batches)
    with pytest.raises(ValueError):
        T.train_groups(cfg, pol, tgts, opts, batches)             # intention optimizers missing
    opts_i = [torch.optim.SGD(n.parameters(), lr=0.01, momentum=0.9) for n in pol.intention_nets]
    with pytest.raises(_lib.SimqError):
        T.train_groups(cfg, pol, tgts, opts, batches, opts_i)


def test_host_batch_compacts_next_states_like_the_reference():
    batch = synth.synth_batch(8, 4, 2, 5, terminal_every=4)
    hb = T.HostBatch(8, 4).fill(batch)
    assert hb.Bn == 6 and hb.nonfinal.tolist() == [1, 1, 1, 0, 1, 1, 1, 0]
    nf = [n for n in batch.next_state if n is not None]        # train.py:112
    assert np.array_equal(hb.ns[:6].numpy(), np.stack(nf)) and np.array_equal(hb.s.numpy(), np.stack(batch.state))
    assert hb.action.tolist() == list(batch.action)
    with pytest.raises(ValueError):
        T.HostBatch(4, 4).fill(batch)
    sh = T.shard_batch(batch, 1, 2)
    assert len(sh.state) == 4 and sh.action == batch.action[4:] and sh.next_state[3] is None


class _RefReplayBuffer:
    """The reference's ReplayBuffer, train.py:28-45, restated for the comparison below."""

    def __init__(self, capacity):
        self.capacity, self.buffer, self.position = capacity, [], 0

    def push(self, *args):
        if len(self.buffer) < self.capacity:
            self.buffer.append(None)
        self.buffer[self.position] = T.Transition(*args)
        self.position = (self.position + 1) % self.capacity

    def sample(self, batch_size):
        import random
        return T.Transition(*zip(*random.sample(self.buffer, batch_size)))

    def __len__(self):
        return len(self.buffer)


def test_replay_buffer_matches_reference_semantics():
    import pickle
    import random
    from spatial_intention_maps_b200.replay import ReplayBuffer
    ours, ref = ReplayBuffer(6, device='cpu'), _RefReplayBuffer(6)
    tr = synth.synth_batch(10, 4, 2, 3, terminal_every=3)
    for i in range(10):                                         # wraps around: capacity 6
        for b in (ours, ref):
            b.push(tr.state[i], tr.action[i], tr.reward[i], tr.next_state[i])
        assert len(ours) == len(ref)
    for seed in (0, 1):
        random.seed(seed); a = ours.sample(4)
        random.seed(seed); b = ref.sample(4)
        at = a.to_transition()
        assert at.action == b.action and np.allclose(at.reward, b.reward)
        for x, y in zip(at.state, b.state):
            assert np.array_equal(x, y)
        for x, y in zip(at.next_state, b.next_state):
            assert (x is None and y is None) or np.array_equal(x, y)
        out_s, out_ns = torch.zeros(4, 96, 96, 4), torch.zeros(4, 96, 96, 4)
        action, reward, nf, Bn = ours.gather(a, out_s, out_ns)
        assert Bn == sum(n is not None for n in b.next_state) and list(action) == list(b.action)
        assert np.array_equal(out_ns[:Bn].numpy(), np.stack([n for n in b.next_state if n is not None]))
    clone = pickle.loads(pickle.dumps(ours))                    # train.py:331 checkpoints the buffers with torch.save
    assert len(clone) == 6 and clone.position == ours.position
    random.seed(5); x = clone.sample(3).to_transition()
    random.seed(5); y = ours.sample(3).to_transition()
    assert x.action == y.action and all(np.array_equal(p, q) for p, q in zip(x.state, y.state))


def test_reference_checkpoint_formats_round_trip(tmp_path):
    """policy_*.pth.tar as written by train.py:313-321 ({'timestep', 'state_dicts': [DataParallel state_dict]}) loads
    through DQNPolicy (policies.py:25-33), and an optimizer state_dict (train.py:324, 200-210) re-binds to the flat
    momentum vector."""
    st = O.make_state(5, 2, 9)
    ckpt = {'timestep': 123, 'state_dicts': [{'module.' + k: v for k, v in st.items()}]}
    path = str(tmp_path / 'policy_00000123.pth.tar')
    torch.save(ckpt, path)

    class Cfg:
        robot_config = [{'lifting_robot': 4}]
        num_input_channels, batch_size, final_exploration = 5, 8, 0.01
        checkpoint_path, policy_path = 'x', path
    pol = policies.DQNPolicy(Cfg(), train=True, device='cpu')
    net = pol.policy_nets[0].module
    assert net.training and torch.equal(net.conv3.weight, st['conv3.weight']) and int(net.bn2.num_batches_tracked) == 3
    # what our modules save is what the reference expects back
    again = {'timestep': 124, 'state_dicts': [pol.policy_nets[0].state_dict()]}
    torch.save(again, path)
    back = torch.load(path, map_location='cpu')['state_dicts'][0]
    assert list(back.keys()) == list(ckpt['state_dicts'][0].keys())
    assert all(torch.equal(back[k], ckpt['state_dicts'][0][k]) for k in back)
    # optimizer resume: buffers replaced by load_state_dict are copied into / re-bound to the flat momentum vector
    opt = torch.optim.SGD(pol.policy_nets[0].parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    T._momentum_views(net, opt)
    sd = opt.state_dict()
    for i, stt in sd['state'].items():
        stt['momentum_buffer'] = torch.full_like(stt['momentum_buffer'], float(i + 1))
    opt.load_state_dict(sd)
    T._momentum_views(net, opt)
    po = net._layout[2]
    assert float(net.flat_momentum[po[0]]) == 1.0 and float(net.flat_momentum[po[3]]) == 4.0 and net.momentum_initialized
    p0 = net._tr_cache[0]
    assert opt.state[p0]['momentum_buffer'].data_ptr() == net.flat_momentum.data_ptr()


def test_profile_tools_read_the_committed_launch_list():
    """tools/launch_summary.py and tools/per_layer_table.py (the scripts behind profiles/*_summary.md and the per-layer
    roofline tables) still parse the committed ncu launch list."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csv = os.path.join(root, 'profiles', 'r1_v13_launches.csv')
    out = subprocess.run([sys.executable, os.path.join(root, 'tools', 'launch_summary.py'), csv], capture_output=True, text=True, check=True).stdout
    assert 'conv2w_umma_kernel' in out and out.startswith('Total ')
    out = subprocess.run([sys.executable, os.path.join(root, 'tools', 'per_layer_table.py'), csv], capture_output=True, text=True, check=True).stdout
    assert 'Backward (B = 128)' in out and 'of the per-layer roofline (time-weighted)' in out


def test_host_batch_stages_states_before_next_states():
    """HostBatch.fill(after_states=...): when the callback runs (the caller starts the upload of s there) the states and the
    per-sample vectors are complete and the next states have not been touched yet; afterwards everything is staged."""
    batch = synth.synth_batch(32, 4, 2, 9, terminal_every=4)
    hb = T.HostBatch(32, 4)
    hb.ns.fill_(float('nan'))
    seen = {}

    def after_states():
        seen['s_ok'] = all(np.array_equal(hb.s.numpy()[i], batch.state[i]) for i in range(32))
        seen['vec_ok'] = hb.Bn == 24 and hb.action.tolist() == list(batch.action) and int(hb.nonfinal.sum()) == 24
        seen['ns_untouched'] = bool(torch.isnan(hb.ns).all())
    hb.fill(batch, after_states=after_states)
    assert seen == {'s_ok': True, 'vec_ok': True, 'ns_untouched': True}
    nxt = [n for n in batch.next_state if n is not None]
    assert all(np.array_equal(hb.ns.numpy()[j], nxt[j]) for j in range(24))


def test_deepcopy_gives_an_independent_network_with_its_own_flat_storage():
    """``copy.deepcopy(policy_net)`` is the common target-network idiom: the copy's parameters and buffers must be views into
    ITS OWN flat vectors (the library reads those), not clones detached from them, and it must not share the context."""
    import copy
    net = networks.FCN(4, 2)
    net.flat_momentum = torch.ones_like(net.flat_params)
    cp = copy.deepcopy(net)
    lo, hi = cp.flat_params.data_ptr(), cp.flat_params.data_ptr() + 4 * cp.flat_params.numel()
    for _, p in cp.trainable():
        assert lo <= p.data_ptr() < hi
    assert lo <= cp.conv3.weight.data_ptr() < hi and cp.flat_params.data_ptr() != net.flat_params.data_ptr()
    b0 = cp.flat_bn.data_ptr()
    assert b0 <= cp.bn1.running_mean.data_ptr() < b0 + 4 * cp.flat_bn.numel()
    assert cp._ctx is None and cp.flat_momentum.data_ptr() != net.flat_momentum.data_ptr()
    assert all(torch.equal(a, b) for a, b in zip(net.state_dict().values(), cp.state_dict().values()))
    with torch.no_grad():
        cp.conv3.weight.add_(1.0)                      # writes through the parameter land in the copy's flat vector only
    off = cp._layout[2][-3]
    assert abs(float((cp.flat_params[off:off + 64] - net.flat_params[off:off + 64]).min()) - 1.0) < 1e-6
    sd = net.state_dict()
    cp.load_state_dict(sd)
    assert torch.equal(cp.flat_params, net.flat_params)


def test_mark_params_changed_bumps_the_version_data_writes_do_not():
    net = networks.FCN(4, 2)
    v0 = net.params_version
    for p in net.parameters():
        p.data.mul_(0.5)                               # invisible to autograd's version counters (Polyak-style update)
    assert net.params_version == v0
    net.mark_params_changed()
    assert net.params_version > v0


BENCH_LINES = ['r2_final_bench.json', 'r2_final_bench_n2.json', 'r2_final_bench_n4.json', 'r2_final_bench_n8.json']


@pytest.mark.parametrize('name', BENCH_LINES)
def test_committed_bench_lines_follow_the_contract(name):
    """The bench lines kept under profiles/ are what DESIGN.md quotes: each must be ONE JSON line with the driver contract's keys,
    internally consistent (value = samples per step / time, roofline.frac = achieved / peak, e2e with its copy sizes), measured with
    clocks the contract accepts, and -- at N > 1 -- carry a passed multi-rank parity check."""
    import json
    path = os.path.join(ROOT, 'profiles', name)
    text = open(path).read().strip()
    assert '\n' not in text, 'one JSON line'
    d = json.loads(text)
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline', 'dtype',
              'data', 'config', 'e2e', 'gpu_launches', 'roofline', 'clocks'):
        assert k in d, k
    assert d['unit'] == 'samples/s' and d['higher_is_better'] is True and d['scaling'] == 'weak' and d['data'] == 'synthetic'
    assert d['vs_baseline'] is None                                  # BASELINE.md has no published number for this metric
    assert d['warmup'] >= 3 and d['steps'] >= 1 and d['gpu_launches'] > 0
    n = d['n_gpus']
    assert d['config']['global_batch'] == 128 * n and 'workload' in d['config'] and 'model' not in d['config']
    assert abs(d['value'] - d['config']['global_batch'] / d['ms_per_step'] * 1e3) <= 1e-6 * d['value']
    r = d['roofline']
    assert r['bound'] in ('hbm', 'tensor') and r['unit'] in ('GB/s', 'TFLOP/s')
    assert abs(r['frac'] - r['achieved'] / r['peak']) < 1e-9 and 0.0 < r['frac'] < 1.0
    assert r['issued_frac'] < 1.4                                    # issued MMAs against the sustained peak: plausibility bound
    e = d['e2e']
    assert e['unit'] == d['unit'] and e['h2d_bytes_per_step'] > 0 and e['d2h_bytes_per_step'] > 0
    assert e['value'] <= d['value'] * 1.02                           # the end-to-end number cannot beat the device-resident one
    bad = {'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'}
    assert not bad & set(d['clocks']['reasons'])
    if n == 1:
        cb = d['cpu_baseline']
        assert cb['kind'] == 'reference' and cb['cores'] >= 1 and cb['value'] > 0 and 'sample' in cb
        assert cb['argmax_check_vs_cpu_oracle']['eval_argmax_equal'] == cb['argmax_check_vs_cpu_oracle']['samples']
    else:
        assert d['config']['parity_check']['ok'] is True


def test_committed_reference_arm_line():
    import json
    d = json.loads(open(os.path.join(ROOT, 'profiles', 'r2_final_bench_reference_arm.json')).read().strip())
    mine = json.loads(open(os.path.join(ROOT, 'profiles', 'r2_final_bench.json')).read().strip())
    assert d['impl'] == 'reference' and d['metric'] == mine['metric'] and d['unit'] == mine['unit']
    assert d['config']['workload'] == mine['config']['workload'] and d['higher_is_better'] is True
    assert d['cpu_baseline']['kind'] == 'reference' and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}


def test_timeline_tool_reproduces_the_committed_analysis():
    """profiles/r2_timeline.md quotes tools/timeline.py's analysis of the committed kernel timeline: re-derive it from the CSV."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('timeline_tool', os.path.join(ROOT, 'tools', 'timeline.py'))
    tl = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tl)
    r = tl.analyse(tl.load_csv(os.path.join(ROOT, 'profiles', 'r2_timeline_after.csv')))
    assert r['kernels'] == 243                                           # the launches of one step
    assert abs(r['span'] / 1e3 - 19.59) < 0.02 and abs(r['tensor_idle'] / 1e3 - 1.30) < 0.02
    assert r['tensor_sum'] > r['tensor_union'] > 0                       # tensor-core kernels of different lanes overlap
    assert r['no_kernel'] < 0.05e3                                       # the graph leaves no launch gaps
    text = open(os.path.join(ROOT, 'profiles', 'r2_timeline.md')).read()
    assert tl.report(r).splitlines()[0] in text
    # a synthetic case: two tensor kernels with an elementwise kernel between them
    s = tl.analyse([(0.0, 10.0, 'conv_umma_kernel<64, 65, 3>'), (10.0, 14.0, 'bn_apply_kernel<0>'), (15.0, 25.0, 'wgrad_umma_kernel<64, 2>')])
    assert s['span'] == 25.0 and s['tensor_union'] == 20.0 and s['tensor_idle'] == 5.0 and abs(s['no_kernel'] - 1.0) < 1e-9
    assert dict(s['fill']) == {'bn_apply_kernel<0>': 4.0}
