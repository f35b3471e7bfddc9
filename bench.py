"""Benchmark of the DQN Q-map training step (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl simq|reference] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one full ``train.train`` update (train.py:108-141) on one synthetic replay minibatch:
online forward on s, train-mode online forward on s' (Double-DQN arg-max), eval-mode target forward,
TD target + SmoothL1, backward, global grad-norm clip, momentum-SGD.  Workload at every N: config c3 of
SURVEY.md §8 (BASELINE.json configs[2]: pushing_4-large_empty, C=5 input channels incl. the intention
map, A=1, gamma 0.85, batch 128 PER GPU -> weak scaling), every 64th transition terminal.  Every timed loop
rotates FOUR distinct minibatches (the net keeps learning: no overfitted, low-activity operands).

``value``  : samples/s with the batches already resident in HBM (device part of the step only).
``e2e``    : samples/s through the public API call a user makes (``train.train`` semantics): pinned host
             batch -> H2D copies -> step -> D2H read of (loss, td_error), all inside the timed region.
``roofline``: the dominant kernel class (tcgen05 conv/dgrad), algorithmic FLOPs / CUDA-event time measured on
             the launching stream during the timed steps, against MEASURED_PEAKS.json's sustained bf16 peak;
             also the literal BASELINE.json metric (forward + backward only) and forward only.
``cpu_baseline``: the reference's own ``train.train`` (oracle/_ref, staged by oracle/build_ref.py) on the host cores,
             bounded sample, plus -- single GPU only -- the SAME unmodified reference step on this GPU through
             PyTorch/cuDNN (fp32 and TF32): the library path this framework replaces.
``config.parity_check`` (N > 1, outside every timed region): rank r's gradients / BN statistics equal a solo run on its
             shard bit for bit, the all-reduced gradient equals the mean of the per-rank gradients, parameters after the
             update are bit-identical on all ranks (DataParallel semantics of policies.py:39-41).
``--impl reference``: the reference path on the host cores, same workload (batch 128) and a batch-16 sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'DQN Q-map train-step samples/sec (3 fwd + bwd + clip + SGD, batch 128/GPU, 96x96)'
UNIT = 'samples/s'
C_IN, A_OUT, GAMMA, TERMINAL_EVERY = 5, 1, 0.85, 64
FWD_GFLOP = 13.021 - 0.0006            # per sample, C=5, A=1
STEP_GFLOP = 5 * FWD_GFLOP - 0.0723    # 3 forwards + backward (2x forward, no stem dgrad)
CPU_SAMPLE_B = 16
N_BATCHES = 4                          # distinct minibatches rotated through every timed loop


def workload_name(B):
    return f'c3 pushing_4-large_empty: C={C_IN} A={A_OUT} gamma={GAMMA} batch={B}/GPU double-DQN, every {TERMINAL_EVERY}th terminal'


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get('bf16_tflops_sustained', 1408.1), d.get('bf16_tflops', 1664.5), d.get('hbm_gbs', 6446.3), 'measured'
    return 1400.0, 1590.0, 6650.0, 'fallback'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons, pw = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            except Exception:
                continue
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': sorted(reasons),
                'samples': len(sm), 'power_w_max': max(pw) if pw else None}


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference's own train.train (oracle/_ref) on the host cores; oracle port when oracle/_ref is absent
# ---------------------------------------------------------------------------------------------
def cpu_port_steps(steps, warmup, B, terminal_every):
    import torch
    from oracle import fcn_oracle as O
    from spatial_intention_maps_b200 import synth
    from tests.gpu_checks import batch_tensors
    pol = O.make_state(C_IN, A_OUT, 0, perturb=False)
    tgt = O.clone_state(pol)
    mom = None
    batches = [synth.synth_batch(B, C_IN, A_OUT, 1234 + i, terminal_every=terminal_every) for i in range(N_BATCHES)]
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        r = O.dqn_step(pol, tgt, mom, *batch_tensors(batches[i % N_BATCHES]), discount=GAMMA)
        mom = r['momentum']
        if i >= warmup:
            ts.append(time.perf_counter() - t0)
    total = sum(ts)
    return {'value': B * len(ts) / total, 'unit': UNIT, 'ms_per_step': total / len(ts) * 1e3, 'device': 'cpu', 'batch': B, 'steps': steps,
            'loss': r['loss']}


def cpu_reference(steps128, steps16, warmup):
    """Runs in a process that cannot see the GPUs (the reference picks cuda whenever it can, train.py:24)."""
    import torch
    from oracle import ref_runner as R
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if R.available():
        kind = 'reference'
        main = R.time_train(C_IN, A_OUT, 128, GAMMA, steps128, warmup, TERMINAL_EVERY, n_batches=min(N_BATCHES, steps128))
        small = R.time_train(C_IN, A_OUT, CPU_SAMPLE_B, GAMMA, steps16, 1, 8, n_batches=N_BATCHES) if steps16 else None
    else:
        kind = 'port'
        main = cpu_port_steps(steps128, warmup, 128, TERMINAL_EVERY)
        small = cpu_port_steps(steps16, 1, CPU_SAMPLE_B, 8) if steps16 else None
    assert main['device'] == 'cpu', main
    what = "the reference's own train.train (oracle/_ref: train.py:108-141 unmodified, torch CPU fp32)" if kind == 'reference' else \
        'oracle port of train.train (oracle/_ref not staged), torch CPU fp32'
    sample = (f'{steps128} timed + {warmup} warm-up steps of the SAME workload (batch 128, C={C_IN} A={A_OUT}, double-DQN, every '
              f'{TERMINAL_EVERY}th transition terminal) through {what}, {cores} threads')
    return {'value': main['value'], 'unit': UNIT, 'cores': cores, 'kind': kind, 'sample': sample, 'ms_per_step': main['ms_per_step'],
            'steps': steps128, 'warmup': warmup,
            'batch16': None if small is None else {'value': small['value'], 'unit': UNIT, 'ms_per_step': small['ms_per_step'],
                                                   'note': f'{steps16} steps of batch {CPU_SAMPLE_B} (the size at which the CPU path is fastest per sample), every 8th terminal'}}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    os.environ['CUDA_VISIBLE_DEVICES'] = ''          # BEFORE torch is imported: this arm is the reference on the HOST cores
    # bounded: the step takes ~4.5 s at batch 128 on 16 cores
    steps = max(1, min(args.steps, 20))
    warmup = max(1, min(args.warmup, 2))
    cb = cpu_reference(steps, 0 if args.cpu_leg else 12, warmup)
    line = {'impl': 'reference', 'metric': METRIC, 'value': cb['value'], 'unit': UNIT, 'n_gpus': args.gpus, 'steps': steps, 'warmup': warmup,
            'ms_per_step': cb['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic', 'config': {'workload': workload_name(128), 'global_batch': 128, 'parallelism': 'host cores'},
            'cpu_baseline': cb,
            'e2e': {'value': cb['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}
    print(json.dumps(line), flush=True)


def cpu_leg_subprocess(steps=3, warmup=1):
    """cpu_baseline of the GPU arm: the reference arm in a child process that cannot see the GPUs, bounded to ~20 s."""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES='')
    for k in ('RANK', 'LOCAL_RANK', 'WORLD_SIZE', 'MASTER_ADDR', 'MASTER_PORT'):
        env.pop(k, None)
    out = subprocess.run([sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--steps', str(steps), '--warmup', str(warmup),
                          '--cpu-leg'], env=env, capture_output=True, text=True, timeout=600)
    for ln in reversed(out.stdout.strip().splitlines()):
        if ln.startswith('{'):
            return json.loads(ln)['cpu_baseline']
    raise RuntimeError('cpu leg printed no JSON line: ' + out.stderr[-400:])


# ---------------------------------------------------------------------------------------------
# GPU arm helpers
# ---------------------------------------------------------------------------------------------
def make_nets(networks, torch, dev, C, A, B, seed=0):
    torch.manual_seed(seed)
    pol = networks.FCN(C, A, max_batch=B).to(dev).train()
    tgt = networks.FCN(C, A, max_batch=B)
    tgt.load_state_dict(pol.state_dict())
    tgt = tgt.to(dev).eval()
    opt = torch.optim.SGD(pol.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    return pol, tgt, opt


def raw_train_step(T, _lib, C, pol, tgt, db, B, apply_update, out2):
    """simq_train_step called directly (no collective): the solo / local leg of the multi-rank parity check."""
    if pol.flat_momentum is None:
        import torch
        pol.flat_momentum = torch.zeros_like(pol.flat_params)
    L = _lib.lib()
    _lib.check(L.simq_train_step(
        pol.ctx(B).handle, _lib.ptr(pol.flat_params), _lib.ptr(pol.flat_bn), _lib.ptr(pol.flat_nbt), _lib.ptr(tgt.flat_params),
        _lib.ptr(tgt.flat_bn), tgt.params_version, _lib.ptr(pol.flat_grad()), _lib.ptr(pol.flat_momentum), _lib.ptr(db.s), _lib.ptr(db.ns),
        _lib.X_NHWC, _lib.ptr(db.action), _lib.ptr(db.reward), _lib.ptr(db.nonfinal), B, db.Bn, GAMMA, 0.01, 0.9, 1e-4, 100.0, 1, 1,
        1 if apply_update else 0, _lib.ptr(out2), _lib.stream_ptr()), 'simq_train_step')
    pol.mark_params_changed()


def dp_parity_check(dev, world, rank, C, A, B, seed):
    """SURVEY.md §8e / policies.py:39-41, on the hardware, outside any timed region.  Every rank: (1) SOLO step on its shard
    (gradients only) -> local gradients + BN statistics; (2) the data-parallel step of train.train_step_device on an identical
    fresh replica.  Checks: the DP replica's BN running statistics / counters equal the solo run's bit for bit (per-replica
    batch statistics); the all-reduced gradient equals the mean of the gathered solo gradients (<= 1e-6 of its max) and is
    bit-identical on all ranks; parameters and momentum after the update are bit-identical on all ranks."""
    import torch
    import torch.distributed as dist
    from spatial_intention_maps_b200 import _lib, networks, synth, train as T
    batch = synth.synth_batch(B, C, A, seed + rank, terminal_every=8)
    hb = T.HostBatch(B, C).fill(batch)
    db = T.DeviceBatch(B, C, dev).upload(hb)
    torch.cuda.synchronize()
    solo, solo_t, _ = make_nets(networks, torch, dev, C, A, B, seed=seed)        # same seed on every rank: identical replicas
    out2 = torch.zeros(2, device=dev)
    raw_train_step(T, _lib, None, solo, solo_t, db, B, False, out2)
    g_local = solo.flat_grad().clone()
    pol, tgt, opt = make_nets(networks, torch, dev, C, A, B, seed=seed)
    T.train_step_device(pol, tgt, opt, db, B, GAMMA, 100, True)
    torch.cuda.synchronize()
    g_red = pol.flat_grad()                                                      # all-reduced (clip coefficient 1 at these norms)
    gathered = [torch.empty_like(g_local) for _ in range(world)]
    dist.all_gather(gathered, g_local)
    mean = torch.stack(gathered).double().mean(0)
    gn = float(mean.norm())
    coef = min(1.0, 100.0 / (gn + 1e-6))
    err_mean = float((g_red.double() - mean * coef).abs().max() / mean.abs().max().clamp_min(1e-30))

    def same_on_all_ranks(t):
        v = t.contiguous().view(torch.int32).to(torch.int64)
        chk = torch.stack([v.sum(), (v * (torch.arange(v.numel(), device=dev) % 8191 + 1)).sum()])
        allc = [torch.empty_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        return all(bool(torch.equal(allc[0], c)) for c in allc)

    res = {'workload': f'C={C} A={A} batch={B}/rank x {world} ranks',
           'bn_stats_equal_solo_run_bitwise': bool(torch.equal(pol.flat_bn, solo.flat_bn) and torch.equal(pol.flat_nbt, solo.flat_nbt)),
           'allreduced_grad_vs_mean_of_rank_grads_maxrel': err_mean,
           'allreduced_grad_identical_on_all_ranks': same_on_all_ranks(g_red),
           'params_identical_on_all_ranks': same_on_all_ranks(pol.flat_params),
           'momentum_identical_on_all_ranks': same_on_all_ranks(pol.flat_momentum),
           'loss_report_is_mean_of_rank_losses': None, 'grad_norm': gn}
    losses = [torch.empty_like(out2) for _ in range(world)]
    dist.all_gather(losses, out2)
    res['loss_report_is_mean_of_rank_losses'] = bool(
        float((torch.stack(losses).double().mean(0) - db.out2.double()).abs().max()) <= 1e-6 * max(1.0, float(out2.abs().max())))
    res['ok'] = bool(res['bn_stats_equal_solo_run_bitwise'] and err_mean <= 1e-6 and res['allreduced_grad_identical_on_all_ranks']
                     and res['params_identical_on_all_ranks'] and res['momentum_identical_on_all_ranks']
                     and res['loss_report_is_mean_of_rank_losses'])
    del solo, solo_t, pol, tgt, opt, db, gathered
    torch.cuda.empty_cache()
    return res


def c4_groups(dev, world, rank, steps):
    """Config c4 of BASELINE.json (lifting_2_pushing_2: two heterogeneous Q-networks, A=2 and A=1, batch 256 EACH split over the
    ranks, gradient all-reduce per network) through train.train_groups -- the training block of train.py:253-263."""
    import types
    import torch
    from spatial_intention_maps_b200 import policies, synth, train as T
    Bg = 256
    B = Bg // world
    cfg = types.SimpleNamespace(batch_size=B, num_input_channels=C_IN, use_double_dqn=True, grad_norm_clipping=100,
                                robot_config=[{'lifting_robot': 2}, {'pushing_robot': 2}], discount_factors=[0.85, 0.85],
                                final_exploration=0.01, checkpoint_path=None, policy_path=None, use_predicted_intention=False)
    torch.manual_seed(0)
    pol = policies.DQNPolicy(cfg, train=True, device=dev, max_batch=B)
    tgts = pol.build_policy_nets()
    for t, n in zip(tgts, pol.policy_nets):
        n.train()
        t.load_state_dict(n.state_dict()); t.eval()
    opts = [torch.optim.SGD(n.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4) for n in pol.policy_nets]
    batches = [[T.shard_batch(synth.synth_batch(Bg, C_IN, A, 500 + 10 * k + i, terminal_every=TERMINAL_EVERY), rank, world)
                for i, A in enumerate((2, 1))] for k in range(2)]
    for k in range(3):
        info = T.train_groups(cfg, pol, tgts, opts, batches[k % 2])
    torch.cuda.synchronize()
    import torch.distributed as dist
    dist.barrier()
    t0 = time.perf_counter()
    for k in range(steps):
        info = T.train_groups(cfg, pol, tgts, opts, batches[k % 2])
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    ms = float(dt) * 1e3 / steps
    checks = [dp_parity_check(dev, world, rank, C_IN, A, B, 700 + A) for A in (2, 1)]
    del pol, tgts, opts
    torch.cuda.empty_cache()
    return {'workload': f'c4 lifting_2_pushing_2: two Q-networks (A=2 and A=1), C={C_IN}, batch {Bg} each = {B}/rank x {world} ranks, '
                        'train_groups (host staging + H2D + both updates + one sync per call)',
            'ms_per_call': ms, 'value': 2 * Bg / (ms * 1e-3), 'unit': UNIT + ' (both networks)', 'train_info': info,
            'parity_check': checks, 'parity_ok': all(c['ok'] for c in checks)}


def same_gpu_reference(B):
    """The UNMODIFIED reference step (oracle/_ref: train.train on policies.DQNPolicy nets) on THIS GPU through PyTorch/cuDNN, as the
    reference runs whenever a GPU is visible (train.py:23-24: cudnn.benchmark, device = cuda).  fp32 = allow_tf32 False (the only
    parity-grade library setting); tf32 = PyTorch's default for convolutions (misses the 1e-3 bar in train mode)."""
    from oracle import ref_runner as R
    if not R.available():
        return {'error': 'oracle/_ref not staged'}
    out = {}
    for name, tf32 in (('fp32', False), ('tf32', True)):
        out[name] = R.time_train(C_IN, A_OUT, B, GAMMA, 6, 3, TERMINAL_EVERY, allow_tf32=tf32)
    out['note'] = ("the reference's own train.train (oracle/_ref, unmodified) on cuda tensors: host staging + H2D + eager cuDNN / ATen kernels + two "
                   '.item() syncs per call, i.e. comparable with `e2e` / `train_call` of this line')
    return out


def argmax_check(dev, B=16):
    """Raw per-sample arg-max equality count against the CPU oracle (no near-tie escape hatch), eval and train mode."""
    import torch
    from oracle import fcn_oracle as O
    from spatial_intention_maps_b200 import networks, synth
    st = O.make_state(C_IN, A_OUT, 105)
    net = networks.FCN(C_IN, A_OUT, max_batch=B)
    net.load_state_dict(st)
    net = net.to(dev)
    x = O.hwc_to_nchw(list(synth.synth_states(B, C_IN, 105)))
    out = {'samples': B}
    for name, training in (('eval', False), ('train', True)):
        net.train(training)
        with torch.no_grad():
            ref = O.forward(O.clone_state(st), x, training)
            q = net(x.to(dev)).cpu()
        out[f'{name}_argmax_equal'] = int((q.view(B, -1).argmax(1) == ref.view(B, -1).argmax(1)).sum())
        out[f'{name}_qmap_maxrel'] = float((q - ref).abs().max() / ref.abs().max())
    del net
    torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_simq(args):
    import ctypes as C
    import random
    import types
    import torch
    import torch.distributed as dist
    from spatial_intention_maps_b200 import _lib, networks, synth, train as T
    from spatial_intention_maps_b200.replay import ReplayBuffer

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus and world > 1:
        raise SystemExit(f'--gpus {args.gpus} but WORLD_SIZE={world}')
    if args.gpus > 1 and world == 1:
        raise SystemExit('for --gpus N > 1 launch with: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        # rank 0 prints ONE JSON line on stdout: NCCL's own output (version banner, warnings) goes to stderr.  NCCL honours
        # NCCL_DEBUG_FILE only above the VERSION level, and prints its banner at VERSION and WARN
        if os.environ.get('NCCL_DEBUG', '').upper() in ('', 'VERSION'):
            os.environ['NCCL_DEBUG'] = 'WARN'
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        dist.init_process_group('nccl', device_id=dev)

    B = args.batch
    if args.global_batch:                            # strong scaling as the MAIN line
        if args.global_batch % world:
            raise SystemExit(f'--global-batch {args.global_batch} is not divisible by {world} ranks')
        B = args.global_batch // world
    L = _lib.lib()
    warm = max(2 * N_BATCHES, args.warmup)           # every rotated batch is replayed from its own captured graph: eager + capture first

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def workload(Bw, Cw, Aw, seed):
        pol, tgt, opt = make_nets(networks, torch, dev, Cw, Aw, Bw)          # replicas are made identical by the first DP step (sync_replicas)
        batches = [synth.synth_batch(Bw, Cw, Aw, seed + 97 * rank + i, terminal_every=TERMINAL_EVERY, uniform=args.uniform_input)
                   for i in range(N_BATCHES)]
        hbs = [T.HostBatch(Bw, Cw).fill(b) for b in batches]
        dbs = [T.DeviceBatch(Bw, Cw, dev).upload(h) for h in hbs]
        torch.cuda.synchronize()
        return pol, tgt, opt, batches, hbs, dbs

    pol, tgt, opt, batches, hbs, dbs = workload(B, C_IN, A_OUT, 1234)

    def step_device(i):
        T.train_step_device(pol, tgt, opt, dbs[i % N_BATCHES], B, GAMMA, 100, True)

    db_e = T.DeviceBatch(B, C_IN, dev)

    def step_e2e(i):
        db_e.upload(hbs[i % N_BATCHES])
        T.train_step_device(pol, tgt, opt, db_e, B, GAMMA, 100, True)
        db_e.out2_host.copy_(db_e.out2, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(db_e.out2_host[0])

    for i in range(warm):
        step_device(i)
    # ---- device-resident throughput ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = pol.ctx(B).launches()
    ms_dev = timed(step_device, args.steps)          # each step replays one CUDA graph
    launches = pol.ctx(B).launches() - launches0
    # ---- end to end through host buffers ----
    for i in range(3):
        step_e2e(i)
    ms_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    loss = float(db_e.out2_host[0])
    # per-kernel-class CUDA-event timing of K more steps (events around every tensor-core launch force the eager launch path
    # and the serial schedule, so this pass is separate from the one that defines `value`)
    L.simq_profile(1, None, None, None)
    ms_prof = timed(step_device, args.steps)
    pm, pf, pl, pi = (C.c_double * 2)(), (C.c_double * 2)(), (C.c_longlong * 2)(), (C.c_double * 2)()
    L.simq_profile(0, pm, pf, pl)
    L.simq_profile_issued(pi)
    # ---- the literal user call train.train(cfg, ..., batch of numpy arrays, ...) incl. host staging ----
    cfg = types.SimpleNamespace(batch_size=B, grad_norm_clipping=100, use_double_dqn=True)

    def train_call(i):
        T.train(cfg, pol, tgt, opt, batches[i % N_BATCHES], None, GAMMA)
    train_call(0)
    ncall = max(2, args.steps // 2)
    barrier()
    t0 = time.perf_counter()
    for i in range(ncall):
        train_call(i)
    ms_call = (time.perf_counter() - t0) * 1e3 / ncall
    # ---- e2e with the device-resident replay buffer: minibatch gathered in HBM, no 47 MB upload ----
    replay = None
    try:
        buf = ReplayBuffer(N_BATCHES * B, dev)
        for b in batches:
            for j in range(B):
                buf.push(b.state[j], b.action[j], b.reward[j], b.next_state[j])
        random.seed(5)
        samples = [buf.sample(B) for _ in range(N_BATCHES)]

        def replay_call(i):
            T.train(cfg, pol, tgt, opt, samples[i % N_BATCHES], None, GAMMA)
        for i in range(N_BATCHES + 1):
            replay_call(i)
        ms_rep = timed(replay_call, ncall) / ncall
        replay = {'value': world * B / (ms_rep * 1e-3), 'unit': UNIT, 'ms_per_step': ms_rep,
                  'note': 'train.train on a replay.ReplayBuffer sample: device gather (simq_gather_rows) + step + sync + 8-byte D2H; '
                          'the 2*B states are uploaded once, when the environment produces them'}
        del buf, samples
    except Exception as e:  # noqa: BLE001
        replay = {'error': f'{type(e).__name__}: {e}'}
    # ---- the same step with three operand terms in EVERY backward GEMM (simq_set_backward_terms(3, 3, 0)) ----
    full3 = None
    if not args.no_serial:
        pol.set_backward_terms(3, 3, 0)
        for i in range(warm):
            step_device(i)
        ms_full3 = timed(step_device, args.steps)
        pol.set_backward_terms(2, 2, 512)
        for i in range(warm):
            step_device(i)
        full3 = {'ms_per_step': ms_full3 / args.steps, 'value': world * B * args.steps / (ms_full3 * 1e-3), 'unit': UNIT,
                 'note': 'simq_set_backward_terms(3, 3, 0): dy keeps both bf16 planes in every weight-gradient / input-gradient GEMM (3 MMAs per '
                         'product); the default uses dy.hi only in the weight-gradient GEMMs and the layer-4 input-gradient convolutions '
                         '(gradient error vs the float64 twin +6 %, profiles/r2_bwd_terms_probe.md); forward passes identical'}
    # ---- the same step under the serial schedule (one stream; simq_set_schedule) ----
    serial = None
    if not args.no_serial:
        pol.set_schedule('serial')
        for i in range(warm):
            step_device(i)
        ms_serial = timed(step_device, args.steps)
        pol.set_schedule('lanes')
        for i in range(warm):
            step_device(i)
        serial = {'ms_per_step': ms_serial / args.steps, 'value': world * B * args.steps / (ms_serial * 1e-3), 'unit': UNIT}
    # ---- forward + backward only (the literal BASELINE.json wording) and forward only ----
    x = dbs[0].s.permute(0, 3, 1, 2)
    q_grad = torch.zeros((B, A_OUT, 96, 96), device=dev)
    q_grad.view(B, -1)[:, 7] = 1.0 / B

    def fwd_bwd(i):
        opt.zero_grad(set_to_none=True)
        q = pol(x)
        q.backward(q_grad)
    fwd_bwd(0)
    nfb = max(4, args.steps // 2)
    ms_fb = timed(fwd_bwd, nfb) / nfb

    def fwd_only(i):
        with torch.no_grad():
            pol(x)
    fwd_only(0)
    ms_f = timed(fwd_only, nfb) / nfb
    # measured all-reduce of the flat gradient vector alone (N > 1)
    allreduce_ms = None
    if world > 1:
        g = pol.flat_grad_ext()
        for _ in range(3):
            dist.all_reduce(g, op=dist.ReduceOp.AVG)
        allreduce_ms = timed(lambda i: dist.all_reduce(g, op=dist.ReduceOp.AVG), 10) / 10
    del q_grad, db_e
    workspace_gb = float(L.simq_workspace_bytes(pol.ctx(B).handle)) / 1e9
    del pol, tgt, opt, dbs
    torch.cuda.empty_cache()

    # ---- multi-rank numerical parity on the hardware (outside every timed region) ----
    parity = None
    if world > 1:
        try:
            parity = dp_parity_check(dev, world, rank, C_IN, A_OUT, 32, 900)
        except Exception as e:  # noqa: BLE001
            parity = {'ok': False, 'error': f'{type(e).__name__}: {e}'}
    c4 = None
    if world == 2 and not args.no_c4:
        try:
            c4 = c4_groups(dev, world, rank, max(4, args.steps // 4))
        except Exception as e:  # noqa: BLE001
            c4 = {'parity_ok': False, 'error': f'{type(e).__name__}: {e}'}

    # ---- strong scaling (config c5: 1024 samples per step over ALL ranks), every N incl. 1 ----
    strong = None
    if not args.no_strong and not args.global_batch and 1024 % world == 0:
        try:
            Bs = 1024 // world
            p2, t2, o2, _, _, d2 = workload(Bs, C_IN, A_OUT, 4321)
            for i in range(2 * N_BATCHES):
                T.train_step_device(p2, t2, o2, d2[i % N_BATCHES], Bs, GAMMA, 100, True)
            ns = max(4, 512 // Bs)
            ms_s = timed(lambda i: T.train_step_device(p2, t2, o2, d2[i % N_BATCHES], Bs, GAMMA, 100, True), ns) / ns
            strong = {'global_batch': 1024, 'batch_per_gpu': Bs, 'n_gpus': world, 'ms_per_step': ms_s, 'value': 1024 / (ms_s * 1e-3),
                      'unit': UNIT, 'scaling': 'strong', 'steps': ns,
                      'workload': 'c5 lifting_4-large_empty synthetic: 1024 samples per step over all ranks (network of c3: C=5, A=1)'}
            del p2, t2, o2, d2
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001
            strong = {'error': f'{type(e).__name__}: {e}'}

    # ---- the north_star's headline shape c* (C=8 input channels, A=2), same step, same batch, single GPU ----
    cstar = None
    if not args.no_cstar and world == 1:
        try:
            p3, t3, o3, _, _, d3 = workload(B, 8, 2, 8765)
            for i in range(2 * N_BATCHES):
                T.train_step_device(p3, t3, o3, d3[i % N_BATCHES], B, GAMMA, 100, True)
            n3 = max(4, args.steps // 2)
            ms3 = timed(lambda i: T.train_step_device(p3, t3, o3, d3[i % N_BATCHES], B, GAMMA, 100, True), n3) / n3
            cstar = {'workload': f'c* C=8 A=2 gamma={GAMMA} batch={B} double-DQN', 'ms_per_step': ms3, 'value': B / (ms3 * 1e-3), 'unit': UNIT}
            del p3, t3, o3, d3
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001
            cstar = {'error': f'{type(e).__name__}: {e}'}

    amax = None
    same_gpu = None
    if world == 1:
        try:
            amax = argmax_check(dev)
        except Exception as e:  # noqa: BLE001
            amax = {'error': f'{type(e).__name__}: {e}'}
        if not args.no_torch_gpu:
            try:
                same_gpu = same_gpu_reference(B)
            except Exception as e:  # noqa: BLE001
                same_gpu = {'error': f'{type(e).__name__}: {e}'}
            torch.cuda.empty_cache()

    if rank == 0:
        sustained, burst, hbm, how = peaks()
        traffic, traffic_note = None, None
        tp = os.path.join(ROOT, 'profiles', 'roofline_traffic.json')
        if os.path.exists(tp):
            tj = json.load(open(tp))
            traffic = tj['dram_bytes']
            traffic_note = (f"{tj['launch']}: {tj['dram_bytes'] / 1e6:.0f} MB DRAM vs {tj['algorithmic_bytes'] / 1e6:.0f} MB algorithmic, "
                            f"tensor pipe {tj['tensor_pipe_active_pct']:.0f} % active ({tj['source']})")
        conv_tf = (pf[0] / (pm[0] * 1e-3)) / 1e12 if pm[0] > 0 else 0.0
        wgrad_tf = (pf[1] / (pm[1] * 1e-3)) / 1e12 if pm[1] > 0 else 0.0
        per_step = ms_dev / args.steps
        value = world * B * args.steps / (ms_dev * 1e-3)
        e2e = world * B * args.steps / (ms_e2e * 1e-3)
        e2e_ms = ms_e2e / args.steps
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': warm,
            'ms_per_step': per_step, 'higher_is_better': True, 'scaling': 'strong' if args.global_batch else 'weak', 'vs_baseline': None,
            'dtype': 'bf16 hi+lo split operands, f32 accumulate: 3 tcgen05 MMAs per product in every forward pass (Q-map / arg-max parity); '
                     'backward: 2 MMAs (dy.hi x split operand) in the weight-gradient GEMMs and the layer-4 input-gradient convolutions, 3 elsewhere',
            'data': 'synthetic (U[0,1) states)' if args.uniform_input else 'synthetic',
            'config': {'workload': workload_name(B), 'global_batch': world * B, 'parallelism': f'dp{world}',
                       'batches_rotated': N_BATCHES,
                       'l2': 'no flush: a step streams >3 GB of activations per GPU through the 126 MB L2, evicting the 47 MB batch',
                       'step_gflop_per_sample': STEP_GFLOP, 'workspace_gb': workspace_gb,
                       'parity_check': parity, 'strong_scaling': strong, 'c4_two_networks_2gpu': c4, 'c_star': cstar},
            'e2e': {'value': e2e, 'unit': UNIT, 'h2d_bytes_per_step': hbs[0].h2d_bytes(), 'd2h_bytes_per_step': 8, 'ms_per_step': e2e_ms,
                    'train_call': {'ms_per_call': ms_call, 'value': world * B / (ms_call * 1e-3), 'unit': UNIT,
                                   'note': 'wall clock of train.train(cfg, policy_net, target_net, optimizer, Transition of numpy arrays, ...) on rank 0: '
                                           'threaded staging into pinned memory + the e2e step + sync'},
                    'device_replay': replay, 'allreduce_ms': allreduce_ms},
            'gpu_launches': int(launches),
            'roofline': {'bound': 'tensor', 'kernel': 'conv2w_umma_kernel / conv2_umma_kernel / conv_umma_kernel (tcgen05 conv + dgrad, all launches of the step)',
                         'achieved': conv_tf, 'peak': sustained / 1.0, 'unit': 'TFLOP/s', 'frac': conv_tf / sustained, 'traffic': traffic,
                         'traffic_note': traffic_note, 'peak_source': f'{how} bf16 sustained',
                         'launches': int(pl[0]), 'ms_per_step_in_kernel': pm[0] / args.steps,
                         'share_of_step': (pm[0] / args.steps) / (ms_prof / args.steps), 'ms_per_step_profiled_pass': ms_prof / args.steps,
                         'profiled_pass': 'eager launches, serial schedule (one stream), CUDA events around every tensor-core launch: a '
                                          'kernel\'s events time that kernel alone; `value` comes from the graphed multi-lane pass',
                         'note': 'achieved = ALGORITHMIC FLOPs (2*valid_pixels*N*K*taps) / time; the kernels issue 3 (forward, most dgrads) or 2 '
                                 '(layer-4 dgrads) bf16 MMAs per product over 625/576 padded rows: issued_frac counts those',
                         'issued_tflops': (pi[0] / (pm[0] * 1e-3)) / 1e12 if pm[0] > 0 else 0.0,
                         'issued_frac': (pi[0] / (pm[0] * 1e-3)) / 1e12 / sustained if pm[0] > 0 else 0.0,
                         'algorithmic_ceiling_frac': pf[0] / pi[0] if pi[0] > 0 else None,
                         'wgrad_kernel': {'achieved': wgrad_tf, 'frac': wgrad_tf / sustained, 'launches': int(pl[1]),
                                          'ms_per_step_in_kernel': pm[1] / args.steps,
                                          'issued_frac': (pi[1] / (pm[1] * 1e-3)) / 1e12 / sustained if pm[1] > 0 else 0.0},
                         'step_tflops_algorithmic': STEP_GFLOP * 1e-3 * value,
                         'step_frac_of_peak': STEP_GFLOP * 1e-3 * value / world / sustained,
                         'fwd_bwd_only': {'value': world * B / (ms_fb * 1e-3), 'unit': UNIT, 'ms': ms_fb,
                                          'note': 'the literal BASELINE.json metric: online forward on s + backward through the autograd.Function route'},
                         'forward_only': {'value': world * B / (ms_f * 1e-3), 'unit': UNIT, 'ms': ms_f, 'note': 'train-mode BN, no_grad'},
                         'serial_schedule': serial, 'three_term_backward': full3},
            'clocks': clocks, 'loss': loss,
        }
        if world == 1:
            cb = None
            if not args.no_cpu:
                try:
                    cb = cpu_leg_subprocess()
                except Exception as e:  # noqa: BLE001
                    cb = {'value': None, 'unit': UNIT, 'cores': os.cpu_count(), 'kind': 'reference', 'sample': f'failed: {type(e).__name__}: {e}'}
            cb = cb or {'value': None, 'unit': UNIT, 'cores': os.cpu_count(), 'kind': 'reference', 'sample': 'skipped (--no-cpu)'}
            cb['argmax_check_vs_cpu_oracle'] = amax
            if same_gpu is not None:
                cb['same_gpu_torch_cudnn'] = same_gpu
                for k in ('fp32', 'tf32'):
                    if isinstance(same_gpu.get(k), dict):
                        same_gpu[k]['this_framework_speedup'] = same_gpu[k]['ms_per_step'] / ms_call
            line['cpu_baseline'] = cb
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='simq', choices=['simq', 'reference'])
    ap.add_argument('--batch', type=int, default=128, help='per-GPU minibatch')
    ap.add_argument('--global-batch', type=int, default=0, help='strong scaling as the main line: this many samples per step over ALL ranks; '
                    'default 0 = weak scaling with --batch per GPU (a c5 strong-scaling sub-record is measured either way)')
    ap.add_argument('--uniform-input', action='store_true', help='U[0,1) states instead of the modelled overhead / distance / intention maps')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--cpu-leg', action='store_true', help='(internal) --impl reference as the bounded cpu_baseline leg of the GPU arm')
    ap.add_argument('--no-torch-gpu', action='store_true', help='skip the same-GPU PyTorch/cuDNN run of the reference step')
    ap.add_argument('--no-serial', action='store_true', help='skip the serial-schedule measurement')
    ap.add_argument('--no-cstar', action='store_true', help='skip the C=8 / A=2 measurement')
    ap.add_argument('--no-strong', action='store_true', help='skip the c5 strong-scaling sub-record')
    ap.add_argument('--no-c4', action='store_true', help='skip the c4 two-network record (2 GPUs only)')
    ap.add_argument('--quick', action='store_true', help='only the headline measurements (profiling runs under ncu)')
    args = ap.parse_args()
    if args.quick:
        args.no_cpu = args.no_torch_gpu = args.no_serial = args.no_cstar = args.no_strong = args.no_c4 = True
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_simq(args)


if __name__ == '__main__':
    main()
