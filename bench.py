"""Benchmark of the DQN Q-map training step (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl simq|reference] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one full ``train.train`` update (train.py:108-141) on one synthetic replay minibatch:
online forward on s, train-mode online forward on s' (Double-DQN arg-max), eval-mode target forward,
TD target + SmoothL1, backward, global grad-norm clip, momentum-SGD.  Workload at every N: config c3 of
SURVEY.md §8 (BASELINE.json configs[2]: pushing_4-large_empty, C=5 input channels incl. the intention
map, A=1, gamma 0.85, batch 128 PER GPU -> weak scaling), every 64th transition terminal.

``value``  : samples/s with the batch already resident in HBM (device part of the step only).
``e2e``    : samples/s through the public API call a user makes (``train.train`` semantics): pinned host
             batch -> H2D copies -> step -> D2H read of (loss, td_error), all inside the timed region.
``roofline``: the dominant kernel (tcgen05 conv/dgrad), algorithmic FLOPs / CUDA-event time measured on
             the launching stream during the timed steps, against MEASURED_PEAKS.json's sustained bf16 peak.
``--impl reference``: the reference path on the host cores (oracle port of train.train: the reference is
             Python and /root/reference does not exist on the GPU box), bounded sample of the same workload.
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
# CPU arm: oracle port of train.train on the host cores
# ---------------------------------------------------------------------------------------------
def cpu_steps(steps, warmup, B=CPU_SAMPLE_B):
    import torch
    from oracle import fcn_oracle as O
    from spatial_intention_maps_b200 import synth
    from tests.gpu_checks import batch_tensors
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    pol = O.make_state(C_IN, A_OUT, 0, perturb=False)
    tgt = O.clone_state(pol)
    mom = None
    ts = []
    for i in range(warmup + steps):
        batch = synth.synth_batch(B, C_IN, A_OUT, 1234 + i, terminal_every=8)
        t0 = time.perf_counter()
        r = O.dqn_step(pol, tgt, mom, *batch_tensors(batch), discount=GAMMA)
        mom = r['momentum']
        if i >= warmup:
            ts.append(time.perf_counter() - t0)
    total = sum(ts)
    return B * len(ts) / total, total / len(ts), cores


def torch_gpu_steps(dev, B, steps, warmup, allow_tf32):
    """The reference's step through stock PyTorch library kernels (cuDNN / ATen) on the SAME GPU: the oracle port of
    train.train with its tensors on the device.  Informational ("the library path to beat"): it is
    neither the product path nor the reference arm."""
    import torch
    from oracle import fcn_oracle as O
    from spatial_intention_maps_b200 import synth
    from tests.gpu_checks import batch_tensors
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = allow_tf32
    torch.backends.cudnn.benchmark = True                                   # train.py:23
    try:
        pol = {k: v.to(dev) for k, v in O.make_state(C_IN, A_OUT, 0, perturb=False).items()}
        tgt = {k: v.clone() for k, v in pol.items()}
        batch = synth.synth_batch(B, C_IN, A_OUT, 1234, terminal_every=TERMINAL_EVERY)
        tens = [t.to(dev) for t in batch_tensors(batch)]
        mom = None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(warmup + steps):
            if i == warmup:
                torch.cuda.synchronize(); e0.record()
            r = O.dqn_step(pol, tgt, mom, *tens, discount=GAMMA)
            mom = r['momentum']
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return {'value': B / (ms * 1e-3), 'unit': UNIT, 'ms_per_step': ms}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = old


def cstar_steps(dev, B, steps):
    """The full train step on the north_star's headline shape (SURVEY.md section 8 row c*: C=8, A=2, gamma 0.85, batch B)."""
    import torch
    from spatial_intention_maps_b200 import networks, synth, train as T
    Cs, As = 8, 2
    pol = networks.FCN(Cs, As, max_batch=B).to(dev).train()
    tgt = networks.FCN(Cs, As, max_batch=B)
    tgt.load_state_dict(pol.state_dict())
    tgt = tgt.to(dev).eval()
    opt = torch.optim.SGD(pol.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    hb = T.HostBatch(B, Cs).fill(synth.synth_batch(B, Cs, As, 4321, terminal_every=TERMINAL_EVERY))
    db = T.DeviceBatch(B, Cs, dev).upload(hb)
    for _ in range(3):
        T.train_step_device(pol, tgt, opt, db, B, GAMMA, 100, True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(steps):
        T.train_step_device(pol, tgt, opt, db, B, GAMMA, 100, True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    loss = float(db.out2[0])
    del pol, tgt, opt, db
    torch.cuda.empty_cache()
    return {'workload': f'c* C={Cs} A={As} gamma={GAMMA} batch={B} double-DQN', 'ms_per_step': ms, 'value': B / (ms * 1e-3), 'unit': UNIT,
            'loss': loss}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    # bounded sample: batches of 16 -- the size at which the CPU path is FASTEST per sample (measured on the 16-core box:
    # 38.5 samples/s at batch 16, 29.2 at the config's batch 128, which no longer fits the caches) -- at most 40 timed steps
    steps, warmup, Bc = max(1, min(args.steps, 40)), max(1, min(args.warmup, 3)), CPU_SAMPLE_B
    v, per_step, cores = cpu_steps(steps, warmup, Bc)
    sample = (f'{steps} timed + {warmup} warm-up steps of batch {Bc} (same network C={C_IN} A={A_OUT}, double-DQN, every 8th transition '
              f'terminal), oracle port of train.train, torch CPU fp32, {cores} threads')
    line = {'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': steps, 'warmup': warmup,
            'ms_per_step': per_step * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic', 'config': {'workload': workload_name(args.batch), 'cpu_sample_batch': Bc},
            'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_simq(args):
    import ctypes as C
    import torch
    import torch.distributed as dist
    from spatial_intention_maps_b200 import _lib, networks, synth, train as T

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
    if args.global_batch:                            # strong scaling: fixed global batch split over the ranks
        if args.global_batch % world:
            raise SystemExit(f'--global-batch {args.global_batch} is not divisible by {world} ranks')
        B = args.global_batch // world
    torch.manual_seed(0)
    pol = networks.FCN(C_IN, A_OUT, max_batch=B).to(dev).train()
    tgt = networks.FCN(C_IN, A_OUT, max_batch=B)
    tgt.load_state_dict(pol.state_dict())
    tgt = tgt.to(dev).eval()
    if world > 1:                                    # identical replicas
        dist.broadcast(pol.flat_params, 0); dist.broadcast(tgt.flat_params, 0)
    opt = torch.optim.SGD(pol.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    batch = synth.synth_batch(B, C_IN, A_OUT, 1234 + rank, terminal_every=TERMINAL_EVERY, uniform=args.uniform_input)
    hb = T.HostBatch(B, C_IN).fill(batch)
    db = T.DeviceBatch(B, C_IN, dev).upload(hb)
    L = _lib.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        T.train_step_device(pol, tgt, opt, db, B, GAMMA, 100, True)

    def step_e2e():
        db.upload(hb)
        T.train_step_device(pol, tgt, opt, db, B, GAMMA, 100, True)
        db.out2_host.copy_(db.out2, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(db.out2_host[0])

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(3, args.warmup)):
        step_device()
    # ---- device-resident throughput + per-kernel-class event timing ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = pol.ctx(B).launches()
    ms_dev = timed(step_device, args.steps)          # the step replays as one CUDA graph
    launches = pol.ctx(B).launches() - launches0
    # per-kernel-class CUDA-event timing of the same K steps (events around every tensor-core launch force the
    # eager launch path, so this pass is separate from the one that defines `value`)
    L.simq_profile(1, None, None, None)
    ms_prof = timed(step_device, args.steps)
    pm, pf, pl = (C.c_double * 2)(), (C.c_double * 2)(), (C.c_longlong * 2)()
    L.simq_profile(0, pm, pf, pl)
    # ---- end to end through host buffers ----
    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    loss = float(db.out2_host[0])
    # ---- informational: the literal user call train.train(cfg, ..., batch of numpy arrays, ...) incl. host staging ----
    import types
    cfg = types.SimpleNamespace(batch_size=B, grad_norm_clipping=100, use_double_dqn=True)

    def train_call():
        T.train(cfg, pol, tgt, opt, batch, None, GAMMA)
    train_call()
    t0 = time.perf_counter()
    ncall = max(2, args.steps // 2)
    for _ in range(ncall):
        train_call()
    ms_call = (time.perf_counter() - t0) * 1e3 / ncall
    # ---- informational: the same step under the serial schedule (one stream; simq_set_schedule) ----
    serial = None
    if not args.no_serial:
        pol.set_schedule('serial')
        for _ in range(3):
            step_device()
        ms_serial = timed(step_device, args.steps)
        pol.set_schedule('lanes')
        for _ in range(2):
            step_device()
        serial = {'ms_per_step': ms_serial / args.steps, 'value': world * B * args.steps / (ms_serial * 1e-3), 'unit': UNIT,
                  'note': 'simq_set_schedule(SERIAL): every kernel of the step on one stream; the headline uses the default two-lane '
                          'schedule (forward(s) beside forwards(s\'), weight gradients beside the dgrad chain), bit-identical results'}
    # forward+backward only (the literal BASELINE.json wording), informational
    x = db.s.permute(0, 3, 1, 2)

    def fwd_bwd():
        opt.zero_grad(set_to_none=True)
        q = pol(x)
        q.backward(q_grad)
    q_grad = torch.zeros((B, A_OUT, 96, 96), device=dev)
    q_grad.view(B, -1)[:, 7] = 1.0 / B
    fwd_bwd()
    ms_fb = timed(fwd_bwd, max(2, args.steps // 2)) / max(2, args.steps // 2)

    def fwd_only():
        with torch.no_grad():
            pol(x)
    fwd_only()
    ms_f = timed(fwd_only, max(2, args.steps // 2)) / max(2, args.steps // 2)

    # ---- informational: the north_star's headline shape c* (C=8 input channels, A=2), same step, same batch ----
    cstar = None
    if not args.no_cstar and world == 1:
        try:
            cstar = cstar_steps(dev, B, max(2, args.steps // 2))
        except Exception as e:  # noqa: BLE001
            cstar = {'error': f'{type(e).__name__}: {e}'}

    # ---- informational: opt-in bf16 mode (1 MMA per product; does NOT meet the parity bar, never the headline) ----
    fast = None
    if not args.no_fast and world == 1:          # single-GPU only: informational, must never endanger the multi-rank line
        try:
            pol.eval()
            with torch.no_grad():
                q_par = pol(x[:8])
                pol.set_precision('bf16')
                q_b16 = pol(x[:8])
            pol.train()
            for _ in range(3):
                step_device()
            ms_fast = timed(step_device, args.steps)
            L.simq_profile(1, None, None, None)
            timed(step_device, max(2, args.steps // 4))
            fm, ff, fl_ = (C.c_double * 2)(), (C.c_double * 2)(), (C.c_longlong * 2)()
            L.simq_profile(0, fm, ff, fl_)
            pol.set_precision('parity')
            qerr = float((q_b16 - q_par).abs().max() / q_par.abs().max())
            agree = float((q_b16.view(8, -1).argmax(1) == q_par.view(8, -1).argmax(1)).float().mean())
            fast = {'value': world * B * args.steps / (ms_fast * 1e-3), 'unit': UNIT, 'ms_per_step': ms_fast / args.steps,
                    'conv_tflops_algorithmic': (ff[0] / (fm[0] * 1e-3)) / 1e12 if fm[0] > 0 else None,
                    'qmap_maxnorm_err_vs_parity_mode': qerr, 'argmax_agreement_vs_parity_mode': agree,
                    'note': 'simq_set_precision(BF16): hi planes only, one tcgen05 MMA per product; fails the 1e-3 parity bar -> not the headline'}
        except Exception as e:  # noqa: BLE001
            fast = {'error': f'{type(e).__name__}: {e}'}
            try:
                pol.set_precision('parity')
            except Exception:  # noqa: BLE001
                pass

    # ---- informational: the same step through stock PyTorch library kernels (cuDNN) on this GPU ----
    torch_gpu = None
    if args.torch_gpu and not args.no_torch_gpu and world == 1:
        try:
            del q_grad
            torch.cuda.empty_cache()
            torch_gpu = {'fp32': torch_gpu_steps(dev, B, 5, 2, False), 'tf32': torch_gpu_steps(dev, B, 5, 2, True),
                         'note': 'oracle port of train.train on cuda tensors (cuDNN / ATen eager kernels, cudnn.benchmark as train.py:23, '
                                 'incl. two .item() syncs per step); fp32 = allow_tf32 False (the parity-grade setting), tf32 = PyTorch default '
                                 '(fails the 1e-3 bar in train mode, SURVEY.md 7.2-1)'}
        except Exception as e:  # noqa: BLE001
            torch_gpu = {'error': f'{type(e).__name__}: {e}'}

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
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(3, args.warmup),
            'ms_per_step': per_step, 'higher_is_better': True, 'scaling': 'strong' if args.global_batch else 'weak', 'vs_baseline': None,
            'dtype': 'bf16 hi+lo split operands (3 tcgen05 MMAs per product), f32 accumulate',
            'data': 'synthetic (U[0,1) states)' if args.uniform_input else 'synthetic',
            'config': {'workload': workload_name(B), 'global_batch': world * B, 'parallelism': f'dp{world}',
                       'l2': 'no flush: a step streams >3 GB of activations per GPU through the 126 MB L2, evicting the 47 MB batch',
                       'step_gflop_per_sample': STEP_GFLOP,
                       'workspace_gb': float(L.simq_workspace_bytes(pol.ctx(B).handle)) / 1e9},
            'e2e': {'value': e2e, 'unit': UNIT, 'h2d_bytes_per_step': hb.h2d_bytes(), 'd2h_bytes_per_step': 8,
                    'ms_per_step': ms_e2e / args.steps},
            'gpu_launches': int(launches),
            'roofline': {'bound': 'tensor', 'kernel': 'conv2w_umma_kernel / conv2_umma_kernel / conv_umma_kernel (tcgen05 conv + dgrad, all launches of the step)', 'achieved': conv_tf, 'peak': sustained / 1.0,
                         'unit': 'TFLOP/s', 'frac': conv_tf / sustained, 'traffic': traffic, 'traffic_note': traffic_note, 'peak_source': f'{how} bf16 sustained',
                         'launches': int(pl[0]), 'ms_per_step_in_kernel': pm[0] / args.steps,
                         'share_of_step': (pm[0] / args.steps) / (ms_prof / args.steps), 'ms_per_step_profiled_pass': ms_prof / args.steps,
                         'profiled_pass': 'eager launches, serial schedule (one stream), CUDA events around every tensor-core launch: a '
                                          'kernel\'s events time that kernel alone; `value` comes from the graphed two-lane pass',
                         'note': 'algorithmic FLOPs (2*valid_pixels*N*K*taps); the kernel issues 3 bf16 MMAs per product over 625/576 padded rows, '
                                 'so issued tensor work = 3.26x algorithmic: issued_frac = frac*3.26',
                         'issued_frac': conv_tf * 3 * 625 / 576 / sustained,
                         'parity_mode_ceiling_frac': 576.0 / (3 * 625), 'frac_of_parity_mode_ceiling': conv_tf * 3 * 625 / 576 / sustained,
                         'wgrad_kernel': {'achieved': wgrad_tf, 'frac': wgrad_tf / sustained, 'launches': int(pl[1]),
                                          'ms_per_step_in_kernel': pm[1] / args.steps}},
            'step_tflops_algorithmic': STEP_GFLOP * 1e-3 * value,
            'fwd_bwd_only': {'value': world * B / (ms_fb * 1e-3), 'unit': UNIT, 'ms': ms_fb},
            'forward_only': {'value': world * B / (ms_f * 1e-3), 'unit': UNIT, 'ms': ms_f, 'note': 'train-mode BN, no_grad'},
            'serial_schedule': serial, 'c_star': cstar,
            'train_call': {'ms_per_call': ms_call, 'value': world * B / (ms_call * 1e-3), 'unit': UNIT,
                           'note': 'wall clock of train.train(cfg, policy_net, target_net, optimizer, Transition of numpy arrays, ...) on rank 0: '
                                   'threaded staging into pinned memory (the upload of s starts while s\' is still being staged) + e2e step + sync'},
            'clocks': clocks, 'loss': loss, 'bf16_fast_mode': fast, 'torch_cudnn_same_gpu': torch_gpu,
        }
        if world == 1 and not args.no_cpu:
            try:
                v, per, cores = cpu_steps(24, 2)             # ~10 s of CPU work
                line['cpu_baseline'] = {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                                        'sample': f'24 timed + 2 warm-up steps of batch {CPU_SAMPLE_B} of the same step (oracle port of train.train, torch CPU fp32)'}
            except Exception as e:  # noqa: BLE001
                line['cpu_baseline'] = {'value': None, 'unit': UNIT, 'cores': os.cpu_count(), 'kind': 'port', 'sample': f'failed: {type(e).__name__}: {e}'}
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
    ap.add_argument('--global-batch', type=int, default=0, help='strong scaling: this many samples per step over ALL ranks (e.g. 1024 = config c5); '
                    'default 0 = weak scaling with --batch per GPU')
    ap.add_argument('--uniform-input', action='store_true', help='U[0,1) states instead of the modelled overhead / distance / intention maps')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--no-fast', action='store_true', help='skip the informational bf16-mode measurement')
    ap.add_argument('--torch-gpu', action='store_true', help='also time the step through eager PyTorch/cuDNN kernels on this GPU (informational; '
                    'runs the oracle port on cuda tensors, so it is opt-in: profiles/r1_v14_bench_with_cudnn_comparator.json holds the numbers)')
    ap.add_argument('--no-torch-gpu', action='store_true', help='(default; kept for older command lines)')
    ap.add_argument('--no-serial', action='store_true', help='skip the informational serial-schedule measurement')
    ap.add_argument('--no-cstar', action='store_true', help='skip the informational C=8 / A=2 measurement')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_simq(args)


if __name__ == '__main__':
    main()
