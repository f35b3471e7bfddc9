"""Latency probes on the GPU box: train step at several batch sizes, batch-1 policy step."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from spatial_intention_maps_b200 import networks, synth, train as T, policies
from tests.gpu_checks import Cfg

dev = torch.device('cuda', 0)
for B in [int(a) for a in sys.argv[1:]] or [32, 128]:
    pol = networks.FCN(5, 1, max_batch=B).to(dev).train()
    tgt = networks.FCN(5, 1, max_batch=B); tgt.load_state_dict(pol.state_dict()); tgt = tgt.to(dev).eval()
    opt = torch.optim.SGD(pol.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    batch = synth.synth_batch(B, 5, 1, 1, terminal_every=max(2, B // 2))
    hb = T.HostBatch(B, 5).fill(batch); db = T.DeviceBatch(B, 5, dev).upload(hb)
    for _ in range(5): T.train_step_device(pol, tgt, opt, db, B, 0.85, 100, True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(20): T.train_step_device(pol, tgt, opt, db, B, 0.85, 100, True)
    t_issue = (time.perf_counter() - t0) / 20
    e1.record(); torch.cuda.synchronize()
    gpu = e0.elapsed_time(e1) / 20
    cfg = Cfg(B, 5)
    t0 = time.perf_counter()
    for _ in range(10): T.train(cfg, pol, tgt, opt, batch, None, 0.85)
    full = (time.perf_counter() - t0) / 10
    print(f'B={B}: device step {gpu:.2f} ms ({B / gpu * 1e3:.0f} samples/s), host issue time {t_issue * 1e3:.2f} ms/step, '
          f'train() wall incl. batch assembly {full * 1e3:.2f} ms ({B / full:.0f} samples/s)', flush=True)
    del pol, tgt, opt
    torch.cuda.empty_cache()

pol = policies.DQNPolicy(Cfg(1, 5), train=False)
s = synth.synth_states(1, 5, 0)[0]
for _ in range(5): pol.step([[s]], exploration_eps=0.0)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50): pol.step([[s]], exploration_eps=0.0)
print(f'policy.step batch 1: {(time.perf_counter() - t0) / 50 * 1e3:.3f} ms per call (host wall, incl. H2D of the state and D2H of the action)')
