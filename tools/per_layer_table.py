"""Per-layer roofline tables from an ncu launch list (tools/launch_summary.load) of one eager train step:

  * forward : the first train-mode forward that starts inside the list (stem_im2col .. head_up2) -- with the default bench
              workload that is the online pass on s' (126 of the 128 samples are non-terminal),
  * backward: up2_adj .. strip_stem (dgrad and wgrad GEMMs in issue order, the split reduction counted with its wgrad).

    python tools/per_layer_table.py profiles/r1_v11_launches.csv [B_backward=128] [B_forward=126]

Bound per launch = max(algorithmic FLOPs / tensor ceiling of that launch, compulsory bytes / measured copy bandwidth); the
fraction printed is bound time / measured time.  The tensor ceiling of a launch is the measured sustained bf16 peak divided by
the MMAs it issues per algorithmic product: operand terms (3, or 2 for the weight-gradient GEMMs and the layer-4 input-gradient
convolutions: last template argument of the kernel name) x 625/576 (pitch-25 halo rows).  ncu per-launch times are cold-cache, serialised and NOT under the
steady-state power cap of the whole step: fractions above 1.0 against the SUSTAINED peak are expected for the big GEMMs.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from launch_summary import load  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = load(sys.argv[1])
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
BF = int(sys.argv[3]) if len(sys.argv) > 3 else 126
names = [r[0] for r in rows]
pk = os.path.join(ROOT, 'MEASURED_PEAKS.json')
peaks = json.load(open(pk)) if os.path.exists(pk) else {}
PEAK, HBM = peaks.get('bf16_tflops_sustained', 1408.1), peaks.get('hbm_gbs', 6446.3)
CEIL = PEAK * 576 / (3 * 625)

STAGES = []                      # (cin, planes, has_ds) per block, network order
inpl = 64
for planes in (64, 128, 256, 512):
    STAGES.append((inpl, planes, inpl != planes))
    STAGES.append((planes, planes, False))
    inpl = planes


def is_gemm(n):
    return (('conv' in n and 'umma' in n) or n.startswith(('wgrad_umma', 'wgrad2_umma', 'wgrad_fma', 'conv_fma'))) and 'reduce' not in n


def terms_of(kernel_name):
    """Operand terms = last template argument of the tcgen05 kernels (conv*_umma_kernel<.., TERMS>, wgrad*_umma_kernel<.., TERMS>)."""
    if '<' not in kernel_name:
        return 3
    try:
        args = kernel_name[kernel_name.index('<') + 1:kernel_name.rindex('>')].split(',')
        return int(args[1] if kernel_name.startswith(('conv2w_umma_kernel', 'conv1w_umma_kernel')) else args[-1])      # conv{1,2}w_umma_kernel<FL, TERMS, BN>
    except (ValueError, IndexError):
        return 3


def emit(title, layers, launches):
    assert len(layers) == len(launches), (title, len(layers), len(launches))
    print(f'\n### {title}\n')
    print('| layer | kernel | terms | us | TFLOP/s (algorithmic) | of its tensor ceiling (%.0f at 3 terms, %.0f at 2) | GB/s (compulsory bytes) | of %.0f GB/s | binding: bound / measured |' % (CEIL, CEIL * 1.5, HBM))
    print('|---|---|---|---|---|---|---|---|---|')
    tt = tb = tf = 0.0
    for (name, fl, byts), (kn, t) in zip(layers, launches):
        terms = terms_of(kn)
        ceil = PEAK * 576 / (terms * 625)
        tfl, gbs = fl / (t * 1e-3) / 1e12, byts / (t * 1e-3) / 1e9
        t_tensor, t_hbm = fl / (ceil * 1e12) * 1e3, byts / (HBM * 1e9) * 1e3          # ms
        bound = max(t_tensor, t_hbm)
        tt += t; tb += bound; tf += fl
        print(f'| {name} | `{kn.split("<")[0]}` | {terms} | {t * 1e3:.1f} | {tfl:.0f} | {tfl / ceil:.2f} | {gbs:.0f} | {gbs / HBM:.2f} | '
              f'{"tensor" if t_tensor >= t_hbm else "HBM"}: **{bound / t:.2f}** |')
    print(f'| all {len(layers)} launches | | | {tt * 1e3:.0f} | {tf / (tt * 1e-3) / 1e12:.0f} | | | | time-weighted: **{tb / tt:.2f}** |')
    return tt, tb


def conv_entry(name, cin, cout, taps, pix, rows_in, k_eff=None):
    fl = 2.0 * pix * (k_eff or cin) * cout * taps
    return (name, fl, rows_in * cin * 4.0 + rows_in * cout * 4.0 + taps * cin * cout * 4.0)


# ---------------- forward ----------------
start = next(i for i, n in enumerate(names) if n.startswith('stem_im2col'))
end = names.index('head_up2_kernel', start)
fwd = [(n, t) for n, t, _ in rows[start:end] if is_gemm(n)]
L = [conv_entry('stem conv1 7x7/2 as GEMM (K=245 padded to 256)', 256, 64, 1, BF * 2304, BF * 2304, 245)]
for cin, planes, ds in STAGES:
    L.append(conv_entry(f'conv3x3 {cin}->{planes}', cin, planes, 9, BF * 576, BF * 625))
    L.append(conv_entry(f'conv3x3 {planes}->{planes}', planes, planes, 9, BF * 576, BF * 625))
    if ds:
        L.append(conv_entry(f'downsample 1x1 {cin}->{planes}', cin, planes, 1, BF * 576, BF * 625))
L += [conv_entry('head conv1 1x1 512->128', 512, 128, 1, BF * 576, BF * 625), conv_entry('head conv2 1x1 128->32 at 48x48', 128, 32, 1, BF * 2304, BF * 2304)]
print(f'Peaks: measured sustained bf16 {PEAK:.0f} TFLOP/s -> tensor ceiling {CEIL:.0f} at 3 MMAs per product, {CEIL * 1.5:.0f} at 2 (625/576 halo padding included); copy bandwidth {HBM:.0f} GB/s.')
ft, fb = emit(f'Train-mode forward (B = {BF})', L, fwd)

# ---------------- backward ----------------
bs = names.index('up2_adj_kernel')
be = names.index('strip_stem_kernel', bs)
ops = []
for n, t, _ in rows[bs:be]:
    if is_gemm(n):
        ops.append([n, t])
    elif n.startswith('wgrad_reduce') and ops:
        ops[-1][1] += t                      # the deterministic split reduction belongs to the wgrad before it


def wg(name, cout, cin, taps, pix, rows_):
    return (name, 2.0 * pix * cout * cin * taps, rows_ * cout * 4.0 + rows_ * cin * 4.0 + taps * cin * cout * 4.0)


def dg(name, cout, cin, taps, pix, rows_):   # dY [rows][cout] -> dX [rows][cin]
    return (name, 2.0 * pix * cout * cin * taps, rows_ * cout * 4.0 + rows_ * cin * 4.0 + taps * cin * cout * 4.0)


P24, R25, P48 = B * 576, B * 625, B * 2304
Lb = [wg('wgrad head conv2 (32 x 128, dy zero-padded to 64 channels)', 32, 128, 1, P48, P48), dg('dgrad head conv2 (K zero-padded to 64)', 32, 128, 1, P48, P48),
      wg('wgrad head conv1 128 x 512 1x1', 128, 512, 1, P24, R25), dg('dgrad head conv1 128->512 1x1', 128, 512, 1, P24, R25)]
for cin, planes, ds in reversed(STAGES):
    Lb.append(wg(f'wgrad conv2 {planes} x {planes} 3x3', planes, planes, 9, P24, R25))
    if ds:
        Lb.append(wg(f'wgrad downsample {planes} x {cin} 1x1', planes, cin, 1, P24, R25))
    Lb.append(dg(f'dgrad conv2 {planes}->{planes} 3x3', planes, planes, 9, P24, R25))
    Lb.append(wg(f'wgrad conv1 {planes} x {cin} 3x3', planes, cin, 9, P24, R25))
    Lb.append(dg(f'dgrad conv1 {planes}->{cin} 3x3', planes, cin, 9, P24, R25))
    if ds:
        Lb.append(dg(f'dgrad downsample {planes}->{cin} 1x1', planes, cin, 1, P24, R25))
Lb.append(wg('wgrad stem 64 x 256 (im2col GEMM)', 64, 256, 1, P48, P48))
bt, bb = emit(f'Backward (B = {B}): dgrad and wgrad GEMMs in issue order (wgrad incl. its split reduction)', Lb, [tuple(o) for o in ops])
print(f'\nForward + backward GEMM launches: {1e3 * (ft + bt):.0f} us measured, {1e3 * (fb + bb):.0f} us at the per-layer bounds -> **{(fb + bb) / (ft + bt):.2f}** of the per-layer roofline (time-weighted).')

# ---------------- HBM-bound elementwise classes of the whole list ----------------
print('\n### Elementwise kernels (whole list)\n')
agg = {}
for n, t, _ in rows:
    if not is_gemm(n):
        a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += t
tot = sum(t for _, t, _ in rows)
print('| kernel | launches | ms | share of the list |')
print('|---|---|---|---|')
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f'| `{k}` | {c} | {t:.3f} | {100 * t / tot:.1f} % |')
