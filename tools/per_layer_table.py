"""Per-layer roofline table of the first train-mode forward found in an ncu launch list (tools/launch_summary.load)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from launch_summary import load  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = load(sys.argv[1])
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
names = [r[0] for r in rows]
start = next(i for i, n in enumerate(names) if n.startswith('stem_im2col_kernel'))
end = names.index('head_up2_kernel', start)
convs = [(n, t) for n, t, _ in rows[start:end] if 'conv' in n and 'umma' in n]
peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {}
PEAK, HBM = peaks.get('bf16_tflops_sustained', 1408.1), peaks.get('hbm_gbs', 6446.3)
CEIL = PEAK * 576 / (3 * 625)
# (name, Cin, Cout, taps, output pixels, input rows, rows are pitch-25?)  bytes: split-bf16 input (4 B/elem) + fp32 output
layers = [('stem conv1 7x7/2 as GEMM (K=245 padded to 256)', 256, 64, 1, B * 2304, B * 2304)]
inpl = 64
for planes in (64, 128, 256, 512):
    for b in range(2):
        cin = inpl if b == 0 else planes
        layers.append((f'conv3x3 {cin}->{planes}', cin, planes, 9, B * 576, B * 625))
        layers.append((f'conv3x3 {planes}->{planes}', planes, planes, 9, B * 576, B * 625))
        if b == 0 and inpl != planes:
            layers.append((f'downsample 1x1 {cin}->{planes}', cin, planes, 1, B * 576, B * 625))
    inpl = planes
layers += [('head conv1 1x1 512->128', 512, 128, 1, B * 576, B * 625), ('head conv2 1x1 128->32 at 48x48', 128, 32, 1, B * 2304, B * 2304)]
assert len(layers) == len(convs), (len(layers), len(convs))
print('| layer | kernel | us | TFLOP/s (algorithmic) | tensor bound: of parity-mode ceiling %.0f | GB/s (compulsory bytes) | HBM bound: of %.0f GB/s | binding |' % (CEIL, HBM))
print('|---|---|---|---|---|---|---|---|')
tt = tf = 0.0
for (name, cin, cout, taps, pix, rows_in), (kn, t) in zip(layers, convs):
    fl = 2.0 * pix * min(cin, 245 if 'stem' in name else cin) * cout * taps
    byts = rows_in * cin * 4.0 + rows_in * cout * 4.0 + taps * cin * cout * 4.0
    tfl, gbs = fl / (t * 1e-3) / 1e12, byts / (t * 1e-3) / 1e9
    t_tensor, t_hbm = fl / (CEIL * 1e12), byts / (HBM * 1e9)
    binding = 'tensor' if t_tensor >= t_hbm else 'HBM'
    frac = (tfl / CEIL) if binding == 'tensor' else (gbs / HBM)
    tt += t; tf += fl
    print(f'| {name} | `{kn.split("<")[0]}` | {t * 1e3:.1f} | {tfl:.0f} | {tfl / CEIL:.2f} | {gbs:.0f} | {gbs / HBM:.2f} | {binding}: **{frac:.2f}** |')
print(f'| all 22 tensor-core launches | | {tt * 1e3:.0f} | {tf / (tt * 1e-3) / 1e12:.0f} | {tf / (tt * 1e-3) / 1e12 / CEIL:.2f} | | | |')
