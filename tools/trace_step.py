"""Cross-kernel timeline of one unet2 model step inside the graph-replayed rollout (bf16 tcgen05 path).

    DLWPCS_TC_TRACE=1 python tools/trace_step.py [--batch 64] [--n 48] [--steps 6]

Every CTA of every conv launch records %globaltimer stamps (include/dlwpcs.h: dlwpcs_trace_read).  Prints, for each layer
of the LAST model step, relative to the first CTA entry of that step (microseconds, min / median / max over CTAs):
entry, prologue done, dependency passed, first patch landed, first accumulators, last epilogue, exit -- i.e. where the
time between two layers goes (launch gap, prologue, pipeline fill, tail imbalance)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
os.environ.setdefault('DLWPCS_TC_TRACE', '1')
import numpy as np  # noqa: E402
import torch  # noqa: E402
from dlwp_cs_b200 import _lib  # noqa: E402
from dlwp_cs_b200.unet import CubeSphereUNet2, RolloutEngine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=64)
ap.add_argument('--n', type=int, default=48)
ap.add_argument('--steps', type=int, default=6)
ap.add_argument('--out', default=None)
args = ap.parse_args()
_lib.load()
dev = torch.device('cuda:0')
torch.manual_seed(1)
model = CubeSphereUNet2(18, 14, base=32).to(dev)
eng = RolloutEngine(model, args.batch, args.n, args.steps, forcing_channels=4, dtype=torch.bfloat16, use_graph=True)
g = torch.Generator().manual_seed(0)
eng.load_inputs(torch.randn(args.batch, 6, args.n, args.n, 14, generator=g), torch.rand(args.batch, 6, args.n, args.n, 4, generator=g))
for _ in range(3):
    eng.launch()
torch.cuda.synchronize()
eng.launch()
torch.cuda.synchronize()
tr = _lib.trace_read(reset=False)
# the graph's kernel nodes keep the trace slots they were captured with (after the 11 eager warm-up launches); every
# replay overwrites them, so the buffer holds the stamps of the last replay
names = [p[0] for p in eng.plan]
if getattr(eng, 'fused_head', None) is not None:          # the 1x1 output layer runs inside the launch in front of it
    names = names[:eng.fused_head] + [names[eng.fused_head] + '+' + names[eng.fused_head + 1]] + names[eng.fused_head + 2:]
nl = len(names)
# keep the launches of the last captured model step; subtract in integers (ns since the epoch do not fit a float64)
total = tr.shape[0]
raw = tr[total - nl:total]
valid = raw[:, :, 0] > 0
t0i = raw[:, :, 0][valid].min()
last = np.where(raw > 0, raw - t0i, 0).astype(np.float64)
last[:, :, 7] = raw[:, :, 7]
t0 = 0.0
rows = []
labels = ['entry', 'prologue', 'dep', 'patch0', 'acc0', 'epi_last', 'exit']
print('%-22s %5s | ' % ('layer', 'ctas') + ' | '.join('%-22s' % l for l in labels) + ' | tiles')
prev_exit = None
for i, nm in enumerate(names):
    v = valid[i]
    r = (last[i][v] - t0) / 1e3
    cells = []
    d = {'layer': nm, 'ctas': int(v.sum())}
    for s, l in enumerate(labels):
        col = r[:, s]
        cells.append('%6.1f %6.1f %6.1f' % (col.min(), np.median(col), col.max()))
        d[l] = [round(float(col.min()), 2), round(float(np.median(col)), 2), round(float(col.max()), 2)]
    tiles = last[i][v][:, 7]
    d['tiles'] = [int(tiles.min()), int(tiles.max())]
    if prev_exit is not None:
        d['gap_prev_last_exit_to_first_patch0_us'] = round(float(r[:, 3].min() - prev_exit), 2)
    prev_exit = float(r[:, 6].max())
    rows.append(d)
    print('%-22s %5d | ' % (nm, v.sum()) + ' | '.join(cells) + ' | %d-%d' % (tiles.min(), tiles.max()))
step_us = (last[:, :, 6][valid].max() - t0) / 1e3
busy = sum(float(np.median((last[i][valid[i]][:, 5] - last[i][valid[i]][:, 3]))) for i in range(nl)) / 1e3
print('step: %.1f us; sum over layers of median (last epilogue - first patch) = %.1f us' % (step_us, busy))
if args.out:
    json.dump({'batch': args.batch, 'n': args.n, 'step_us': step_us, 'layers': rows}, open(args.out, 'w'), indent=1)
