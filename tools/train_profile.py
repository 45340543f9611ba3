"""Kernel-level breakdown of one training step (unet2 C48, batch 32, bf16) with torch.profiler (CUPTI)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
from dlwp_cs_b200 import _lib
from dlwp_cs_b200.unet import CubeSphereUNet2
from dlwp_cs_b200.train import DataParallelTrainer
_lib.load()
dev = torch.device('cuda:0')
tb = int(os.environ.get('BATCH', '32'))
dt = torch.bfloat16 if os.environ.get('DTYPE', 'bf16') == 'bf16' else torch.float32
m = CubeSphereUNet2(18, 14, base=32).to(dev)
torch.manual_seed(1)
tr = DataParallelTrainer(m, lr=1e-3)
g = torch.Generator().manual_seed(1)
xs = torch.randn(tb, 6, 48, 48, 18, generator=g).to(dev).to(dt)
ts = torch.randn(tb, 6, 48, 48, 14, generator=g).to(dev).to(dt)
for _ in range(3):
    tr.step(xs, ts)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        tr.step(xs, ts)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=40, max_name_column_width=90))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    tr.step(xs, ts)
e1.record(); e1.synchronize()
print('ms/step', e0.elapsed_time(e1) / 5, 'samples/s', tb * 5e3 / e0.elapsed_time(e1))
