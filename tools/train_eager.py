"""A few eager (no CUDA graph) bf16 training steps of unet2 at C48, batch 32 -- the command ncu profiles for the
backward kernels (wgrad / dgrad / halo scatter-add)."""
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
m = CubeSphereUNet2(18, 14, base=32).to(dev)
torch.manual_seed(1)
tr = DataParallelTrainer(m, lr=1e-3, use_graph=False)
g = torch.Generator().manual_seed(1)
xs = torch.randn(tb, 6, 48, 48, 18, generator=g).to(dev).bfloat16()
ts = torch.randn(tb, 6, 48, 48, 14, generator=g).to(dev).bfloat16()
for _ in range(int(os.environ.get('STEPS', '3'))):
    tr.step(xs, ts)
torch.cuda.synchronize()
print('ok')
