"""One model step of the bf16 rollout engine, layer by layer with a synchronize after each (debugging aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
from dlwp_cs_b200 import _lib
from dlwp_cs_b200.unet import CubeSphereUNet2, RolloutEngine
_lib.load()
dev = torch.device('cuda:0')
n = int(os.environ.get('N_FACE', '48'))
batch = int(os.environ.get('BATCH', '64'))
model = CubeSphereUNet2(18, 14, base=32).to(dev)
eng = RolloutEngine(model, batch, n, 2, forcing_channels=4, dtype=torch.bfloat16, use_graph=False)
g = torch.Generator().manual_seed(0)
eng.load_inputs(torch.randn(batch, 6, n, n, 14, generator=g), torch.rand(batch, 6, n, n, 4, generator=g))
torch.cuda.synchronize()
for name, d, s0, s1, dst, packed in eng.plan:
    out = eng.ring[0] if dst == 'out' else eng.buf[dst]
    _lib.conv2d_fwd(d, eng._src(s0, 0), eng._src(s1, 0), packed, out=out)
    torch.cuda.synchronize()
    print(name, 'ok', float(out.float().abs().max()), flush=True)
