import os, sys, time
sys.path[:0] = ['/root/repo']
import torch
from dlwp_cs_b200 import _lib
from dlwp_cs_b200.unet import CubeSphereUNet2, RolloutEngine
_lib.load()
dev = torch.device('cuda:0')
torch.manual_seed(1)
model = CubeSphereUNet2(18, 14, base=32).to(dev)
B = int(sys.argv[1]); graph = int(sys.argv[2]); steps = int(sys.argv[3])
eng = RolloutEngine(model, B, 48, steps, forcing_channels=4, dtype=torch.bfloat16, use_graph=bool(graph))
g = torch.Generator().manual_seed(0)
eng.load_inputs(torch.randn(B, 6, 48, 48, 14, generator=g), torch.rand(B, 6, 48, 48, 4, generator=g))
torch.cuda.synchronize()
print('built', B, graph, steps, flush=True)
for it in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.launch(); e1.record(); torch.cuda.synchronize()
    print('launch', it, 'ms', e0.elapsed_time(e1), 'us/step', 1e3 * e0.elapsed_time(e1) / steps, flush=True)
if len(sys.argv) > 4:
    import numpy as np
    hs = torch.randn(B, 6, 48, 48, 14, generator=g).bfloat16().pin_memory(); hf = torch.rand(B, 6, 48, 48, 4, generator=g).bfloat16().pin_memory()
    for it in range(2):
        t0 = time.time(); h = eng.run_to_host(hs, hf); torch.cuda.synchronize(); print('to_host', it, time.time() - t0, flush=True)
