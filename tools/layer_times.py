"""Per-layer device time of one unet2 model step at batch 64 (bf16 tcgen05 path); prints one JSON line.
Used with DLWPCS_TC_KNOCK=<mask> for knock-out bottleneck analysis (timings under a knock-out are not results)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
import bench as B
from dlwp_cs_b200 import _lib
from dlwp_cs_b200.unet import CubeSphereUNet2, RolloutEngine
_lib.load()
dev = torch.device('cuda:0')
n = int(os.environ.get('N_FACE', '48'))
batch = int(os.environ.get('BATCH', '64'))
model = CubeSphereUNet2(18, 14, base=32).to(dev)
torch.manual_seed(1)
eng = RolloutEngine(model, batch, n, 2, forcing_channels=4, dtype=torch.bfloat16, use_graph=False)
g = torch.Generator().manual_seed(0)
eng.load_inputs(torch.randn(batch, 6, n, n, 14, generator=g), torch.rand(batch, 6, n, n, 4, generator=g))
eng.launch()
torch.cuda.synchronize()
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
lt = B.time_layers(eng, flush, reps=20)
print(json.dumps({'knock': os.environ.get('DLWPCS_TC_KNOCK', '0'), 'total_us': round(1e3 * sum(m for _, m in lt), 1),
                  'layers_us': {k: round(1e3 * m, 1) for k, m in lt}}))
