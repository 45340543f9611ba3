"""Knock-out timing of the tcgen05 wgrad kernel (DLWPCS_WG_KNOCK=<mask>) for two layer shapes; graph-timed."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
from dlwp_cs_b200 import _lib
_lib.load()
B = 32
dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(0)
res = {}
for (n, ci, co) in ((48, 32, 32), (24, 64, 64), (24, 128, 64), (48, 64, 32)):
    d = _lib.make_desc(B, n, ci, co, (3, 3), (1, 1), (1, 1), 1, False, True, False, True, _lib.ACT_CAPPED_LEAKY_RELU, 0.1,
                       10.0, _lib.BF16, _lib.BF16)
    x = torch.randn(B, 6, n, n, ci, generator=g).to(dev).bfloat16()
    y = torch.randn(B, 6, n, n, co, generator=g).to(dev).bfloat16()
    dy = torch.randn(B, 6, n, n, co, generator=g).to(dev).bfloat16()
    fn = lambda: _lib.conv2d_wgrad(d, x, dy, y)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(10):
            fn()
    gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); e1.record(); e1.synchronize()
    res['%d:%d->%d' % (n, ci, co)] = round(100 * e0.elapsed_time(e1), 1)
print(json.dumps({'knock': os.environ.get('DLWPCS_WG_KNOCK', '0'), 'wgrad+reduce_us': res}))
