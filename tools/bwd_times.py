"""Per-layer device time of forward / dgrad / wgrad of the unet2 layers at training shapes (batch 32, bf16), through the
C ABI.  One JSON line."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
from dlwp_cs_b200 import _lib
from dlwp_cs_b200.unet import unet2_layer_specs
_lib.load()
B = int(os.environ.get('BATCH', '32'))
n0 = int(os.environ.get('N_FACE', '48'))
edges = {'conv_2d_1': 1, 'conv_2d_1_2': 1, 'conv_2d_2': 2, 'conv_2d_2_2': 2, 'conv_2d_5_2': 4, 'conv_2d_5': 4,
         'conv_2d_6_2': 2, 'conv_2d_6': 2, 'conv_2d_7': 1, 'conv_2d_7_2': 1, 'conv_2d_8': 1}
dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(0)


def timed(fn, reps=10):
    """GPU time per call: the calls are captured into a CUDA graph so that Python / allocator overhead is excluded."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    e1.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps


out = {}
tot = [0.0, 0.0, 0.0]
for name, k, ci, co in unet2_layer_specs(18, 14, 32):
    n = n0 // edges[name]
    last = name == 'conv_2d_8'
    d = _lib.make_desc(B, n, ci, co, (k, k), (1, 1), (1, 1), 0 if last else 1, False, True, False, True,
                       _lib.ACT_NONE if last else _lib.ACT_CAPPED_LEAKY_RELU, 0.1, 10.0, _lib.BF16, _lib.BF16)
    x = torch.randn(B, 6, n, n, ci, generator=g).to(dev).bfloat16()
    ws = [(torch.randn(k, k, ci, co, generator=g) * 0.05).to(dev) for _ in range(2)]
    bs = [torch.zeros(co, device=dev) for _ in range(2)]
    packed = _lib.pack_weights(d, ws[0], ws[1], None, bs[0], bs[1], None)
    packed_t = _lib.pack_weights(d, ws[0], ws[1], None, transposed=True)
    y = _lib.conv2d_fwd(d, x, None, packed)
    dy = torch.randn(y.shape, generator=g).to(dev).bfloat16()
    tf = timed(lambda: _lib.conv2d_fwd(d, x, None, packed, out=y))
    td = timed(lambda: _lib.conv2d_dgrad(d, dy, y, packed_t))
    tw = timed(lambda: _lib.conv2d_wgrad(d, x, dy, y))
    out[name] = [round(tf, 1), round(td, 1), round(tw, 1)]
    tot = [tot[0] + tf, tot[1] + td, tot[2] + tw]
print(json.dumps({'batch': B, 'fwd_dgrad_wgrad_us': out, 'total_us': [round(t, 1) for t in tot]}))
