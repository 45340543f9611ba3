"""Device time of single CubeSphereConv2D launches (bf16, 3x3, halo 1, batch 64) at shapes that isolate the kernels:
prints one JSON line per shape with us per launch, TFLOP/s and algorithmic GB/s (SURVEY.md 8(d) bytes).
    python tools/stress_shape.py            # row-streamed kernel where it applies
    DLWPCS_RS=0 python tools/stress_shape.py    # classic kernel"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
from dlwp_cs_b200 import _lib
_lib.load()
dev = torch.device('cuda:0')
batch = int(os.environ.get('BATCH', '64'))
shapes = [(48, 32, 32), (48, 64, 32), (48, 64, 64), (24, 64, 64), (96, 32, 32)]
if os.environ.get('SHAPES'):
    shapes = [tuple(int(v) for v in t.split(',')) for t in os.environ['SHAPES'].split(';')]
reps = int(os.environ.get('REPS', '7'))
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for n, cin, cout in shapes:
    b = batch if n <= 48 else batch // 4
    g = torch.Generator().manual_seed(n + cin)
    x = torch.randn(b, 6, n, n, cin, generator=g).bfloat16().to(dev)
    w = [(torch.randn(3, 3, cin, cout, generator=g) * 0.05).to(dev) for _ in range(2)]
    bs = [torch.zeros(cout, device=dev) for _ in range(2)]
    d = _lib.make_desc(b, n, cin, cout, (3, 3), (1, 1), (1, 1), 1, False, True, False, True, _lib.ACT_CAPPED_LEAKY_RELU, 0.1, 10.0,
                       _lib.BF16, _lib.BF16)
    packed = _lib.pack_weights(d, w[0], w[1], None, bs[0], bs[1], None)
    y = _lib.conv2d_fwd(d, x, None, packed)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(4):                     # input + output of four launches: > L2 at the larger shapes anyway
            _lib.conv2d_fwd(d, x, None, packed, out=y)
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) / 4)
    ms = sorted(ts)[len(ts) // 2]
    flop = 2.0 * b * 6 * n * n * 9 * cin * cout
    byts = b * 6 * n * n * (cin + cout) * 2
    print(json.dumps({'rs': os.environ.get('DLWPCS_RS', '1'), 'n': n, 'batch': b, 'cin': cin, 'cout': cout, 'us': round(1e3 * ms, 1),
                      'tflops': round(flop / ms / 1e9, 1), 'gbs': round(byts / ms / 1e6, 1)}))
