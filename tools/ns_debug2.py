import os, sys, time
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__)))]
import torch
from dlwp_cs_b200 import _lib as lib
lib.load()
B, n, cin, cout = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), 32
reps = int(sys.argv[4])
g = torch.Generator().manual_seed(0)
x = torch.randn(B, 6, n, n, cin, generator=g).bfloat16().cuda()
ws = [(torch.randn(3, 3, cin, cout, generator=g) * 0.1).cuda() for _ in range(2)]
bs = [(torch.randn(cout, generator=g) * 0.1).cuda() for _ in range(2)]
d = lib.make_desc(B, n, cin, cout, (3, 3), (1, 1), (1, 1), 1, False, True, False, True, lib.ACT_CAPPED_LEAKY_RELU, 0.1, 10.0, lib.BF16, lib.BF16)
packed = lib.pack_weights(d, ws[0], ws[1], None, bs[0], bs[1], None)
y = lib.conv2d_fwd(d, x, None, packed)
torch.cuda.synchronize()
print('first ok', flush=True)
ref = y.clone()
for r in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        lib.conv2d_fwd(d, x, None, packed, out=y)
    e1.record(); torch.cuda.synchronize()
    print(r, 'us/launch', 1e3 * e0.elapsed_time(e1) / 50, 'same', bool(torch.equal(y, ref)), flush=True)
