"""Debug helper: per-face / per-pixel error map of one bf16 tensor-core conv case against the oracle."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle')]
import numpy as np
import torch
import cs_oracle as O
from dlwp_cs_b200 import _lib as lib


def run(n, cin, cout, k, halo, batch=2, flip=True, indep=False, seed=0):
    g = torch.Generator().manual_seed(seed + n * 1000 + cin * 10 + cout)
    x = torch.randn(batch, 6, n, n, cin, generator=g).bfloat16()
    nw = 3 if indep else 2
    ws = [(torch.randn(k, k, cin, cout, generator=g) * 0.2).bfloat16().float() for _ in range(nw)]
    bs = [torch.randn(cout, generator=g) * 0.2 for _ in range(nw)]
    w_np = ws[2] if indep else None
    b_np = bs[2] if indep else None
    dd = lambda t: None if t is None else t.double()
    ref = O.cube_sphere_conv2d(O.cube_sphere_pad(x.double(), halo), dd(ws[0]), dd(ws[1]), dd(w_np), dd(bs[0]), dd(bs[1]),
                               dd(b_np), flip_north_pole=flip)
    cu = lambda t: None if t is None else t.cuda()
    d = lib.make_desc(batch, n, cin, cout, (k, k), (1, 1), (1, 1), halo, False, flip, indep, True, lib.ACT_NONE, 0.1, 10.0,
                      lib.BF16, lib.F32)
    packed = lib.pack_weights(d, cu(ws[0]), cu(ws[1]), cu(w_np), cu(bs[0]), cu(bs[1]), cu(b_np))
    for rep in range(3):
        y = lib.conv2d_fwd(d, x.cuda(), None, packed)
        torch.cuda.synchronize()
        err = (y.double().cpu() - ref).abs().amax(dim=-1).numpy()       # (B,6,H,W)
        print('rep', rep, 'max err per (batch, face):')
        print(np.round(err.max(axis=(2, 3)), 4))
    b, f = np.unravel_index(err.max(axis=(2, 3)).argmax(), err.shape[:2])
    print('worst face map (b=%d f=%d):' % (b, f))
    print(np.round(err[b, f], 3))


if __name__ == '__main__':
    flip = '--noflip' not in sys.argv
    indep = '--indep' in sys.argv
    a = [int(v) for v in sys.argv[1:] if not v.startswith('--')]
    run(*a, flip=flip, indep=indep)
