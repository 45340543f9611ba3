"""GPU parity: the CUDA path (through the C ABI in libdlwpcs.so) against the oracle and the reference-generated golden
vectors.  float32 bar from BASELINE.json north_star: rtol 1e-5 against the float64 oracle on identical float32 inputs
(atol = 1e-5 x the output's max magnitude, since individual outputs pass through zero).  Index work is bit-exact."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import cs_oracle as O  # noqa: E402
from tests.golden.cases import CONV_CASES  # noqa: E402

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
RTOL = 1e-5


def close(actual, expected, rtol=RTOL):
    a = actual.detach().double().cpu().numpy() if torch.is_tensor(actual) else np.asarray(actual, dtype=np.float64)
    e = expected.detach().double().cpu().numpy() if torch.is_tensor(expected) else np.asarray(expected, np.float64)
    assert a.shape == e.shape, (a.shape, e.shape)
    np.testing.assert_allclose(a, e, rtol=rtol, atol=rtol * max(float(np.abs(e).max()), 1e-30))


@pytest.fixture(scope='module')
def cs():
    import dlwp_cs_b200
    from dlwp_cs_b200 import _lib, functional
    _lib.load()
    return dlwp_cs_b200, _lib, functional


@pytest.mark.parametrize('n,p,c,dtype', [(4, 1, 1, torch.float32), (5, 2, 3, torch.float32), (8, 3, 4, torch.float32),
                                         (48, 1, 18, torch.float32), (48, 1, 32, torch.bfloat16),
                                         (24, 1, 7, torch.bfloat16), (12, 2, 64, torch.float32)])
def test_pad_fwd_bit_exact(cs, n, p, c, dtype):
    _, _, F = cs
    x = torch.randn(2, 6, n, n, c).to(dtype)
    ref = O.cube_sphere_pad(x, p)
    got = F.cube_sphere_pad(x.cuda(), p).cpu()
    assert torch.equal(got, ref)


def test_pad_layer_channels_first_and_empty_batch(cs):
    pkg, _, _ = cs
    x = torch.randn(2, 3, 6, 6, 6)
    got = pkg.CubeSpherePadding2D(2)(x.cuda()).cpu()          # default data_format is channels_first (custom.py:1074)
    assert torch.equal(got, O.cube_sphere_pad(x, 2, 'channels_first'))
    e = pkg.CubeSpherePadding2D(1, data_format='channels_last')(torch.zeros(0, 6, 4, 4, 2).cuda())
    assert tuple(e.shape) == (0, 6, 6, 6, 2)


@pytest.mark.parametrize('n,p,c', [(4, 1, 2), (6, 2, 3), (48, 1, 8)])
def test_pad_bwd_scatter_add(cs, n, p, c):
    _, _, F = cs
    x = torch.randn(2, 6, n, n, c, dtype=torch.float64, requires_grad=True)
    g = torch.randn(2, 6, n + 2 * p, n + 2 * p, c, dtype=torch.float64)
    O.cube_sphere_pad(x, p).backward(g)
    xc = x.detach().float().cuda().requires_grad_(True)
    F.cube_sphere_pad(xc, p).backward(g.float().cuda())
    close(xc.grad, x.grad.float().double(), rtol=1e-6)


@pytest.mark.parametrize('case', CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv_layer_vs_reference_golden(cs, case):
    """CubeSphereConv2D module (both data formats) against outputs of the reference's own call() (tests/golden)."""
    pkg, _, _ = cs
    name, kw, b, h, cin = case
    g = np.load(os.path.join(G, 'conv_cases.npz'))
    names = [k for k in ('equatorial_kernel', 'polar_kernel', 'north_pole_kernel', 'equatorial_bias', 'polar_bias',
                         'north_pole_bias') if name + '.' + k in g.files]
    weights32 = [g[name + '.' + k].astype(np.float32) for k in names]
    x32 = g[name + '.x'].astype(np.float32)
    # float64 oracle on the float32-rounded inputs (identical inputs on both sides)
    t = {k: torch.from_numpy(w).double() for k, w in zip(names, weights32)}
    ref = O.cube_sphere_conv2d(torch.from_numpy(x32).double(), t['equatorial_kernel'], t['polar_kernel'],
                               t.get('north_pole_kernel'), t.get('equatorial_bias'), t.get('polar_bias'),
                               t.get('north_pole_bias'), strides=kw.get('strides', 1), padding=kw.get('padding', 'valid'),
                               dilation=kw.get('dilation_rate', 1), flip_north_pole=kw.get('flip_north_pole', True))
    np.testing.assert_allclose(ref.numpy(), g[name + '.y_cl'], rtol=1e-5, atol=1e-5)   # golden used float64 inputs
    for fmt in ('channels_last', 'channels_first'):
        layer = pkg.CubeSphereConv2D(data_format=fmt, **kw).cuda()
        layer.set_weights(weights32)
        xin = torch.from_numpy(x32)
        if fmt == 'channels_first':
            xin = xin.permute(0, 4, 1, 2, 3).contiguous()
        y = layer(xin.cuda())
        assert tuple(y.shape) == tuple(layer.compute_output_shape(tuple(xin.shape)))
        ycl = y if fmt == 'channels_last' else y.permute(0, 2, 3, 4, 1)
        close(ycl, ref)


def test_cfg1_pad_conv_vs_reference_golden(cs):
    """BASELINE.json configs[0]: C48, 3->3 channels, batch 1; unfused (pad kernel + conv kernel) and fused."""
    pkg, lib, F = cs
    g = np.load(os.path.join(G, 'padconv_cfg1.npz'))
    x = torch.from_numpy(g['x']).cuda()
    w = [torch.from_numpy(g[k]).cuda() for k in ('w_eq', 'w_pol', 'b_eq', 'b_pol')]
    xp = F.cube_sphere_pad(x, 1)
    assert np.array_equal(xp.cpu().numpy(), g['xp'])
    y_unfused = F.cube_sphere_conv2d(xp, w[0], w[1], None, w[2], w[3], None)
    y_fused = F.cube_sphere_conv2d(x, w[0], w[1], None, w[2], w[3], None, halo=1)
    close(y_unfused, g['y'])
    close(y_fused, g['y'])
    assert torch.equal(y_fused, y_unfused)
    # the host-buffer C entry point (numpy in / numpy out)
    d = lib.make_desc(1, 48, 3, 3, halo=1)
    yh = lib.conv2d_fwd_host(d, g['x'], g['w_eq'], g['w_pol'], None, g['b_eq'], g['b_pol'], None)
    close(yh, g['y'])


@pytest.mark.parametrize('n,cin,cout,k,halo', [(48, 18, 32, 3, 1), (24, 64, 64, 3, 1), (12, 128, 64, 3, 1),
                                               (48, 32, 14, 1, 0), (10, 5, 7, 3, 2), (8, 3, 40, 5, 2)])
def test_fused_halo_conv_vs_oracle(cs, n, cin, cout, k, halo):
    _, _, F = cs
    g = torch.Generator().manual_seed(n * 1000 + cin)
    x = torch.randn(2, 6, n, n, cin, generator=g)
    w = [torch.randn(k, k, cin, cout, generator=g) * 0.1 for _ in range(2)]
    bs = [torch.randn(cout, generator=g) * 0.1 for _ in range(2)]
    ref = O.capped_leaky_relu(O.cube_sphere_conv2d(O.cube_sphere_pad(x.double(), halo), w[0].double(), w[1].double(),
                                                   None, bs[0].double(), bs[1].double(), None))
    y = F.cube_sphere_conv2d(x.cuda(), w[0].cuda(), w[1].cuda(), None, bs[0].cuda(), bs[1].cuda(), None, halo=halo,
                             activation=('capped_leaky_relu', 0.1, 10.0))
    close(y, ref)


def test_fused_sources_pool_upsample_concat(cs):
    """The U-Net's pool / upsample / concat folded into the conv's input sampling (Azure/train_cs.py:282-299)."""
    _, lib, _ = cs
    g = torch.Generator().manual_seed(7)
    n = 12
    big = torch.randn(2, 6, 2 * n, 2 * n, 8, generator=g)       # to be average-pooled
    small = torch.randn(2, 6, n // 2, n // 2, 12, generator=g)  # to be upsampled
    same = torch.randn(2, 6, n, n, 4, generator=g)
    w = [torch.randn(3, 3, 16, 8, generator=g) * 0.1 for _ in range(2)]
    b = [torch.randn(8, generator=g) * 0.1 for _ in range(2)]
    cases = [((small, lib.SRC_UP2), (same, lib.SRC_SAME), torch.cat([O.upsample_2x2(small), same], -1)),
             ((big, lib.SRC_POOL2), (big[..., :8], lib.SRC_POOL2), None)]
    # case 1: upsample (+) skip
    (s0, m0), (s1, m1), xin = cases[0]
    d = lib.make_desc(2, n, 16, 8, halo=1, c0=12, mode0=m0, c1=4, mode1=m1)
    packed = lib.pack_weights(d, w[0].cuda(), w[1].cuda(), None, b[0].cuda(), b[1].cuda(), None)
    y = lib.conv2d_fwd(d, s0.cuda(), s1.cuda(), packed)
    ref = O.cube_sphere_conv2d(O.cube_sphere_pad(xin.double(), 1), w[0].double(), w[1].double(), None, b[0].double(),
                               b[1].double(), None)
    close(y, ref)
    # case 2: average pool
    w8 = [t[:, :, :8].contiguous() for t in w]
    d = lib.make_desc(2, n, 8, 8, halo=1, mode0=lib.SRC_POOL2)
    packed = lib.pack_weights(d, w8[0].cuda(), w8[1].cuda(), None, b[0].cuda(), b[1].cuda(), None)
    y = lib.conv2d_fwd(d, big.cuda(), None, packed)
    ref = O.cube_sphere_conv2d(O.cube_sphere_pad(O.avg_pool_2x2(big.double()), 1), w8[0].double(), w8[1].double(), None,
                               b[0].double(), b[1].double(), None)
    close(y, ref)


BWD_CASES = [
    dict(n=8, cin=4, cout=6, k=3, halo=1, act=True),
    dict(n=12, cin=16, cout=32, k=3, halo=1, act=True),
    dict(n=6, cin=3, cout=5, k=3, halo=2, act=False, flip=False),
    dict(n=6, cin=3, cout=5, k=3, halo=1, act=True, indep=True),
    dict(n=9, cin=8, cout=4, k=1, halo=0, act=False),
    dict(n=9, cin=5, cout=4, k=3, halo=0, act=True, same=True),
    dict(n=10, cin=4, cout=4, k=3, halo=0, act=False, dil=2),
    dict(n=7, cin=2, cout=3, k=5, halo=2, act=True, nobias=True),
    dict(n=48, cin=18, cout=32, k=3, halo=1, act=True, batch=1),
]


@pytest.mark.parametrize('c', BWD_CASES, ids=[str(i) for i in range(len(BWD_CASES))])
def test_dgrad_wgrad_vs_oracle_autograd(cs, c):
    _, _, F = cs
    n, cin, cout, k, halo = c['n'], c['cin'], c['cout'], c['k'], c['halo']
    bsz = c.get('batch', 2)
    g = torch.Generator().manual_seed(n * 100 + cin * 10 + cout)
    x = torch.randn(bsz, 6, n, n, cin, generator=g)
    nw = 3 if c.get('indep') else 2
    ws = [torch.randn(k, k, cin, cout, generator=g) * 0.3 for _ in range(nw)]
    bs = None if c.get('nobias') else [torch.randn(cout, generator=g) * 0.3 for _ in range(nw)]
    padding, dil, flip = 'same' if c.get('same') else 'valid', c.get('dil', 1), c.get('flip', True)

    def run(oracle):
        cast = (lambda t: t.double().requires_grad_(True)) if oracle else (lambda t: t.cuda().requires_grad_(True))
        xi, wi = cast(x), [cast(w) for w in ws]
        bi = None if bs is None else [cast(b) for b in bs]
        w_np = wi[2] if nw == 3 else None
        b_eq, b_pol, b_np = (None, None, None) if bi is None else (bi[0], bi[1], bi[2] if nw == 3 else None)
        if oracle:
            y = O.cube_sphere_conv2d(O.cube_sphere_pad(xi, halo), wi[0], wi[1], w_np, b_eq, b_pol, b_np,
                                     padding=padding, dilation=dil, flip_north_pole=flip)
            if c['act']:
                y = O.capped_leaky_relu(y)
        else:
            y = F.cube_sphere_conv2d(xi, wi[0], wi[1], w_np, b_eq, b_pol, b_np, padding=padding,
                                     dilation_rate=(dil, dil), flip_north_pole=flip, halo=halo,
                                     activation=('capped_leaky_relu', 0.1, 10.0) if c['act'] else None)
        return xi, wi, bi, y

    xo, wo, bo, yo = run(True)
    gy = torch.randn(yo.shape, generator=g)
    # keep away from the activation's kinks so that float32 rounding cannot flip a derivative
    yo.backward(gy.double())
    xc, wc, bc, yc = run(False)
    close(yc, yo)
    yc.backward(gy.cuda())
    close(xc.grad, xo.grad)
    for a, b in zip(wc, wo):
        close(a.grad, b.grad)
    if bs is not None:
        for a, b in zip(bc, bo):
            close(a.grad, b.grad)


def test_errors(cs):
    pkg, lib, F = cs
    with pytest.raises(ValueError):
        F.cube_sphere_pad(torch.zeros(1, 5, 4, 4, 2).cuda(), 1)              # face axis != 6
    with pytest.raises(ValueError):
        F.cube_sphere_pad(torch.zeros(1, 6, 4, 5, 2).cuda(), 1)              # H != W
    with pytest.raises(ValueError):
        layer = pkg.CubeSphereConv2D(4, 3, data_format='channels_last', in_channels=3).cuda()
        layer(torch.zeros(1, 6, 8, 8, 5).cuda())                              # channel mismatch
    with pytest.raises(lib.DlwpcsError):
        layer = pkg.CubeSphereConv2D(4, 9, data_format='channels_last', in_channels=3).cuda()
        layer(torch.zeros(1, 6, 8, 8, 3).cuda())                              # kernel larger than the face


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_resampling_kernels_vs_oracle(cs, dtype):
    """dlwpcs_pool2 / dlwpcs_up2cat_fwd / _bwd (AveragePooling3D((1,2,2)), UpSampling3D((1,2,2)) + concatenate,
    train_cs.py:197-198, 293) against the oracle's slicing restatement, forward and adjoint."""
    from dlwp_cs_b200 import functional as F
    g = torch.Generator().manual_seed(17)
    b, n, ca, cb = 2, 8, 16, 8
    x = torch.randn(b, 6, 2 * n, 2 * n, cb, generator=g).to(dtype)
    a = torch.randn(b, 6, n // 2, n // 2, ca, generator=g).to(dtype)
    bb = torch.randn(b, 6, n, n, cb, generator=g).to(dtype)
    tol = 1e-6 if dtype == torch.float32 else 2.0 ** -8

    def close(u, v):
        np.testing.assert_allclose(u.double().cpu().numpy(), v.double().numpy(), rtol=tol, atol=tol)
    # forward
    xc = x.cuda().requires_grad_(True)
    y = F.avg_pool_2x2(xc)
    close(y.detach(), O.avg_pool_2x2(x.double()))
    ac, bc = a.cuda().requires_grad_(True), bb.cuda().requires_grad_(True)
    t = F.upsample_concat(ac, bc)
    assert torch.equal(t.detach().cpu(), torch.cat([O.upsample_2x2(a), bb], dim=-1))
    # adjoints against autograd of the oracle
    gy = torch.randn(y.shape, generator=g).to(dtype)
    gt = torch.randn(t.shape, generator=g).to(dtype)
    y.backward(gy.cuda())
    t.backward(gt.cuda())
    xo = x.double().requires_grad_(True)
    O.avg_pool_2x2(xo).backward(gy.double())
    ao, bo = a.double().requires_grad_(True), bb.double().requires_grad_(True)
    torch.cat([O.upsample_2x2(ao), bo], dim=-1).backward(gt.double())
    close(xc.grad, xo.grad)
    np.testing.assert_allclose(ac.grad.double().cpu().numpy(), ao.grad.numpy(), rtol=tol, atol=4 * tol)
    assert torch.equal(bc.grad.cpu().double(), bo.grad)
