"""GPU parity of the bf16 tcgen05 path (cs_tc.cu) through the C ABI, against the float64 oracle on IDENTICAL inputs: x and
the weights are rounded to bfloat16 first, so what is measured is the kernel (fp32 accumulation in tensor memory), not
the input quantisation.  Tolerances (BASELINE.json north_star: 1e-3 for bf16):
  * float32-stored output:  rtol 1e-3 of the output's max magnitude (observed ~1e-6),
  * bfloat16-stored output: additionally one bf16 rounding of the result, 2^-8 relative per element.
Cases go from a single tensor-core instruction to the fully fused U-Net layers so that a failure localises."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import cs_oracle as O  # noqa: E402

RTOL = 1e-3


@pytest.fixture(scope='module')
def lib():
    from dlwp_cs_b200 import _lib
    _lib.load()
    return _lib


def bf(t):
    return t.to(torch.bfloat16)


def check(y, ref, stored_bf16=False):
    a = y.detach().double().cpu().numpy()
    e = ref.detach().double().cpu().numpy()
    assert a.shape == e.shape, (a.shape, e.shape)
    scale = max(float(np.abs(e).max()), 1e-30)
    if stored_bf16:
        np.testing.assert_allclose(a, e, rtol=2.0 ** -8, atol=RTOL * scale)
    else:
        np.testing.assert_allclose(a, e, rtol=RTOL, atol=RTOL * scale)
        assert float(np.abs(a - e).max()) <= 2e-5 * scale          # what fp32 accumulation actually delivers


def run_case(lib, n, cin, cout, k, halo, batch=2, act=False, same=False, dil=1, flip=True, indep=False, bias=True,
             out_dtype=torch.float32, seed=0):
    g = torch.Generator().manual_seed(seed + n * 1000 + cin * 10 + cout)
    x = bf(torch.randn(batch, 6, n, n, cin, generator=g))
    nw = 3 if indep else 2
    ws = [bf(torch.randn(k, k, cin, cout, generator=g) * 0.2).float() for _ in range(nw)]
    bs = [torch.randn(cout, generator=g) * 0.2 for _ in range(nw)] if bias else None
    w_np = ws[2] if indep else None
    b_eq, b_pol, b_np = (None, None, None) if bs is None else (bs[0], bs[1], bs[2] if indep else None)
    dd = lambda t: None if t is None else t.double()
    ref = O.cube_sphere_conv2d(O.cube_sphere_pad(x.double(), halo), dd(ws[0]), dd(ws[1]), dd(w_np), dd(b_eq), dd(b_pol),
                               dd(b_np), padding='same' if same else 'valid', dilation=dil, flip_north_pole=flip)
    if act:
        ref = O.capped_leaky_relu(ref)
    cu = lambda t: None if t is None else t.cuda()
    d = lib.make_desc(batch, n, cin, cout, (k, k), (1, 1), (dil, dil), halo, same, flip, indep, bias,
                      lib.ACT_CAPPED_LEAKY_RELU if act else lib.ACT_NONE, 0.1, 10.0, lib.BF16,
                      lib.F32 if out_dtype == torch.float32 else lib.BF16)
    packed = lib.pack_weights(d, cu(ws[0]), cu(ws[1]), cu(w_np), cu(b_eq), cu(b_pol), cu(b_np))
    y = lib.conv2d_fwd(d, x.cuda(), None, packed)
    torch.cuda.synchronize()
    assert y.dtype == out_dtype
    check(y, ref, stored_bf16=out_dtype == torch.bfloat16)


# name, kwargs -- ordered by how much of the kernel they exercise
TC_CASES = [
    ('gemm_k16_n16', dict(n=16, cin=16, cout=16, k=1, halo=0, bias=False)),              # one MMA per m-block
    ('gemm_k64_n32', dict(n=16, cin=64, cout=32, k=1, halo=0)),                           # 4 K-steps, bias
    ('gemm_k128_n64_2chunks', dict(n=16, cin=128, cout=64, k=1, halo=0)),                 # two K chunks
    ('gemm_ragged_m', dict(n=10, cin=32, cout=16, k=1, halo=0)),                          # 100 rows: partial m-block
    ('conv3_valid_nohalo', dict(n=12, cin=16, cout=16, k=3, halo=0)),                     # taps = row offsets
    ('conv3_halo', dict(n=12, cin=32, cout=32, k=3, halo=1, act=True)),                   # fused halo gather
    ('conv3_halo_c48', dict(n=48, cin=32, cout=32, k=3, halo=1, act=True)),
    ('conv3_halo_64_64', dict(n=24, cin=64, cout=64, k=3, halo=1, act=True)),
    ('conv3_halo_128_64', dict(n=24, cin=128, cout=64, k=3, halo=1, act=True)),           # streamed weight ring
    ('conv3_halo_64_128', dict(n=12, cin=64, cout=128, k=3, halo=1, act=True)),
    ('conv3_scalar_18_32', dict(n=48, cin=18, cout=32, k=3, halo=1, act=True, batch=1)),  # scalar gather path
    ('conv1_head_32_14', dict(n=48, cin=32, cout=14, k=1, halo=0, batch=1)),              # masked output channels
    ('conv3_same', dict(n=9, cin=16, cout=24, k=3, halo=0, same=True)),
    ('conv3_dil2', dict(n=10, cin=8, cout=16, k=3, halo=0, dil=2)),
    ('conv5_halo2', dict(n=8, cin=24, cout=40, k=5, halo=2, act=True)),
    ('conv3_noflip', dict(n=6, cin=16, cout=16, k=3, halo=1, flip=False)),
    ('conv3_indep_np', dict(n=6, cin=16, cout=16, k=3, halo=1, indep=True)),
    ('conv3_bf16_out', dict(n=24, cin=64, cout=64, k=3, halo=1, act=True, out_dtype=torch.bfloat16)),
    ('conv3_96', dict(n=96, cin=32, cout=32, k=3, halo=1, act=True, batch=1)),
    ('conv3_k48', dict(n=8, cin=48, cout=16, k=3, halo=1)),                               # K chunk of 48
    ('conv3_cin256', dict(n=8, cin=256, cout=32, k=3, halo=1)),                           # 4 chunks
]


@pytest.mark.parametrize('name,kw', TC_CASES, ids=[c[0] for c in TC_CASES])
def test_tc_conv_vs_oracle(lib, name, kw):
    run_case(lib, **kw)


def test_tc_fused_sources(lib):
    """upsample (+) skip concat and 2x2 average pooling folded into the gather (Azure/train_cs.py:282-299)."""
    g = torch.Generator().manual_seed(11)
    n = 12
    big = bf(torch.randn(2, 6, 2 * n, 2 * n, 16, generator=g))
    small = bf(torch.randn(2, 6, n // 2, n // 2, 16, generator=g))
    same = bf(torch.randn(2, 6, n, n, 8, generator=g))
    w = [bf(torch.randn(3, 3, 24, 16, generator=g) * 0.1).float() for _ in range(2)]
    b = [torch.randn(16, generator=g) * 0.1 for _ in range(2)]
    # upsample (+) skip
    d = lib.make_desc(2, n, 24, 16, halo=1, c0=16, mode0=lib.SRC_UP2, c1=8, mode1=lib.SRC_SAME, x_dtype=lib.BF16)
    packed = lib.pack_weights(d, w[0].cuda(), w[1].cuda(), None, b[0].cuda(), b[1].cuda(), None)
    y = lib.conv2d_fwd(d, small.cuda(), same.cuda(), packed)
    xin = torch.cat([O.upsample_2x2(small), same], -1).double()
    ref = O.cube_sphere_conv2d(O.cube_sphere_pad(xin, 1), w[0].double(), w[1].double(), None, b[0].double(),
                               b[1].double(), None)
    check(y, ref)
    # average pool: the pooled value is rounded to bf16 before the MMA (as if the pooled tensor were stored in bf16)
    w16 = [t[:, :, :16].contiguous() for t in w]
    d = lib.make_desc(2, n, 16, 16, halo=1, mode0=lib.SRC_POOL2, x_dtype=lib.BF16)
    packed = lib.pack_weights(d, w16[0].cuda(), w16[1].cuda(), None, b[0].cuda(), b[1].cuda(), None)
    y = lib.conv2d_fwd(d, big.cuda(), None, packed)
    pooled = bf(O.avg_pool_2x2(big.float())).double()
    ref = O.cube_sphere_conv2d(O.cube_sphere_pad(pooled, 1), w16[0].double(), w16[1].double(), None, b[0].double(),
                               b[1].double(), None)
    check(y, ref)


def test_tc_rollout_vs_oracle(lib):
    """Two steps of the bf16 device-resident U-Net rollout against the float64 oracle (bf16-rounded weights/inputs):
    what is left is the bf16 storage of every activation, ~2^-9 per layer."""
    from dlwp_cs_b200.unet import CubeSphereUNet2, RolloutEngine
    n, b, cp, cf, steps, base = 16, 2, 6, 2, 2, 16
    params = {k: bf(v).float() for k, v in O.make_unet2_params(cp + cf, cp, base=base, seed=3).items()}
    model = CubeSphereUNet2(cp + cf, cp, base=base).cuda()
    model.load_oracle_params(params)
    g = torch.Generator().manual_seed(5)
    state = bf(torch.randn(b, 6, n, n, cp, generator=g))
    forcing = bf(torch.rand(b, 6, n, n, cf, generator=g))
    ref = O.rollout({k: v.double() for k, v in params.items()}, state.double(), forcing.double(), steps)
    eng = RolloutEngine(model, b, n, steps, forcing_channels=cf, dtype=torch.bfloat16)
    out = eng.run(state.cuda(), forcing.cuda())
    torch.cuda.synchronize()
    err = (out.double().cpu() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 3e-2, err
    # the host-streaming variant (forecast chunks copied to pinned memory while later steps run) delivers the same bits
    for chunk in (1, 2):
        host = eng.run_to_host(state.pin_memory(), forcing.pin_memory(), chunk=chunk)
        torch.cuda.synchronize()
        assert not host.is_cuda and host.is_pinned()
        assert torch.equal(host, out.cpu())


BWD_TC_CASES = [
    dict(n=12, cin=32, cout=32, k=3, halo=1, act=True),
    dict(n=24, cin=64, cout=32, k=3, halo=1, act=True),
    dict(n=8, cin=16, cout=24, k=3, halo=1, act=False, flip=False),
    dict(n=8, cin=16, cout=16, k=3, halo=1, act=True, indep=True),
    dict(n=16, cin=32, cout=16, k=1, halo=0, act=False),
    dict(n=9, cin=16, cout=8, k=3, halo=0, act=True, same=True),
    dict(n=48, cin=24, cout=32, k=3, halo=1, act=True, batch=1),
    # shapes that exercise the tcgen05 wgrad's job splits: two 64-channel input blocks, two output-channel groups,
    # channel counts that are not multiples of 8 (element-wise loads), the 1x1 head, a 5x5 window, 3 images
    dict(n=24, cin=128, cout=64, k=3, halo=1, act=True),
    dict(n=12, cin=64, cout=128, k=3, halo=1, act=True),
    dict(n=24, cin=64, cout=64, k=3, halo=1, act=True, batch=3),
    dict(n=16, cin=18, cout=32, k=3, halo=1, act=True),
    dict(n=16, cin=32, cout=14, k=1, halo=0, act=False),
    dict(n=12, cin=16, cout=16, k=5, halo=2, act=True),
    dict(n=48, cin=32, cout=32, k=3, halo=1, act=True, batch=3),
]


@pytest.mark.parametrize('c', BWD_TC_CASES, ids=[str(i) for i in range(len(BWD_TC_CASES))])
def test_tc_backward_vs_oracle_autograd(lib, c):
    """bf16 backward through the differentiable operator: dgrad on the tcgen05 kernel (rotated, channel-swapped weights,
    activation-derivative mask, halo scatter-add), wgrad accumulated in float32 from the bf16 tensors."""
    from dlwp_cs_b200 import functional as F
    n, cin, cout, k, halo = c['n'], c['cin'], c['cout'], c['k'], c['halo']
    bsz = c.get('batch', 2)
    g = torch.Generator().manual_seed(n * 100 + cin * 10 + cout)
    x = bf(torch.randn(bsz, 6, n, n, cin, generator=g))
    nw = 3 if c.get('indep') else 2
    # small weights keep every output below the ReLU cap of 10: the derivative is taken from the bf16-STORED output, and a
    # value within one bf16 step of the cap would round onto it and flip its mask (the kink at 0 is sign-preserving)
    ws = [bf(torch.randn(k, k, cin, cout, generator=g) * 0.04).float() for _ in range(nw)]
    bs = [torch.randn(cout, generator=g) * 0.2 for _ in range(nw)]
    padding, flip = 'same' if c.get('same') else 'valid', c.get('flip', True)

    def run(oracle):
        cast = (lambda t: t.double().requires_grad_(True)) if oracle else (lambda t: t.cuda().requires_grad_(True))
        xi, wi, bi = cast(x), [cast(w) for w in ws], [cast(b) for b in bs]
        w_np, b_np = (wi[2], bi[2]) if nw == 3 else (None, None)
        if oracle:
            y = O.cube_sphere_conv2d(O.cube_sphere_pad(xi, halo), wi[0], wi[1], w_np, bi[0], bi[1], b_np, padding=padding,
                                     flip_north_pole=flip)
            if c['act']:
                y = O.capped_leaky_relu(y)
        else:
            y = F.cube_sphere_conv2d(xi, wi[0], wi[1], w_np, bi[0], bi[1], b_np, padding=padding, flip_north_pole=flip,
                                     halo=halo, activation=('capped_leaky_relu', 0.1, 10.0) if c['act'] else None)
        return xi, wi, bi, y

    xo, wo, bo, yo = run(True)
    gy = bf(torch.randn(yo.shape, generator=g))
    yo.backward(gy.double())
    xc, wc, bc, yc = run(False)
    assert yc.dtype == torch.bfloat16
    check(yc, yo, stored_bf16=True)
    yc.backward(gy.cuda())

    def close(a, e, rtol):
        a, e = a.double().cpu().numpy(), e.double().numpy()
        np.testing.assert_allclose(a, e, rtol=rtol, atol=rtol * max(float(np.abs(e).max()), 1e-30))
    # dx: bf16 storage of dx (2^-8) on top of the bf16 rounding of the masked dy and of the padded-gradient workspace
    close(xc.grad, xo.grad, 1.5e-2)
    # wgrad: float32 accumulation of bf16-exact inputs; the bf16-rounded forward output only moves the activation mask
    # where y is within rounding of a kink, which the random data avoids almost surely
    for a, b in zip(wc, wo):
        close(a.grad, b.grad, 2e-3)
    for a, b in zip(bc, bo):
        close(a.grad, b.grad, 2e-3)


@pytest.mark.parametrize('n,cin,cout', [(12, 32, 32), (24, 64, 64)])
def test_tc_in_kernel_mask_equals_separate_act_bwd(lib, n, cin, cout):
    """C ABI: dgrad / wgrad with the activation derivative applied inside the kernels (register path) against the same
    kernels fed with dy * act'(y) from dlwpcs_act_bwd (asynchronous-copy path, bias sums read back from the dy tile in shared memory)."""
    g = torch.Generator().manual_seed(n + cin)
    b = 2
    x = bf(torch.randn(b, 6, n, n, cin, generator=g)).cuda()
    y = bf(torch.randn(b, 6, n, n, cout, generator=g) * 6).cuda()          # spans both kinks (0 and the cap 10)
    dy = bf(torch.randn(b, 6, n, n, cout, generator=g)).cuda()
    ws = [(torch.randn(3, 3, cin, cout, generator=g) * 0.05).cuda() for _ in range(2)]
    d = lib.make_desc(b, n, cin, cout, (3, 3), (1, 1), (1, 1), 1, False, True, False, True, lib.ACT_CAPPED_LEAKY_RELU, 0.1,
                      10.0, lib.BF16, lib.BF16)
    d0 = lib.copy_desc(d, act=lib.ACT_NONE)
    dym = lib.act_bwd(dy, y, lib.ACT_CAPPED_LEAKY_RELU, 0.1, 10.0)
    packed_t = lib.pack_weights(d, ws[0], ws[1], None, transposed=True)
    assert torch.equal(lib.conv2d_dgrad(d, dy, y, packed_t), lib.conv2d_dgrad(d0, dym, None, packed_t))
    a = lib.conv2d_wgrad(d, x, dy, y)
    e = lib.conv2d_wgrad(d0, x, dym, None)
    for u, v in zip(a, e):
        if u is not None:
            # weights: identical bf16 operands -> identical sums; bias: the in-kernel path sums the unrounded products
            np.testing.assert_allclose(u.cpu().numpy(), v.cpu().numpy(), rtol=2e-3, atol=2e-3 * float(v.abs().max()))


def test_tc_rollout_c96_12var_vs_oracle(lib):
    """BASELINE.json configs[4] shape: C96 faces, 12 variables x 2 time steps (+2 insolation +2 constants), base 32; one
    bf16 model step against the float64 oracle on bf16-rounded weights / inputs."""
    from dlwp_cs_b200.unet import CubeSphereUNet2, RolloutEngine
    n, b, cp, cf, steps, base = 96, 1, 24, 4, 1, 32
    params = {k: bf(v).float() for k, v in O.make_unet2_params(cp + cf, cp, base=base, seed=8).items()}
    model = CubeSphereUNet2(cp + cf, cp, base=base).cuda()
    model.load_oracle_params(params)
    g = torch.Generator().manual_seed(6)
    state = bf(torch.randn(b, 6, n, n, cp, generator=g))
    forcing = bf(torch.rand(b, 6, n, n, cf, generator=g))
    ref = O.rollout({k: v.double() for k, v in params.items()}, state.double(), forcing.double(), steps)
    eng = RolloutEngine(model, b, n, steps, forcing_channels=cf, dtype=torch.bfloat16)
    out = eng.run(state.cuda(), forcing.cuda())
    torch.cuda.synchronize()
    err = (out.double().cpu() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 3e-2, err


def test_tc_conv_ragged_tiles_many_per_cta(lib):
    """A face whose m-blocks do not divide evenly into tiles (19 = 4+4+4+4+3 at C48) with enough images that every CTA
    walks many tiles: exercises the per-m-block accumulator barriers across short last tiles (a phase slip deadlocks or
    reads a stale accumulator)."""
    y_ref = None
    for batch in (40,):
        g = torch.Generator().manual_seed(3)
        x = bf(torch.randn(batch, 6, 48, 48, 32, generator=g))
        ws = [bf(torch.randn(3, 3, 32, 32, generator=g) * 0.1).float() for _ in range(2)]
        bs = [torch.randn(32, generator=g) * 0.1 for _ in range(2)]
        d = lib.make_desc(batch, 48, 32, 32, (3, 3), (1, 1), (1, 1), 1, False, True, False, True, lib.ACT_CAPPED_LEAKY_RELU,
                          0.1, 10.0, lib.BF16, lib.BF16)
        packed = lib.pack_weights(d, ws[0].cuda(), ws[1].cuda(), None, bs[0].cuda(), bs[1].cuda(), None)
        y = lib.conv2d_fwd(d, x.cuda(), None, packed)
        torch.cuda.synchronize()
        sel = [0, 17, batch - 1]                         # oracle on three images only (float64 CPU)
        ref = O.capped_leaky_relu(O.cube_sphere_conv2d(O.cube_sphere_pad(x[sel].double(), 1), ws[0].double(), ws[1].double(),
                                                       None, bs[0].double(), bs[1].double(), None))
        check(y[sel], ref, stored_bf16=True)


def _exact_inputs(b, n, cin, cout, seed):
    """Integer-valued activations and weights in multiples of 1/8: every product and every partial sum of the convolution
    is exactly representable in float32, so the tensor-core result must equal the exact sum whatever the summation order."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randint(-4, 5, (b, 6, n, n, cin), generator=g).to(torch.bfloat16)
    ws = [(torch.randint(-8, 9, (3, 3, cin, cout), generator=g).float() / 8.0) for _ in range(2)]
    return x, ws


def test_tc_full_size_properties_c48_batch64(lib):
    """BASELINE.json full size (C48, batch 64, 32 -> 32 channels) through size-independent properties:
    exactness on exactly-representable data (checked against the float64 oracle on three images), linearity in the input,
    and agreement of the fused halo gather with the standalone padding kernel followed by an un-padded convolution."""
    b, n, cin, cout = 64, 48, 32, 32
    x1, ws = _exact_inputs(b, n, cin, cout, 11)
    x2, _ = _exact_inputs(b, n, cin, cout, 12)
    d = lib.make_desc(b, n, cin, cout, (3, 3), (1, 1), (1, 1), 1, False, True, False, False, lib.ACT_NONE, 0.1, 10.0, lib.BF16,
                      lib.F32)
    packed = lib.pack_weights(d, ws[0].cuda(), ws[1].cuda(), None)
    y1 = lib.conv2d_fwd(d, x1.cuda(), None, packed)
    y2 = lib.conv2d_fwd(d, x2.cuda(), None, packed)
    y12 = lib.conv2d_fwd(d, (x1.float() + x2.float()).to(torch.bfloat16).cuda(), None, packed)     # |x1 + x2| <= 8: exact in bf16
    assert torch.equal(y12, y1 + y2)
    sel = [0, 31, 63]
    ref = O.cube_sphere_conv2d(O.cube_sphere_pad(x1[sel].double(), 1), ws[0].double(), ws[1].double(), None)
    assert torch.equal(y1[sel].double().cpu(), ref)
    # fused halo gather == standalone pad kernel + convolution without halo (bf16 padded tensor, same arithmetic)
    xp = lib.pad_fwd(x1.cuda(), 1)
    d0 = lib.make_desc(b, n + 2, cin, cout, (3, 3), (1, 1), (1, 1), 0, False, True, False, False, lib.ACT_NONE, 0.1, 10.0,
                       lib.BF16, lib.F32)
    y1b = lib.conv2d_fwd(d0, xp, None, lib.pack_weights(d0, ws[0].cuda(), ws[1].cuda(), None))
    assert torch.equal(y1b, y1)


def test_tc_backward_full_size_exactness_c48_batch32(lib):
    """Training size (C48, batch 32): dgrad and wgrad on exactly-representable data equal the float64 oracle's autograd
    -- wgrad sums 32 x 4 x 2304 products per weight in tensor memory and must reproduce the integer-valued exact sum; dgrad
    is checked on three images (its bf16 output stores the exact value as long as it has at most 8 significant bits, so the
    weights are restricted to {-1, 0, 1} there)."""
    b, n, cin, cout = 32, 48, 32, 32
    g = torch.Generator().manual_seed(21)
    x = torch.randint(-2, 3, (b, 6, n, n, cin), generator=g).to(torch.bfloat16)
    dy = torch.randint(-1, 2, (b, 6, n, n, cout), generator=g).to(torch.bfloat16)
    ws = [torch.randint(-1, 2, (3, 3, cin, cout), generator=g).float() for _ in range(2)]
    d = lib.make_desc(b, n, cin, cout, (3, 3), (1, 1), (1, 1), 1, False, True, False, True, lib.ACT_NONE, 0.1, 10.0, lib.BF16,
                      lib.BF16)
    dw_eq, dw_pol, _, db_eq, db_pol, _ = lib.conv2d_wgrad(d, x.cuda(), dy.cuda(), None)
    xo = x.double().requires_grad_(True)
    wo = [w.double().requires_grad_(True) for w in ws]
    bo = [torch.zeros(cout, dtype=torch.float64, requires_grad=True) for _ in range(2)]
    yo = O.cube_sphere_conv2d(O.cube_sphere_pad(xo, 1), wo[0], wo[1], None, bo[0], bo[1], None)
    yo.backward(dy.double())
    assert torch.equal(dw_eq.double().cpu(), wo[0].grad) and torch.equal(dw_pol.double().cpu(), wo[1].grad)
    assert torch.equal(db_eq.double().cpu(), bo[0].grad) and torch.equal(db_pol.double().cpu(), bo[1].grad)
    packed_t = lib.pack_weights(d, ws[0].cuda(), ws[1].cuda(), None, transposed=True)
    dx = lib.conv2d_dgrad(d, dy.cuda(), None, packed_t)
    sel = [0, 13, 31]
    # |dx| <= 9 * 32 * (corner multiplicity 5) ... the bf16 store keeps 8 significant bits: compare with one rounding
    ref = xo.grad[sel]
    got = dx[sel].double().cpu()
    assert torch.equal(got, ref.to(torch.bfloat16).double())


def test_chained_launches_equal_stream_ordered_launches():
    """dlwpcs_conv2d_fwd_chained (per-sample completion counters + dynamic tile scheduler, layer k+1 overlapping the tail of
    layer k) must give bit-identical forecasts to plain stream-ordered launches, with and without a CUDA graph, over
    several steps and an ensemble large enough that consecutive layers really overlap."""
    from dlwp_cs_b200.unet import CubeSphereUNet2, RolloutEngine
    torch.manual_seed(3)
    n, b, cp, cf, steps = 48, 24, 14, 4, 6
    model = CubeSphereUNet2(cp + cf, cp, base=32).cuda()
    g = torch.Generator().manual_seed(0)
    state = torch.randn(b, 6, n, n, cp, generator=g).cuda()
    forcing = torch.rand(b, 6, n, n, cf, generator=g).cuda()
    ref = plain = None
    for chain, graph in ((False, False), (True, False), (True, True), (True, True)):
        eng = RolloutEngine(model, b, n, steps, forcing_channels=cf, dtype=torch.bfloat16, use_graph=graph, chain=chain)
        assert eng.chain == chain
        for _ in range(3):                              # replays reuse the counters: they are re-zeroed inside every run
            out = eng.run(state, forcing).clone()
            torch.cuda.synchronize()
            assert not eng.chain_error
            if not chain:
                plain = out
                continue
            if ref is None:
                ref = out
            assert torch.equal(out, ref), (chain, graph)
    # the stream-ordered engine runs its 3x3 layers with <= 64 input channels on the row-streamed kernel (another float32
    # summation order): equal up to the bf16 rounding of a few layer outputs
    scale = float(plain.float().abs().max())
    assert float((plain.float() - ref.float()).abs().max()) <= 2.0 ** -6 * scale
    # host-streamed variant (events between the chunks) on the chained engine
    host = eng.run_to_host(state.cpu().bfloat16().pin_memory(), forcing.cpu().bfloat16().pin_memory(), chunk=2)
    torch.cuda.synchronize()
    assert torch.equal(host, ref.cpu()) and not eng.chain_error


# ---- float32-accurate tensor-core path (split bf16 x3: dlwpcs_split3 + [w_hi ; w_hi ; w_lo]) ---------------------------
@pytest.mark.parametrize('n,cin,cout,k,halo,act,flip,indep', [
    (48, 3, 3, 3, 1, False, True, False),          # BASELINE configs[0]
    (12, 32, 32, 3, 1, True, True, False),
    (24, 64, 32, 3, 1, True, True, False),
    (12, 128, 64, 3, 1, True, True, False),
    (8, 18, 32, 3, 1, True, False, True),
    (16, 32, 14, 1, 0, False, True, False),
    (12, 16, 24, 5, 2, False, True, False),
])
def test_tc32_conv_float32_accuracy(n, cin, cout, k, halo, act, flip, indep):
    """float32 inputs, float32 weights against the float64 oracle -- on the tensor cores.  Each operand carries 16 mantissa
    bits (hi + lo bf16) and the lo*lo term is dropped: ~2^-17 relative per product, i.e. <= 1e-5 of the output scale up to
    K ~ 600 and ~1.6e-5 at the widest layer of the U-Net (K = 9 x 128), where the test allows 2.5e-5."""
    from dlwp_cs_b200 import functional as F_cs
    g = torch.Generator().manual_seed(n * 100 + cin)
    x = torch.randn(2, 6, n, n, cin, generator=g)
    nw = 3 if indep else 2
    ws = [torch.randn(k, k, cin, cout, generator=g) * 0.2 for _ in range(nw)]
    bs = [torch.randn(cout, generator=g) * 0.2 for _ in range(nw)]
    dd = lambda t: None if t is None else t.double()
    w_np, b_np = (ws[2], bs[2]) if indep else (None, None)
    ref = O.cube_sphere_conv2d(O.cube_sphere_pad(x.double(), halo), dd(ws[0]), dd(ws[1]), dd(w_np), dd(bs[0]), dd(bs[1]),
                               dd(b_np), flip_north_pole=flip)
    if act:
        ref = O.capped_leaky_relu(ref)
    cu = lambda t: None if t is None else t.cuda()
    y = F_cs.cube_sphere_conv2d_tc32(x.cuda(), cu(ws[0]), cu(ws[1]), cu(w_np), cu(bs[0]), cu(bs[1]), cu(b_np),
                                     flip_north_pole=flip, halo=halo,
                                     activation=('capped_leaky_relu', 0.1, 10.0) if act else None)
    assert y.dtype == torch.float32 and tuple(y.shape) == tuple(ref.shape)
    a, e = y.double().cpu().numpy(), ref.numpy()
    scale = float(np.abs(e).max())
    tol = 1e-5 if k * k * cin <= 600 else 2.5e-5
    np.testing.assert_allclose(a, e, rtol=tol, atol=tol * scale)


def test_tc32_cfg1_vs_reference_golden(golden_dir):
    """BASELINE configs[0] against the outputs of the reference's OWN call() methods (tests/golden/padconv_cfg1.npz)."""
    import os
    from dlwp_cs_b200 import functional as F_cs
    g = np.load(os.path.join(golden_dir, 'padconv_cfg1.npz'))
    t = lambda k: torch.from_numpy(g[k]).float().cuda()
    y = F_cs.cube_sphere_conv2d_tc32(t('x'), t('w_eq'), t('w_pol'), None, t('b_eq'), t('b_pol'), None, halo=1)
    ref = g['y']
    np.testing.assert_allclose(y.cpu().numpy(), ref, rtol=1e-5, atol=1e-5 * float(np.abs(ref).max()))


def test_tc32_rollout_100_steps_c48():
    """The float32-accurate tensor-core engine over the 100-step C48 rollout: every step within 1e-4 of the float64 oracle
    (the CUDA-core float32 engine holds the same bound in tests/test_gpu_models.py), no growth along the rollout."""
    from dlwp_cs_b200.unet import CubeSphereUNet2, RolloutEngine
    n, cp, cf, batch, steps = 48, 14, 4, 2, 100
    params = O.make_unet2_params(cp + cf, cp, base=32, seed=1)
    model = CubeSphereUNet2(cp + cf, cp, base=32).cuda()
    model.load_oracle_params(params)
    g = torch.Generator().manual_seed(0)
    state = torch.randn(batch, 6, n, n, cp, generator=g).clamp_(-5, 5)
    forcing = torch.rand(batch, 6, n, n, cf, generator=g)
    with torch.no_grad():
        ref = O.rollout({k: v.double() for k, v in params.items()}, state.double(), forcing.double(), steps, exact=False)
    eng = RolloutEngine(model, batch, n, steps, forcing_channels=cf, dtype=torch.float32, tensor_cores=True)
    assert eng.tc32 and eng.launches_per_step == 22
    got = eng.run(state.cuda(), forcing.cuda()).double().cpu()
    torch.cuda.synchronize()
    err = ((got - ref).abs().flatten(1).max(dim=1).values / ref.abs().flatten(1).max(dim=1).values).numpy()
    print('tc32 rel err at steps 1,2,5,10,50,100:', [float('%.2e' % err[i]) for i in (0, 1, 4, 9, 49, 99)])
    assert err.max() <= 1e-4, err.max()
    assert err[50:].max() <= 2.0 * err[:5].max() + 1e-5


# ---- row-streamed kernel (cs_tc_rs.cu): three kernel rows stacked along N, shift-add by overlapping accumulator ranges -----
def _rs_case(lib, batch, n, cin, cout, srcs=1, indep=False, bias=True, act=True, seed=0, exact=True, flip=True):
    """bf16 in / bf16 out, 3x3, halo 1 -> served by conv_rs_kernel.  exact: activations and weights in {-1, 0, 1} (sums
    below 256 in magnitude are exact in float32 AND in the bf16 output), compared with == against the float64 oracle.
    srcs: 1 plain, 2 = [nearest-upsampled a | b] (Azure/train_cs.py:293, 299), 'pool' = 2x2 mean of a finer tensor."""
    g = torch.Generator().manual_seed(seed + 7 * n + cin)
    if exact:
        mk = lambda *shape: torch.randint(-1, 2, shape, generator=g).float()
    else:
        mk = lambda *shape: torch.randn(*shape, generator=g)
    nw = 3 if indep else 2
    ws = [mk(3, 3, cin, cout) if exact else bf(mk(3, 3, cin, cout) * 0.1).float() for _ in range(nw)]
    bs = [mk(cout) if exact else mk(cout) * 0.1 for _ in range(nw)] if bias else [None] * nw
    if srcs == 1:
        x = bf(mk(batch, 6, n, n, cin))
        xin = x.double()
        x0, x1, c0, c1, m0 = x, None, cin, 0, lib.SRC_SAME
    elif srcs == 'pool':            # exact data: multiples of 4 so that the 2x2 mean stays an integer
        x = bf(mk(batch, 6, 2 * n, 2 * n, cin) * (4.0 if exact else 1.0))
        xin = bf(O.avg_pool_2x2(x.double()).float()).double()      # the mean is rounded to bf16 once, like a stored tensor
        x0, x1, c0, c1, m0 = x, None, cin, 0, lib.SRC_POOL2
    else:
        ca = cin // 2
        a = bf(mk(batch, 6, n // 2, n // 2, ca))
        b2 = bf(mk(batch, 6, n, n, cin - ca))
        xin = torch.cat([O.upsample_2x2(a.double()), b2.double()], dim=-1)
        x0, x1, c0, c1, m0 = a, b2, ca, cin - ca, lib.SRC_UP2
    dd = lambda t: None if t is None else t.double()
    ref = O.cube_sphere_conv2d(O.cube_sphere_pad(xin, 1), dd(ws[0]), dd(ws[1]), dd(ws[2]) if indep else None, dd(bs[0]),
                               dd(bs[1]), dd(bs[2]) if indep else None, flip_north_pole=flip)
    if act:
        ref = O.capped_leaky_relu(ref, 0.5 if exact else 0.1, 1e9 if exact else 10.0)
    cu = lambda t: None if t is None else t.cuda()
    d = lib.make_desc(batch, n, cin, cout, (3, 3), (1, 1), (1, 1), 1, False, flip, indep, bias,
                      lib.ACT_CAPPED_LEAKY_RELU if act else lib.ACT_NONE, 0.5 if exact else 0.1, 1e9 if exact else 10.0,
                      lib.BF16, lib.BF16, c0, m0, c1, lib.SRC_SAME)
    packed = lib.pack_weights(d, cu(ws[0]), cu(ws[1]), cu(ws[2]) if indep else None, cu(bs[0]), cu(bs[1]),
                              cu(bs[2]) if indep else None)
    y = lib.conv2d_fwd(d, cu(x0), cu(x1), packed)
    torch.cuda.synchronize()
    return y, ref, d, (x0, x1, ws, bs)


@pytest.mark.parametrize('batch,n,cin,cout,srcs,indep,bias,act', [
    (2, 8, 16, 32, 1, False, True, True),          # one strip per face group, 16 input channels (32-byte rows)
    (2, 12, 24, 32, 1, False, True, False),        # 24 -> 32 padded input channels
    (3, 24, 24, 32, 1, True, True, True),          # independent north pole; strips that cut images
    (5, 48, 24, 32, 1, False, False, True),        # C48, no bias, 8 strips + ragged pole strips
    (4, 48, 16, 24, 1, False, True, True),         # 24 output channels (48-byte rows)
    (3, 24, 24, 32, 2, False, True, True),         # up-sampled + concatenated sources
    (3, 24, 32, 64, 1, False, True, True),         # 64 output channels: N = 192, 8 slots
    (2, 24, 64, 64, 1, False, True, True),         # 128-byte rows
    (3, 12, 32, 64, 'pool', False, True, True),    # 2x2 mean in the load stage
    (2, 24, 32, 48, 1, False, True, True),         # 48 output channels: 10 slots (ring length not a power of two)
])
def test_rs_kernel_exact_on_integer_data(lib, batch, n, cin, cout, srcs, indep, bias, act):
    y, ref, _, _ = _rs_case(lib, batch, n, cin, cout, srcs, indep, bias, act)
    assert float(ref.abs().max()) <= 256.0
    assert torch.equal(y.double().cpu(), ref)


@pytest.mark.parametrize('batch,n,cin,cout,srcs', [(40, 48, 32, 32, 1), (16, 24, 64, 32, 1), (12, 48, 64, 32, 2),
                                                    (3, 96, 32, 32, 1), (20, 24, 64, 64, 1), (9, 24, 32, 64, 'pool')])
def test_rs_kernel_vs_oracle_and_classic_kernel(lib, batch, n, cin, cout, srcs):
    """Random data at the U-Net's layer shapes (several units per CTA, both weight-group changes, ring wrap-around): within
    one bf16 rounding of the float64 oracle, and within two of the classic kernel (run through the chained entry point,
    which always takes the classic path)."""
    y, ref, d, (x0, x1, ws, bs) = _rs_case(lib, batch, n, cin, cout, srcs, exact=False, seed=5)
    sel = [0, batch // 2, batch - 1]
    check(y[sel], ref[sel], stored_bf16=True)
    cu = lambda t: None if t is None else t.cuda()
    packed_c = lib.pack_weights(d, cu(ws[0]), cu(ws[1]), None, cu(bs[0]), cu(bs[1]), None, transposed=2)
    yc = torch.empty_like(y)
    ctr = torch.zeros(1, dtype=torch.int32, device='cuda')
    lib.conv2d_fwd_chained(d, cu(x0), cu(x1), packed_c, yc, None, None, None, ctr)
    torch.cuda.synchronize()
    diff = (y.float() - yc.float()).abs()
    assert float((diff / (yc.float().abs() + 1e-3)).max()) <= 2.0 ** -6


# ---- fused 1x1 head (dlwpcs_conv2d_fwd_head): the output layer inside the epilogue of the 3x3 layer in front of it -----------
@pytest.mark.parametrize('batch,n,cin,cmid,cout2,srcs,act2', [
    (3, 24, 32, 32, 16, 1, False),        # the U-Net's tail: 32 -> 32 (ReLU) -> 16 linear
    (5, 48, 32, 32, 16, 1, False),        # C48, several units per CTA and both weight-group changes
    (2, 24, 64, 32, 24, 2, True),         # up-sampled + concatenated input, 24 head channels (32 padded), activated head
    (2, 12, 32, 64, 16, 1, False),        # 64 channels in between (128-byte A rows, K = 64)
    (3, 24, 16, 24, 8, 1, False),         # 24 channels in between (padded to 32 with exact zeros), 8 head channels
    (2, 8, 8, 16, 8, 1, False),           # 16 channels in between: the head's K (16) is half the padded A tile
])
def test_fused_head_equals_two_launches(lib, batch, n, cin, cmid, cout2, srcs, act2):
    """Same arithmetic as the two separate launches: one bf16 rounding in between, float32 accumulation of the 1x1 products
    in the same order -> bit-identical."""
    g = torch.Generator().manual_seed(11 * n + cin + cout2)
    mk = lambda *shape: torch.randn(*shape, generator=g)
    if srcs == 1:
        x0, x1, c0, c1, m0 = bf(mk(batch, 6, n, n, cin)).cuda(), None, cin, 0, lib.SRC_SAME
    else:
        ca = cin // 2
        x0, x1 = bf(mk(batch, 6, n // 2, n // 2, ca)).cuda(), bf(mk(batch, 6, n, n, cin - ca)).cuda()
        c0, c1, m0 = ca, cin - ca, lib.SRC_UP2
    w = [bf(mk(3, 3, cin, cmid) * 0.1).float().cuda() for _ in range(2)]
    b_ = [(mk(cmid) * 0.1).cuda() for _ in range(2)]
    wh = [bf(mk(1, 1, cmid, cout2) * 0.2).float().cuda() for _ in range(2)]
    bh = [(mk(cout2) * 0.1).cuda() for _ in range(2)]
    d = lib.make_desc(batch, n, cin, cmid, (3, 3), (1, 1), (1, 1), 1, False, True, False, True, lib.ACT_CAPPED_LEAKY_RELU, 0.1,
                      10.0, lib.BF16, lib.BF16, c0, m0, c1, lib.SRC_SAME)
    dh = lib.make_desc(batch, n, cmid, cout2, (1, 1), (1, 1), (1, 1), 0, False, True, False, True,
                       lib.ACT_CAPPED_LEAKY_RELU if act2 else lib.ACT_NONE, 0.1, 10.0, lib.BF16, lib.BF16)
    assert lib.conv2d_head_fusable(d, dh)
    packed = lib.pack_weights(d, w[0], w[1], None, b_[0], b_[1], None)
    packed_h = lib.pack_weights(dh, wh[0], wh[1], None, bh[0], bh[1], None)
    mid = lib.conv2d_fwd(d, x0, x1, packed)
    two = lib.conv2d_fwd(dh, mid, None, packed_h)
    one = lib.conv2d_fwd_head(d, x0, x1, packed, dh, packed_h)
    torch.cuda.synchronize()
    assert float(two.float().abs().max()) > 0.1
    assert torch.equal(one, two)
    # not fusable: a 3x3 head, a head that reads something else
    d3 = lib.make_desc(batch, n, cmid, cout2, (3, 3), (1, 1), (1, 1), 1, False, True, False, True, lib.ACT_NONE, 0.1, 10.0,
                       lib.BF16, lib.BF16)
    assert not lib.conv2d_head_fusable(d, d3)
    dw = lib.make_desc(batch, n, cmid + 8, cout2, (1, 1), (1, 1), (1, 1), 0, False, True, False, True, lib.ACT_NONE, 0.1, 10.0,
                       lib.BF16, lib.BF16)
    assert not lib.conv2d_head_fusable(d, dw)


def test_rollout_engine_fused_head_is_bit_identical():
    from dlwp_cs_b200.unet import CubeSphereUNet2, RolloutEngine
    torch.manual_seed(3)
    model = CubeSphereUNet2(18, 14, base=32).cuda()
    g = torch.Generator().manual_seed(4)
    state, forcing = torch.randn(3, 6, 24, 24, 14, generator=g), torch.rand(3, 6, 24, 24, 4, generator=g)
    outs = []
    for fuse in (True, False):
        eng = RolloutEngine(model, 3, 24, 3, forcing_channels=4, dtype=torch.bfloat16, fuse_head=fuse)
        assert (eng.fused_head is not None) == fuse
        assert eng.launches_per_step == (10 if fuse else 11)
        outs.append(eng.run(state, forcing).clone())
        torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1])


# ---- pooled second output (dlwpcs_conv2d_fwd_pool): the producer of a pooled layer writes the 2x2 mean itself ----------------
@pytest.mark.parametrize('batch,n,cin,cout,srcs', [(3, 24, 32, 32, 1), (5, 48, 24, 32, 1), (2, 8, 16, 24, 1), (7, 48, 64, 32, 2),
                                                    (1, 12, 32, 16, 1)])
def test_pool_output_equals_pooling_the_output(lib, batch, n, cin, cout, srcs):
    """y is bit-identical to the plain launch, y_pool == the 2x2 mean of the bf16 y taken in float32 and rounded to bf16 once
    (what the fused pooled load of the next layer computes); then the pooled layer gives the same result from either."""
    g = torch.Generator().manual_seed(13 * n + cin)
    mk = lambda *shape: torch.randn(*shape, generator=g)
    if srcs == 1:
        x0, x1, c0, c1, m0 = bf(mk(batch, 6, n, n, cin)).cuda(), None, cin, 0, lib.SRC_SAME
    else:
        ca = cin // 2
        x0, x1 = bf(mk(batch, 6, n // 2, n // 2, ca)).cuda(), bf(mk(batch, 6, n, n, cin - ca)).cuda()
        c0, c1, m0 = ca, cin - ca, lib.SRC_UP2
    w = [bf(mk(3, 3, cin, cout) * 0.1).float().cuda() for _ in range(2)]
    b_ = [(mk(cout) * 0.1).cuda() for _ in range(2)]
    d = lib.make_desc(batch, n, cin, cout, (3, 3), (1, 1), (1, 1), 1, False, True, False, True, lib.ACT_CAPPED_LEAKY_RELU, 0.1,
                      10.0, lib.BF16, lib.BF16, c0, m0, c1, lib.SRC_SAME)
    assert lib.conv2d_pool_fusable(d)
    packed = lib.pack_weights(d, w[0], w[1], None, b_[0], b_[1], None)
    y_ref = lib.conv2d_fwd(d, x0, x1, packed)
    y, yp = lib.conv2d_fwd_pool(d, x0, x1, packed)
    torch.cuda.synchronize()
    assert torch.equal(y, y_ref)
    f = y.float()
    want = (0.25 * (((f[:, :, 0::2, 0::2] + f[:, :, 0::2, 1::2]) + f[:, :, 1::2, 0::2]) + f[:, :, 1::2, 1::2])).bfloat16()
    assert torch.equal(yp, want)
    # the next layer: pooled load of y == plain load of y_pool, bit for bit (classic kernel both times: a 5x5 layer)
    w2 = [bf(mk(5, 5, cout, 16) * 0.1).float().cuda() for _ in range(2)]
    dp = lib.make_desc(batch, n // 2, cout, 16, (5, 5), (1, 1), (1, 1), 2, False, True, False, False, lib.ACT_NONE, 0.1, 10.0,
                       lib.BF16, lib.BF16, cout, lib.SRC_POOL2, 0, lib.SRC_SAME)
    ds = lib.copy_desc(dp, mode0=lib.SRC_SAME)
    pk2 = lib.pack_weights(dp, w2[0], w2[1], None, None, None, None)
    za = lib.conv2d_fwd(dp, y, None, pk2)
    zb = lib.conv2d_fwd(ds, yp, None, pk2)
    torch.cuda.synchronize()
    assert torch.equal(za, zb)
    # not fusable: more than 32 output channels, odd face edge
    assert not lib.conv2d_pool_fusable(lib.copy_desc(d, cout=64))
    assert not lib.conv2d_pool_fusable(lib.make_desc(batch, 9, cin, cout, (3, 3), (1, 1), (1, 1), 1, False, True, False, True,
                                                     lib.ACT_NONE, 0.1, 10.0, lib.BF16, lib.BF16))


def test_rollout_engine_pool_outputs():
    from dlwp_cs_b200.unet import CubeSphereUNet2, RolloutEngine
    torch.manual_seed(3)
    model = CubeSphereUNet2(18, 14, base=32).cuda()
    g = torch.Generator().manual_seed(4)
    state, forcing = torch.randn(3, 6, 24, 24, 14, generator=g), torch.rand(3, 6, 24, 24, 4, generator=g)
    outs = []
    for pool in (True, False):
        eng = RolloutEngine(model, 3, 24, 3, forcing_channels=4, dtype=torch.bfloat16, pool_out=pool)
        assert (len(eng.pool_out) == 1) == pool          # conv_2d_1_2 (32 channels) feeds the pooled conv_2d_2 (on by default)
        outs.append(eng.run(state, forcing).float().clone())
        torch.cuda.synchronize()
    # the pooled layer runs on the other kernel (another float32 summation order): equal up to the bf16 rounding of layers
    scale = float(outs[1].abs().max())
    assert float((outs[0] - outs[1]).abs().max()) <= 2.0 ** -6 * scale
