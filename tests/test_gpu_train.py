"""GPU tests of the training-step kernels and the data-parallel trainer (single rank here; the all-reduce logic is
covered on CPU with gloo in tests/test_train_cpu.py and on 2+ GPUs by bench.py --gpus N)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import cs_oracle as O  # noqa: E402


@pytest.fixture(scope='module')
def lib():
    from dlwp_cs_b200 import _lib
    _lib.load()
    return _lib


def keras_adam(p, g, m, v, lr, b1, b2, eps, t):
    """numpy restatement of keras.optimizers.Adam (the optimizer of Azure/train_cs.py:424)."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    lr_t = lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    return p - lr_t * m / (np.sqrt(v) + eps), m, v


def test_adam_step_matches_keras_formula(lib):
    rng = np.random.default_rng(0)
    n = 100003
    p, g = rng.standard_normal(n), rng.standard_normal(n) * 0.1
    m, v = np.zeros(n), np.zeros(n)
    dp, dm, dv = [torch.tensor(a, dtype=torch.float32).cuda() for a in (p, m, v)]
    for t in range(1, 4):
        gt = g * t
        dg = torch.tensor(gt * 4.0, dtype=torch.float32).cuda()          # summed over 4 "ranks"
        lib.adam_step(dp, dg, dm, dv, 1e-3, 0.9, 0.999, 1e-7, t, grad_scale=0.25)
        p, m, v = keras_adam(p, gt, m, v, 1e-3, 0.9, 0.999, 1e-7, t)
        # the hyper-parameters cross the C ABI as float32: 1 - 0.999f differs from 0.001 by 1.3e-5 relative
        np.testing.assert_allclose(dp.cpu().numpy(), p, rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(dv.cpu().numpy(), v, rtol=5e-5, atol=1e-12)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_mse_loss_grad(lib, dtype):
    g = torch.Generator().manual_seed(2)
    y = torch.randn(3, 6, 8, 8, 5, generator=g).to(dtype)
    t = torch.randn(3, 6, 8, 8, 5, generator=g).to(dtype)
    loss = torch.zeros(1, device='cuda')
    dy = lib.mse_loss_grad(y.cuda(), t.cuda(), loss)
    d = y.double() - t.double()
    assert abs(float(loss) - float((d * d).mean())) < 1e-5 * float((d * d).mean())
    ref = 2 * d / d.numel()
    tol = 1e-6 if dtype == torch.float32 else 2.0 ** -8
    np.testing.assert_allclose(dy.double().cpu().numpy(), ref.numpy(), rtol=tol, atol=tol * float(ref.abs().max()))


def test_trainer_step_matches_oracle_autograd_and_keras_adam(lib):
    """One optimizer step of the fp32 U-Net through the engine == oracle autograd (float64) + the Keras Adam formula."""
    from dlwp_cs_b200.unet import CubeSphereUNet2
    from dlwp_cs_b200.train import DataParallelTrainer
    n, b, cin, cout, base = 8, 2, 5, 3, 8
    params = O.make_unet2_params(cin, cout, base=base, seed=4)
    model = CubeSphereUNet2(cin, cout, base=base).cuda()
    model.load_oracle_params(params)
    names = [k for k, _ in model.named_parameters()]
    g = torch.Generator().manual_seed(9)
    x = torch.randn(b, 6, n, n, cin, generator=g)
    t = torch.randn(b, 6, n, n, cout, generator=g)
    # oracle
    pd = {k: params[k].double().requires_grad_(True) for k in names}
    loss_ref = (O.unet2(pd, x.double()) - t.double()).pow(2).mean()
    loss_ref.backward()
    trainer = DataParallelTrainer(model, lr=1e-3)
    loss = trainer.step(x.cuda(), t.cuda())
    assert abs(float(loss.detach()) - float(loss_ref.detach())) < 1e-5 * float(loss_ref.detach())
    for k, p in model.named_parameters():
        gref = pd[k].grad.numpy()
        # gradient parity (the flat buffer still holds this step's gradients)
        np.testing.assert_allclose(p.grad.double().cpu().numpy(), gref, rtol=1e-4,
                                   atol=1e-5 * max(float(np.abs(gref).max()), 1e-30))
        pref, _, _ = keras_adam(params[k].double().numpy(), gref, 0.0, 0.0, 1e-3, 0.9, 0.999, 1e-7, 1)
        # the first Adam step moves every weight by ~lr * sign(g); tiny gradients make the direction itself ill-conditioned
        big = np.abs(gref) > 1e-4 * np.abs(gref).max()
        np.testing.assert_allclose(p.detach().double().cpu().numpy()[big], pref[big], rtol=0, atol=2e-5)


def test_trainer_graph_replay_equals_eager_steps(lib):
    """Three bf16 optimizer steps replayed from the captured CUDA graph (device-side Adam step counter) leave exactly the
    parameters the eager launches leave; inputs change every step."""
    from dlwp_cs_b200.unet import CubeSphereUNet2
    from dlwp_cs_b200.train import DataParallelTrainer
    n, b, cin, cout, base = 8, 2, 6, 4, 8
    params = O.make_unet2_params(cin, cout, base=base, seed=11)
    g = torch.Generator().manual_seed(3)
    xs = [torch.randn(b, 6, n, n, cin, generator=g).cuda().bfloat16() for _ in range(3)]
    ts = [torch.randn(b, 6, n, n, cout, generator=g).cuda().bfloat16() for _ in range(3)]
    out = []
    for use_graph in (False, True):
        model = CubeSphereUNet2(cin, cout, base=base).cuda()
        model.load_oracle_params(params)
        tr = DataParallelTrainer(model, lr=1e-2, use_graph=use_graph)
        losses = [float(tr.step(x, t)) for x, t in zip(xs, ts)]
        assert int(tr.step_counter) == 3
        out.append((tr.flat.param.clone(), losses))
    assert torch.equal(out[0][0], out[1][0])
    # the loss is a float atomicAdd over blocks: equal up to summation order
    np.testing.assert_allclose(out[0][1], out[1][1], rtol=1e-5)
    assert out[0][1][2] < out[0][1][0]          # and it learns


def test_trainer_sequence_step_matches_oracle_autograd(lib):
    """Two-step ("integration_steps = 2", Azure/train_cs.py:101, 391-426) optimizer step: the same layers applied twice, the
    second input re-packed from the first output + the next insolation + the constants, loss = (mse_0 + mse_1) / 2 --
    float32 engine (graph-replayed) against float64 oracle autograd: loss and every parameter gradient."""
    from dlwp_cs_b200.unet import CubeSphereUNet2
    from dlwp_cs_b200.train import DataParallelTrainer, repack_reference
    n, b, n_var, t_in, n_const, base = 8, 2, 3, 2, 2, 8
    cin, cout = t_in * (n_var + 1) + n_const, t_in * n_var
    params = O.make_unet2_params(cin, cout, base=base, seed=6)
    model = CubeSphereUNet2(cin, cout, base=base).cuda()
    model.load_oracle_params(params)
    names = [k for k, _ in model.named_parameters()]
    g = torch.Generator().manual_seed(10)
    x = torch.randn(b, 6, n, n, cin, generator=g)
    solar = torch.rand(b, t_in, 6, n, n, 1, generator=g)
    ts = [torch.randn(b, 6, n, n, cout, generator=g) for _ in range(2)]
    pd = {k: params[k].double().requires_grad_(True) for k in names}
    y0 = O.unet2(pd, x.double())
    y1 = O.unet2(pd, repack_reference(y0, solar.double(), x.double(), t_in, n_var))
    loss_ref = 0.5 * ((y0 - ts[0].double()).pow(2).mean() + (y1 - ts[1].double()).pow(2).mean())
    loss_ref.backward()
    trainer = DataParallelTrainer(model, lr=1e-3)
    loss = trainer.step_sequence(x.cuda(), [solar.cuda()], [t.cuda() for t in ts], t_in, n_var)
    assert abs(float(loss.detach()) - float(loss_ref.detach())) < 1e-5 * float(loss_ref.detach())
    for k, p in model.named_parameters():
        gref = pd[k].grad.numpy()
        np.testing.assert_allclose(p.grad.double().cpu().numpy(), gref, rtol=1e-4,
                                   atol=1e-5 * max(float(np.abs(gref).max()), 1e-30))
    assert int(trainer.step_counter) == 1


def test_mse_loss_grad_rejects_mismatched_target(lib):
    """A float32 target under a bf16 output (or a smaller target) used to be reinterpreted / read out of bounds."""
    y = torch.randn(2, 6, 8, 8, 8).cuda().bfloat16()
    loss = torch.zeros(1, device='cuda')
    with pytest.raises(lib.DlwpcsError):
        lib.mse_loss_grad(y, torch.randn(2, 6, 8, 8, 8).cuda(), loss)                   # dtype
    with pytest.raises(lib.DlwpcsError):
        lib.mse_loss_grad(y, torch.randn(1, 6, 8, 8, 8).cuda().bfloat16(), loss)       # shape
    with pytest.raises(lib.DlwpcsError):
        lib.mse_loss_grad(y, y.clone(), torch.zeros(2, device='cuda'))                  # accumulator
    lib.mse_loss_grad(y, y.clone(), loss)
    assert float(loss) == 0.0


def test_trainer_hyperparameter_change_recaptures(lib):
    """lr is baked into the captured step: assigning a new one must take effect."""
    from dlwp_cs_b200.unet import CubeSphereUNet2
    from dlwp_cs_b200.train import DataParallelTrainer
    torch.manual_seed(0)
    model = CubeSphereUNet2(6, 4, base=8).cuda()
    tr = DataParallelTrainer(model, lr=1e-3)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 6, 8, 8, 6, generator=g).cuda()
    t = torch.randn(2, 6, 8, 8, 4, generator=g).cuda()
    tr.step(x, t)
    p0 = tr.flat.param.clone()
    tr.lr = 0.0                                  # frozen: the next step must not move the parameters
    assert not tr._graphs
    tr.step(x, t)
    assert torch.equal(tr.flat.param, p0)
    tr.lr = 1e-2
    tr.step(x, t)
    assert not torch.equal(tr.flat.param, p0)
    tr.close()
    assert not tr._graphs


def test_pack_weights2_equals_separate_packs(lib):
    """One launch for the forward and the transposed image; kernels zero-extended to the padded channel counts == packing
    explicitly padded copies."""
    g = torch.Generator().manual_seed(3)
    cin, cout, cinp, coutp = 18, 14, 24, 16
    ws = [torch.randn(3, 3, cin, cout, generator=g).cuda() for _ in range(3)]
    bs = [torch.randn(cout, generator=g).cuda() for _ in range(3)]
    d = lib.make_desc(2, 8, cinp, coutp, (3, 3), halo=1, independent_north_pole=True, x_dtype=lib.BF16, y_dtype=lib.BF16)
    fpad = torch.nn.functional.pad
    wp = [fpad(w, (0, coutp - cout, 0, cinp - cin)) for w in ws]
    bp = [fpad(b, (0, coutp - cout)) for b in bs]
    ref_f = lib.pack_weights(d, wp[0], wp[1], wp[2], bp[0], bp[1], bp[2])
    ref_t = lib.pack_weights(d, wp[0], wp[1], wp[2], transposed=True)
    got_f, got_t = lib.pack_weights2(d, ws[0], ws[1], ws[2], bs[0], bs[1], bs[2])
    assert torch.equal(got_f, ref_f) and torch.equal(got_t, ref_t)
    only_t = lib.pack_weights2(d, ws[0], ws[1], ws[2], forward=False)
    assert only_t[0] is None and torch.equal(only_t[1], ref_t)


@pytest.mark.parametrize('dtype', [torch.bfloat16, torch.float32])
@pytest.mark.parametrize('halo,k', [(1, 3), (0, 1)])
def test_dgrad_act_equals_dgrad_then_activation_derivative(lib, dtype, halo, k):
    """dlwpcs_conv2d_dgrad_act: the mask of the activation that produced the layer's input, fused into the halo
    scatter-add (bf16, halo > 0) or applied in place (1x1 layers / float32) == dgrad followed by dlwpcs_act_bwd."""
    g = torch.Generator().manual_seed(5)
    b, n, cin, cout = 2, 12, 32, 16
    code = lib.dtype_code(dtype)
    x = (torch.randn(b, 6, n, n, cin, generator=g) * 6).to(dtype).cuda()          # values on both sides of 0 and of the cap 10
    dy = torch.randn(b, 6, n, n, cout, generator=g).to(dtype).cuda()
    ws = [torch.randn(k, k, cin, cout, generator=g).cuda() * 0.2 for _ in range(2)]
    d = lib.make_desc(b, n, cin, cout, (k, k), halo=halo, x_dtype=code, y_dtype=code)
    packed_t = lib.pack_weights(d, ws[0], ws[1], transposed=True)
    plain = lib.conv2d_dgrad(d, dy, None, packed_t)
    ref = lib.act_bwd(plain, x, lib.ACT_CAPPED_LEAKY_RELU, 0.1, 10.0)
    got = lib.conv2d_dgrad(d, dy, None, packed_t, x_in=x, in_act=(lib.ACT_CAPPED_LEAKY_RELU, 0.1, 10.0))
    if dtype == torch.bfloat16 and halo > 0:
        # fused: the mask multiplies the float32 sum before the single rounding; the reference rounds twice
        np.testing.assert_allclose(got.float().cpu().numpy(), ref.float().cpu().numpy(), rtol=2.0 ** -7, atol=1e-6)
    else:
        assert torch.equal(got, ref)


def test_batch_pack_equals_per_layer_pack_and_trainer_step():
    """dlwpcs_pack_weights_batch: the images of every layer in one launch are byte-identical to the per-layer packs, and an
    optimizer step that uses them leaves exactly the same parameters as one that packs inside the layers' forward."""
    from dlwp_cs_b200 import _lib, functional as F_cs
    from dlwp_cs_b200.unet import CubeSphereUNet2
    from dlwp_cs_b200.train import DataParallelTrainer
    torch.manual_seed(2)
    model = CubeSphereUNet2(18, 14, base=16).cuda()
    res = F_cs.prepack_model(model, torch.bfloat16)
    F_cs.clear_prepacked()
    layers = model.layers_in_order()
    assert len(res) == len(layers)
    for i, (layer, (packed, packed_t)) in enumerate(zip(layers, res)):
        d, ws = layer.pack_descriptor(torch.bfloat16)
        one, one_t = _lib.pack_weights2(d, *ws, forward=True, transposed=i > 0)
        assert torch.equal(packed, one)
        assert (packed_t is None) == (i == 0)
        if i > 0:
            assert torch.equal(packed_t, one_t)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 6, 16, 16, 18, generator=g).cuda().bfloat16()
    t = torch.randn(2, 6, 16, 16, 14, generator=g).cuda().bfloat16()
    finals = []
    for prepack in (True, False):
        torch.manual_seed(2)
        m = CubeSphereUNet2(18, 14, base=16).cuda()
        tr = DataParallelTrainer(m, lr=1e-3, use_graph=False, distributed=False)
        tr.prepack = prepack
        for _ in range(2):
            tr.step(x, t)
        torch.cuda.synchronize()
        finals.append(torch.cat([p.detach().flatten() for p in m.parameters()]).clone())
        tr.close()
    assert torch.equal(finals[0], finals[1])
