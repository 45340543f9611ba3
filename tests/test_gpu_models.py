"""GPU parity at the wrapper level and at the configurations the bench numbers are quoted on.

  * ``CubeSphereForecaster.predict_timeseries`` (dlwp_cs_b200/models.py) against fixtures produced by the reference's OWN
    ``DLWPTorchNN.predict_timeseries`` (tests/golden/make_golden_timeseries.py): plain / step_sequence / keep_time_dim;
  * the 100-step C48 rollout of BASELINE.json configs[1] (small ensemble): float32 engine and bf16 engine against the
    float64 oracle at EVERY step, with a bound on the growth of the error along the rollout;
  * C96 / 12 variables (configs[4]) multi-step.

The bf16 contract: per layer bf16 activations and weights, fp32 accumulation in tensor memory, ONE rounding to bf16 when
the layer output is stored (the fused 2x2 mean is rounded once more, like a stored pooled tensor).  With the random-init
(glorot) weights the bench uses, the network is contractive (the state settles on a forcing-dependent fixed point), so
the per-step error stays at the single-step level (~5e-3 of the field's max) instead of compounding.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import cs_oracle as O  # noqa: E402


@pytest.fixture(scope='module')
def ts(golden_dir):
    return np.load(os.path.join(golden_dir, 'timeseries.npz'))


def _forecaster(ts, dtype):
    from dlwp_cs_b200.models import CubeSphereForecaster
    from dlwp_cs_b200.unet import CubeSphereUNet2
    n, b, v, td, base, seed = [int(k) for k in ts['meta']]
    c = v * td
    model = CubeSphereUNet2(c, c, base=base).cuda()
    model.load_oracle_params(O.make_unet2_params(c, c, base=base, seed=seed))
    return CubeSphereForecaster(model, time_dim=td, data_format='channels_first', dtype=dtype)


CASES = [('plain5', 5, False, False), ('plain5_keep', 5, False, True), ('seq3', 3, True, False),
         ('seq3_keep', 3, True, True), ('plain2', 2, False, False)]


@pytest.mark.parametrize('name,time_steps,seq,keep', CASES)
def test_predict_timeseries_fp32_vs_reference_method(ts, name, time_steps, seq, keep):
    fc = _forecaster(ts, torch.float32)
    got = fc.predict_timeseries(ts['predictors'], time_steps, step_sequence=seq, keep_time_dim=keep)
    ref = ts[name]
    assert got.shape == ref.shape and got.dtype == np.float32
    scale = float(np.abs(ref).max())
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-5 * scale)


@pytest.mark.parametrize('name,time_steps,seq,keep', CASES[:1] + CASES[2:3])
def test_predict_timeseries_bf16_vs_reference_method(ts, name, time_steps, seq, keep):
    fc = _forecaster(ts, torch.bfloat16)
    got = fc.predict_timeseries(ts['predictors'], time_steps, step_sequence=seq, keep_time_dim=keep)
    ref = ts[name]
    assert got.shape == ref.shape
    assert float(np.abs(got - ref).max()) <= 1.5e-2 * float(np.abs(ref).max())


def test_predict_and_forecast_layout(ts):
    fc = _forecaster(ts, torch.float32)
    one = fc.predict(ts['predictors'])
    series = fc.predict_timeseries(ts['predictors'], 2)                  # time_dim = 2 -> one model step
    np.testing.assert_array_equal(series.shape, (2, 2, 2, 6, 8, 8))
    # step 0 of the series: time step t of sample s is channels [t*V, (t+1)*V) of predict()
    np.testing.assert_allclose(series[1, 0], one[0, 2:4], rtol=0, atol=0)
    f = fc.forecast(ts['predictors'], 4)
    assert f.shape == (4, 2, 2, 6, 8, 8)
    # channels_last predictors give the same numbers
    from dlwp_cs_b200.models import CubeSphereForecaster
    fl = CubeSphereForecaster(fc.model, time_dim=2, data_format='channels_last', dtype=torch.float32)
    one_l = fl.predict(np.ascontiguousarray(ts['predictors'].transpose(0, 2, 3, 4, 1)))
    np.testing.assert_array_equal(one_l.transpose(0, 4, 1, 2, 3), one)


def _rollout_errors(n, cp, cf, batch, steps, seed=1, bias_scale=0.05):
    """relative max error (of the oracle field's max, per step) of the float32 and the bf16 engine vs the float64 oracle."""
    from dlwp_cs_b200.unet import CubeSphereUNet2, RolloutEngine
    params = O.make_unet2_params(cp + cf, cp, base=32, seed=seed, bias_scale=bias_scale)
    model = CubeSphereUNet2(cp + cf, cp, base=32).cuda()
    model.load_oracle_params(params)
    g = torch.Generator().manual_seed(0)
    state = torch.randn(batch, 6, n, n, cp, generator=g).clamp_(-5, 5)
    forcing = torch.rand(batch, 6, n, n, cf, generator=g)
    # the bf16 engine sees bf16-rounded initial conditions; give the oracle the same ones for that comparison
    p64 = {k: v.double() for k, v in params.items()}
    with torch.no_grad():
        ref32 = O.rollout(p64, state.double(), forcing.double(), steps, exact=False)
        ref16 = O.rollout(p64, state.bfloat16().double(), forcing.bfloat16().double(), steps, exact=False)
    out = {}
    for dtype, ref in ((torch.float32, ref32), (torch.bfloat16, ref16)):
        eng = RolloutEngine(model, batch, n, steps, forcing_channels=cf, dtype=dtype)
        got = eng.run(state.cuda(), forcing.cuda()).double().cpu()
        torch.cuda.synchronize()
        assert torch.isfinite(got).all()
        err = (got - ref).abs().flatten(1).max(dim=1).values / ref.abs().flatten(1).max(dim=1).values
        out[dtype] = err.numpy()
    return out


def test_rollout_100_steps_c48_error_growth():
    """BASELINE.json configs[1]: C48, 7 variables x 2 time steps (+2 insolation +2 constants), 100 6-hour steps."""
    err = _rollout_errors(48, 14, 4, 2, 100)
    e32, e16 = err[torch.float32], err[torch.bfloat16]
    print('fp32 rel err at steps 1,2,5,10,50,100:', [float('%.2e' % e32[i]) for i in (0, 1, 4, 9, 49, 99)])
    print('bf16 rel err at steps 1,2,5,10,50,100:', [float('%.2e' % e16[i]) for i in (0, 1, 4, 9, 49, 99)])
    assert e32.max() <= 1e-4, e32.max()                       # observed ~1e-6: float32 accumulation order only
    assert e16.max() <= 1.5e-2, e16.max()                     # one bf16 rounding per stored layer output, 11 layers deep
    # no compounding along the rollout: the last steps are no worse than twice the worst of the first five
    assert e16[50:].max() <= 2.0 * e16[:5].max() + 1e-3
    assert e32[50:].max() <= 2.0 * e32[:5].max() + 1e-5


def test_rollout_c96_12var_multi_step():
    """BASELINE.json configs[4]: C96, 12 variables x 2 time steps in/out (+2 +2), here 4 steps of one member."""
    err = _rollout_errors(96, 24, 4, 1, 4)
    assert err[torch.float32].max() <= 1e-4, err[torch.float32]
    assert err[torch.bfloat16].max() <= 1.5e-2, err[torch.bfloat16]


@pytest.mark.parametrize('arch', ['basic', 'unet', 'unet3', 'unet4'])
def test_other_architectures_forward_and_rollout(arch):
    """The other cubed-sphere networks of Azure/train_cs.py:233-388 through the same layers: module forward (float32) and a
    2-step device-resident rollout (float32 and bf16), whose fused launch plan is derived from the layer program, against
    the oracle's statement-by-statement restatement."""
    from dlwp_cs_b200.unet import CubeSphereCNN, RolloutEngine
    n, b, cp, cf, base = 16, 2, 6, 2, 8
    params = O.make_arch_params(arch, cp + cf, cp, base=base, seed=9)
    model = CubeSphereCNN(arch, cp + cf, cp, base=base).cuda()
    model.load_oracle_params(params)
    g = torch.Generator().manual_seed(4)
    state = torch.randn(b, 6, n, n, cp, generator=g)
    forcing = torch.rand(b, 6, n, n, cf, generator=g)
    p64 = {k: v.double() for k, v in params.items()}
    with torch.no_grad():
        x = torch.cat([state, forcing], dim=-1)
        ref1 = O.cs_network(arch, p64, x.double())
        y = model(x.cuda())
        scale = float(ref1.abs().max())
        assert float((y.double().cpu() - ref1).abs().max()) <= 1e-5 * scale
        ref2 = O.cs_network(arch, p64, torch.cat([ref1, forcing.double()], dim=-1))
    for dtype, tol in ((torch.float32, 1e-4), (torch.bfloat16, 1.5e-2)):
        eng = RolloutEngine(model, b, n, 2, forcing_channels=cf, dtype=dtype)
        # bf16: the 1x1 output layer runs inside the launch of the 3x3 layer in front of it (dlwpcs_conv2d_fwd_head)
        assert eng.launches_per_step == len(model.program) - (1 if eng.fused_head is not None else 0)
        assert (eng.fused_head is not None) == (dtype == torch.bfloat16)
        out = eng.run(state.cuda(), forcing.cuda()).double().cpu()
        torch.cuda.synchronize()
        for t, ref in enumerate((ref1, ref2)):
            err = float((out[t] - ref).abs().max()) / float(ref.abs().max())
            assert err <= tol, (arch, dtype, t, err)
