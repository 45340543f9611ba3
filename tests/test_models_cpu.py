"""Wrapper-level boundary on the CPU (SURVEY.md section 8 a9 / b):

  * the output arithmetic of ``predict_timeseries`` (dlwp_cs_b200.models.assemble_timeseries) against fixtures produced by
    the reference's OWN ``DLWPTorchNN.predict_timeseries`` (tests/golden/make_golden_timeseries.py);
  * the reference's ``DLWPTorchNN.build_model`` (models_torch.py:127-163) resolving ``CubeSpherePadding2D`` /
    ``CubeSphereConv2D`` BY NAME from a ``DLWP.custom`` that re-exports the engine's classes.
    Needs /root/reference, which only exists in the build container -- skipped elsewhere.
"""
import importlib
import os
import sys
import types

import numpy as np
import pytest
import torch

import cs_oracle as O

REFERENCE = '/root/reference'


@pytest.fixture(scope='module')
def ts(golden_dir):
    return np.load(os.path.join(golden_dir, 'timeseries.npz'))


def oracle_steps(ts, steps, step_sequence):
    """The stacked model outputs (steps, B, C, 6, N, N) of the oracle network -- the loop restated independently of the
    reference's code (models.py:272-294)."""
    n, b, v, td, base, seed = [int(k) for k in ts['meta']]
    c = v * td
    params = {k: w.double() for k, w in O.make_unet2_params(c, c, base=base, seed=seed).items()}
    p = torch.from_numpy(ts['predictors']).double()
    outs = []
    with torch.no_grad():
        for _ in range(steps):
            pr = O.unet2(params, p.permute(0, 2, 3, 4, 1)).permute(0, 4, 1, 2, 3).float().double()
            outs.append(pr)
            p = torch.cat([p[:, v:], pr[:, :v]], dim=1) if step_sequence else pr
    return torch.stack(outs).numpy().astype(np.float32)


@pytest.mark.parametrize('name,time_steps,seq,keep', [('plain5', 5, False, False), ('plain5_keep', 5, False, True),
                                                      ('seq3', 3, True, False), ('seq3_keep', 3, True, True),
                                                      ('plain2', 2, False, False)])
def test_assemble_timeseries_matches_reference_method(ts, name, time_steps, seq, keep):
    from dlwp_cs_b200.models import assemble_timeseries
    td = int(ts['meta'][3])
    steps = time_steps if seq else -(-time_steps // td)
    got = assemble_timeseries(oracle_steps(ts, steps, seq), td, seq, keep)
    ref = ts[name]
    assert got.shape == ref.shape and got.dtype == np.float32
    np.testing.assert_allclose(got, ref, rtol=0, atol=2e-6)


def test_forecast_cs_layout():
    from dlwp_cs_b200.models import forecast_cs_layout
    a = np.arange(2 * 3 * 6 * 4 * 4 * 5, dtype=np.float32).reshape(2, 3, 6, 4, 4, 5)
    f = forecast_cs_layout(a, 'channels_last')
    assert f.shape == (2, 3, 5, 6, 4, 4) and f[1, 2, 3, 4, 1, 2] == a[1, 2, 4, 1, 2, 3]
    assert forecast_cs_layout(f, 'channels_first') is not None
    with pytest.raises(ValueError):
        forecast_cs_layout(a[0], 'channels_last')


def test_forecaster_argument_checks():
    from dlwp_cs_b200.models import CubeSphereForecaster
    from dlwp_cs_b200.unet import CubeSphereUNet2
    m = CubeSphereUNet2(4, 4, base=8)
    with pytest.raises(ValueError):
        CubeSphereForecaster(m, time_dim=0)
    with pytest.raises(ValueError):
        CubeSphereForecaster(m, data_format='nchw')
    fc = CubeSphereForecaster(m, time_dim=2, data_format='channels_first')
    with pytest.raises(ValueError):
        fc.predict_timeseries(np.zeros((1, 4, 6, 8, 8), np.float32), 0)
    with pytest.raises(ValueError):                       # 5 channels into a 4-channel model
        fc.predict_timeseries(np.zeros((1, 5, 6, 8, 8), np.float32), 2)
    with pytest.raises(ValueError):                       # faces on the wrong axis for channels_first
        fc.predict_timeseries(np.zeros((1, 6, 8, 8, 4), np.float32), 2)


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason='the reference tree only exists in the build container')
def test_reference_build_model_resolves_engine_layers_by_name():
    """models_torch.py:127-163 unchanged: by-name lookup in torch.nn then DLWP.custom, the activation kwarg popped into
    ``activations``, ``setattr(model, 'layer%d')`` registering the parameters, optimizer over ``model.parameters()``."""
    import torch._dynamo  # noqa: F401  (before the shim: torch probes find_spec('tensorflow') lazily)
    import tf_shim
    tf_shim.load_reference_generators()
    import dlwp_cs_b200.custom as ours
    saved = sys.modules.get('DLWP.custom')
    mod = types.ModuleType('DLWP.custom')                 # DLWP/custom.py re-exports the engine's classes
    mod.CubeSpherePadding2D, mod.CubeSphereConv2D = ours.CubeSpherePadding2D, ours.CubeSphereConv2D
    sys.modules['DLWP.custom'] = mod
    try:
        models_torch = importlib.import_module('DLWP.model.models_torch')
        dlwp = models_torch.DLWPTorchNN(is_convolutional=True, time_dim=2, scaler_type=None, scale_targets=False)
        layers = (('CubeSpherePadding2D', (1,), {'data_format': 'channels_last'}),
                  ('CubeSphereConv2D', (8, 3), {'data_format': 'channels_last', 'activation': 'relu', 'in_channels': 4}),
                  ('CubeSpherePadding2D', (1,), {'data_format': 'channels_last'}),
                  ('CubeSphereConv2D', (4, 3), {'data_format': 'channels_last', 'activation': None,
                                                'independent_north_pole': True, 'flip_north_pole': False}))
        dlwp.build_model(layers, 'Adam', 'MSELoss', optimizer_kwargs={'lr': 1e-3})
        assert [type(l).__name__ for l in dlwp.layers] == ['CubeSpherePadding2D', 'CubeSphereConv2D'] * 2
        assert all(isinstance(l, (ours.CubeSpherePadding2D, ours.CubeSphereConv2D)) for l in dlwp.layers)
        assert dlwp.activations[1] is torch.nn.functional.relu and dlwp.activations[3] is None
        names = [n for n, _ in dlwp.model.named_parameters()]
        assert names[:4] == ['layer1.equatorial_kernel', 'layer1.polar_kernel', 'layer1.equatorial_bias',
                             'layer1.polar_bias']                                  # custom.py:882-914 creation order
        assert 'layer3.north_pole_kernel' in names and 'layer3.north_pole_bias' in names
        assert tuple(dlwp.model.layer1.equatorial_kernel.shape) == (3, 3, 4, 8)    # HWIO
        assert dlwp.model.layer3.has_uninitialized_params()                        # keras-style lazy build (custom.py:871-919)
        # the wrapper's forward reaches the engine's operator, which refuses CPU tensors instead of falling back
        from dlwp_cs_b200 import _lib
        _lib.load()
        with pytest.raises(_lib.DlwpcsError):
            dlwp.predict(np.zeros((1, 6, 8, 8, 4), np.float32))
        dlwp.reset()                                                               # models_torch.py:400-410
    finally:
        if saved is not None:
            sys.modules['DLWP.custom'] = saved
        else:
            sys.modules.pop('DLWP.custom', None)


@pytest.mark.parametrize('arch', ['basic', 'unet', 'unet2', 'unet3', 'unet4'])
def test_architecture_programs_match_the_oracle_restatement(arch):
    """The layer programs the engine derives its fused launch plan from (dlwp_cs_b200.unet.ARCHS) against the statement-by-
    statement restatement of Azure/train_cs.py:233-388 in the oracle: same layers, kernel sizes and channel counts."""
    from dlwp_cs_b200.unet import CubeSphereCNN, arch_program
    prog = arch_program(arch, 18, 14, 32)
    assert [(s['name'], s['kernel'], s['cin'], s['cout']) for s in prog] == O.arch_shapes(arch, 18, 14, 32)
    assert prog[-1]['dst'] == 'out' and prog[-1]['kernel'] == 1 and not prog[-1]['act']
    for s in prog:                       # at most two sources, only the first one resampled: what one fused launch can read
        assert 1 <= len(s['sources']) <= 2 and (len(s['sources']) == 1 or s['sources'][1][2] == 'same')
    m = CubeSphereCNN(arch, 18, 14, 32)
    assert sum(p.numel() for p in m.parameters()) == sum(2 * (k * k * ci * co + co) for _, k, ci, co in O.arch_shapes(arch, 18, 14, 32))
    assert m.levels == (3 if arch == 'unet4' else 2)


def test_unknown_architecture_and_bad_edge():
    from dlwp_cs_b200.unet import CubeSphereCNN, arch_program
    with pytest.raises(ValueError):
        arch_program('unet5', 4, 4)
    assert CubeSphereCNN('unet4', 4, 4, base=8).levels == 3
