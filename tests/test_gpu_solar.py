"""GPU parity of the forced rollout: the insolation kernel against the oracle restatement of
DLWP/util.py:306-364 (itself pinned bit-for-bit against the reference's function), and the device-resident forced
rollout against the oracle loop of TimeSeriesEstimator.predict (extensions.py:259-308)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import cs_oracle as O  # noqa: E402
import cs_solar as S  # noqa: E402


@pytest.fixture(scope='module')
def lib():
    from dlwp_cs_b200 import _lib
    _lib.load()
    return _lib


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_insolation_kernel_vs_oracle(lib, dtype):
    n, b, c, c_first, n_sol = 12, 5, 8, 2, 3
    lat, lon = S.cubed_sphere_latlon(n)
    days = np.array([[0.0, 79.25, 171.5, 265.125, 364.875], [0.25, 79.5, 171.75, 265.375, 365.125],
                     [10.5, 100.0, 200.75, 300.5, 59.99]])
    out = torch.full((b, 6, n, n, c), -7.0, dtype=dtype, device='cuda')
    lib.insolation(out, c_first, n_sol, torch.from_numpy(np.sin(np.radians(lat)).reshape(-1)).cuda(),
                   torch.from_numpy(np.cos(np.radians(lat)).reshape(-1)).cuda(),
                   torch.from_numpy(lon.astype(np.float32).reshape(-1)).cuda(), torch.from_numpy(days).cuda(), S=1.0)
    got = out.float().cpu().numpy()
    assert np.all(got[..., :c_first] == -7.0) and np.all(got[..., c_first + n_sol:] == -7.0)      # other channels untouched
    for k in range(n_sol):
        ref = S.insolation(days[k], lat, lon)                                               # (B,6,n,n) float32
        tol = 2e-6 if dtype == torch.float32 else 2.0 ** -8
        np.testing.assert_allclose(got[..., c_first + k], ref, rtol=tol if dtype == torch.bfloat16 else 0, atol=tol)


def test_forced_rollout_vs_oracle(lib):
    """Three forecast iterations with the insolation recomputed on the device every iteration, reference channel packing
    [vars(t0), sol(t0), vars(t1), sol(t1), constants] (generators.py:880-899), a start date that crosses 31 December."""
    from dlwp_cs_b200.unet import CubeSphereUNet2, RolloutEngine, reference_input_order
    n, b, n_var, t_in, n_const, steps, base = 8, 2, 3, 2, 2, 3, 8
    cp, cf = t_in * n_var, t_in + n_const
    params = O.make_unet2_params(cp + cf, cp, base=base, seed=5)
    model = CubeSphereUNet2(cp + cf, cp, base=base).cuda()
    model.load_oracle_params(params)
    g = torch.Generator().manual_seed(2)
    state = torch.randn(b, 6, n, n, cp, generator=g)
    consts = torch.rand(b, 6, n, n, n_const, generator=g)
    lat, lon = S.cubed_sphere_latlon(n)
    dates = np.array(['2016-12-31T06:00', '2015-06-30T18:00'], dtype='datetime64[s]')
    dt_days = 0.25
    order = reference_input_order(t_in, n_var, n_const)
    # oracle loop
    p = state.double()
    ref = []
    for s in range(steps):
        sol = np.stack([S.insolation([S.day_of_year(d + np.timedelta64(int((s * t_in + k) * dt_days * 86400), 's'))
                                      for d in dates], lat, lon) for k in range(t_in)], axis=-1)        # (B,6,n,n,t_in)
        slots = torch.cat([p, torch.from_numpy(sol).double(), consts.double()], dim=-1)               # engine slot order
        xin = slots[..., order]                                                                        # model channel order
        p = O.unet2({k: v.double() for k, v in params.items()}, xin)
        ref.append(p)
    ref = torch.stack(ref)
    eng = RolloutEngine(model, b, n, steps, forcing_channels=cf, dtype=torch.float32, input_order=order)
    eng.set_solar(lat, lon, dt_days, t_in, t_in, start_dates=dates)
    forcing0 = torch.cat([torch.zeros(b, 6, n, n, t_in), consts], dim=-1)          # insolation slots are filled on the device
    out = eng.run(state.cuda(), forcing0.cuda())
    torch.cuda.synchronize()
    err = (out.double().cpu() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 1e-4, err
