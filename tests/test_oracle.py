"""The oracle (oracle/cs_oracle.py) against the golden vectors produced from the reference's own custom.py
(tests/golden/make_golden.py) and against SURVEY.md Appendix A.  CPU only."""
import os

import numpy as np
import pytest
import torch

import cs_oracle as O
from tests.golden.cases import CONV_CASES, PAD_CASES

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

# CubeSpherePadding2D, N=4, p=1, entry "g:i,j" = source face g, row i, col j
APPENDIX_A = {
    0: """4:3,0 4:3,0 4:3,1 4:3,2 4:3,3 4:3,3
3:0,3 0:0,0 0:0,1 0:0,2 0:0,3 1:0,0
3:1,3 0:1,0 0:1,1 0:1,2 0:1,3 1:1,0
3:2,3 0:2,0 0:2,1 0:2,2 0:2,3 1:2,0
3:3,3 0:3,0 0:3,1 0:3,2 0:3,3 1:3,0
5:0,0 5:0,0 5:0,1 5:0,2 5:0,3 5:0,3""",
    1: """4:3,3 4:3,3 4:2,3 4:1,3 4:0,3 4:0,3
0:0,3 1:0,0 1:0,1 1:0,2 1:0,3 2:0,0
0:1,3 1:1,0 1:1,1 1:1,2 1:1,3 2:1,0
0:2,3 1:2,0 1:2,1 1:2,2 1:2,3 2:2,0
0:3,3 1:3,0 1:3,1 1:3,2 1:3,3 2:3,0
5:0,3 5:0,3 5:1,3 5:2,3 5:3,3 5:3,3""",
    2: """4:0,3 4:0,3 4:0,2 4:0,1 4:0,0 4:0,0
1:0,3 2:0,0 2:0,1 2:0,2 2:0,3 3:0,0
1:1,3 2:1,0 2:1,1 2:1,2 2:1,3 3:1,0
1:2,3 2:2,0 2:2,1 2:2,2 2:2,3 3:2,0
1:3,3 2:3,0 2:3,1 2:3,2 2:3,3 3:3,0
5:3,3 5:3,3 5:3,2 5:3,1 5:3,0 5:3,0""",
    3: """4:0,0 4:0,0 4:1,0 4:2,0 4:3,0 4:3,0
2:0,3 3:0,0 3:0,1 3:0,2 3:0,3 0:0,0
2:1,3 3:1,0 3:1,1 3:1,2 3:1,3 0:1,0
2:2,3 3:2,0 3:2,1 3:2,2 3:2,3 0:2,0
2:3,3 3:3,0 3:3,1 3:3,2 3:3,3 0:3,0
5:3,0 5:3,0 5:2,0 5:1,0 5:0,0 5:0,0""",
    4: """2:0,3 2:0,3 2:0,2 2:0,1 2:0,0 2:0,0
3:0,0 4:0,0 4:0,1 4:0,2 4:0,3 1:0,3
3:0,1 4:1,0 4:1,1 4:1,2 4:1,3 1:0,2
3:0,2 4:2,0 4:2,1 4:2,2 4:2,3 1:0,1
3:0,3 4:3,0 4:3,1 4:3,2 4:3,3 1:0,0
0:0,0 0:0,0 0:0,1 0:0,2 0:0,3 0:0,3""",
    5: """0:3,0 0:3,0 0:3,1 0:3,2 0:3,3 0:3,3
3:3,3 5:0,0 5:0,1 5:0,2 5:0,3 1:3,0
3:3,2 5:1,0 5:1,1 5:1,2 5:1,3 1:3,1
3:3,1 5:2,0 5:2,1 5:2,2 5:2,3 1:3,2
3:3,0 5:3,0 5:3,1 5:3,2 5:3,3 1:3,3
2:3,3 2:3,3 2:3,2 2:3,1 2:3,0 2:3,0""",
}


def _parse(txt, n=4):
    rows = []
    for line in txt.strip().splitlines():
        r = []
        for tok in line.split():
            g, ij = tok.split(':')
            i, j = ij.split(',')
            r.append(int(g) * n * n + int(i) * n + int(j))
        rows.append(r)
    return np.array(rows)


def test_pad_lut_appendix_a():
    lut = O.pad_lut(4, 1)
    for f in range(6):
        np.testing.assert_array_equal(lut[f], _parse(APPENDIX_A[f]))


def test_pad_lut_multiplicity():
    # fan-in of the backward scatter-add
    lut = O.pad_lut(4, 1)
    cnt = np.bincount(lut.ravel(), minlength=96).reshape(6, 4, 4)
    assert cnt.sum() == 216
    np.testing.assert_array_equal(cnt[0], [[4, 2, 2, 4], [2, 1, 1, 2], [2, 1, 1, 2], [4, 2, 2, 4]])
    assert cnt[1][0, 0] == 3 and cnt[3][3, 3] == 3 and cnt[4][0, 0] == 5 and cnt[5][3, 3] == 5


@pytest.mark.parametrize('n,p', PAD_CASES)
def test_pad_lut_vs_reference(n, p):
    g = np.load(os.path.join(G, 'pad_luts.npz'))
    lut = O.pad_lut(n, p)
    np.testing.assert_array_equal(lut, g['cl_n%d_p%d' % (n, p)])
    np.testing.assert_array_equal(lut, g['cf_n%d_p%d' % (n, p)])       # channels_first twin: same geometry


def test_pad_channels_first_matches_last():
    x = torch.randn(2, 6, 5, 5, 3, dtype=torch.float64)
    a = O.cube_sphere_pad(x, 2)
    b = O.cube_sphere_pad(x.permute(0, 4, 1, 2, 3).contiguous(), 2, 'channels_first').permute(0, 2, 3, 4, 1)
    assert torch.equal(a, b)


@pytest.mark.parametrize('case', CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv_vs_reference(case):
    name, kw, b, h, cin = case
    g = np.load(os.path.join(G, 'conv_cases.npz'))
    t = lambda k: torch.from_numpy(g[name + '.' + k]) if (name + '.' + k) in g.files else None
    args = dict(strides=kw.get('strides', 1), padding=kw.get('padding', 'valid'), dilation=kw.get('dilation_rate', 1),
                flip_north_pole=kw.get('flip_north_pole', True))
    x = t('x')
    y = O.cube_sphere_conv2d(x, t('equatorial_kernel'), t('polar_kernel'), t('north_pole_kernel'),
                             t('equatorial_bias'), t('polar_bias'), t('north_pole_bias'), **args)
    np.testing.assert_allclose(y.numpy(), g[name + '.y_cl'], rtol=1e-12, atol=1e-12)
    ycf = O.cube_sphere_conv2d(x.permute(0, 4, 1, 2, 3), t('equatorial_kernel'), t('polar_kernel'),
                               t('north_pole_kernel'), t('equatorial_bias'), t('polar_bias'), t('north_pole_bias'),
                               data_format='channels_first', **args)
    np.testing.assert_allclose(ycf.numpy(), g[name + '.y_cf'], rtol=1e-12, atol=1e-12)
    # oneDNN fast path used for the timed CPU baseline must agree too
    yf = O.cube_sphere_conv2d(x, t('equatorial_kernel'), t('polar_kernel'), t('north_pole_kernel'),
                              t('equatorial_bias'), t('polar_bias'), t('north_pole_bias'), exact=False, **args)
    np.testing.assert_allclose(yf.numpy(), g[name + '.y_cl'], rtol=1e-10, atol=1e-10)
    ks = kw['kernel_size'] if isinstance(kw['kernel_size'], tuple) else (kw['kernel_size'],) * 2
    s, d = args['strides'], args['dilation']
    assert y.shape[2] == O.conv_output_length(h, ks[0], args['padding'], s, d)
    assert y.shape[3] == O.conv_output_length(h, ks[1], args['padding'], s, d)


def test_cfg1_vs_reference():
    g = np.load(os.path.join(G, 'padconv_cfg1.npz'))
    x = torch.from_numpy(g['x'])
    xp = O.cube_sphere_pad(x, 1)
    np.testing.assert_array_equal(xp.numpy(), g['xp'])
    f64 = lambda k: torch.from_numpy(g[k]).double()
    y = O.cube_sphere_conv2d(xp.double(), f64('w_eq'), f64('w_pol'), None, f64('b_eq'), f64('b_pol'), None)
    np.testing.assert_allclose(y.numpy(), g['y'], rtol=1e-12, atol=1e-13)


def test_flip_equals_flipped_kernel_stride1():
    # flipping rows before and after == correlating with the kernel flipped along kh (stride 1)
    x = torch.randn(1, 6, 9, 9, 3, dtype=torch.float64)
    w, wp = torch.randn(3, 3, 3, 4, dtype=torch.float64), torch.randn(3, 3, 3, 4, dtype=torch.float64)
    a = O.cube_sphere_conv2d(x, w, wp, flip_north_pole=True, dilation=2)
    b5 = O.conv2d_tf(x[:, 5], wp.flip(0), dilation=2)
    torch.testing.assert_close(a[:, 5], b5, rtol=1e-12, atol=1e-12)


def test_pad_backward_is_scatter_add():
    # autograd of the gather == index_add over the LUT
    n, p, c = 6, 2, 3
    x = torch.randn(2, 6, n, n, c, dtype=torch.float64, requires_grad=True)
    g = torch.randn(2, 6, n + 2 * p, n + 2 * p, c, dtype=torch.float64)
    O.cube_sphere_pad(x, p).backward(g)
    lut = torch.from_numpy(O.pad_lut(n, p)).reshape(-1)
    ref = torch.zeros(2, 6 * n * n, c, dtype=torch.float64).index_add_(1, lut, g.reshape(2, -1, c))
    torch.testing.assert_close(x.grad.reshape(2, -1, c), ref, rtol=0, atol=1e-13)


def test_unet2_and_rollout_shapes():
    params = O.make_unet2_params(6, 4, base=8, dtype=torch.float64)
    x = torch.randn(1, 6, 8, 8, 4, dtype=torch.float64)
    forcing = torch.rand(1, 6, 8, 8, 2, dtype=torch.float64)
    out = O.rollout(params, x, forcing, 3)
    assert out.shape == (3, 1, 6, 8, 8, 4)
    y0 = O.unet2(params, torch.cat([x, forcing], -1))
    torch.testing.assert_close(out[0], y0)
    fast = O.rollout(params, x, forcing, 3, exact=False, host_hop=True)
    torch.testing.assert_close(out, fast, rtol=1e-9, atol=1e-9)
    # the arrangement timed as the CPU baseline (LUT gather + batched faces) is the same function
    fast2 = O.rollout_fast(params, x, forcing, 3)
    torch.testing.assert_close(out, fast2, rtol=1e-9, atol=1e-9)
    xp = torch.randn(2, 6, 6, 6, 3, dtype=torch.float64)
    assert torch.equal(O.cube_sphere_pad_fast(xp, 2), O.cube_sphere_pad(xp, 2))


def test_insolation_oracle_matches_reference_golden():
    """oracle/cs_solar.insolation against outputs of the reference's own DLWP.util.insolation (util.py:306-364), generated
    by tests/golden/make_golden_solar.py -- bit for bit, 2-d cubed-sphere lat/lon, 1-d lat/lon meshes and daily=True."""
    import os
    import numpy as np
    import cs_solar as S
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'insolation.npz'))
    assert np.array_equal(np.array([S.day_of_year(np.datetime64(x)) for x in g['dates']]), g['days'])
    for n in (4, 12):
        assert np.array_equal(S.insolation(g['days'], g['lat_%d' % n], g['lon_%d' % n]), g['sol_%d' % n])
        assert np.array_equal(S.insolation(g['days'], g['lat_%d' % n], g['lon_%d' % n], S=2.5, daily=True),
                              g['sol_daily_%d' % n])
    assert np.array_equal(S.insolation(g['days'], g['lat_1d'], g['lon_1d']), g['sol_1d'])
    sol = g['sol_12']
    assert sol.min() == 0.0 and 0.9 < sol.max() < 1.1          # night side clipped, sub-solar point ~ S / rho^2


def test_cubed_sphere_latlon_geometry():
    import numpy as np
    import cs_solar as S
    lat, lon = S.cubed_sphere_latlon(8)
    assert lat.shape == lon.shape == (6, 8, 8)
    assert np.all(lat[5] > 35) and np.all(lat[4] < -35) and np.all(np.abs(lat[:4]) < 45.0001)
    assert np.all((lon >= 0) & (lon < 360))
    # unit vectors of all cells cover the sphere evenly enough: mean position is the origin
    v = np.stack([np.cos(np.radians(lat)) * np.cos(np.radians(lon)), np.cos(np.radians(lat)) * np.sin(np.radians(lon)),
                  np.sin(np.radians(lat))])
    assert np.abs(v.reshape(3, -1).mean(axis=1)).max() < 1e-12


def test_feed_oracle_matches_reference_generator_golden():
    """oracle/cs_feed.generate against outputs of the reference's own ArrayDataGenerator.generate
    (DLWP/model/generators.py:872-984), produced by tests/golden/make_golden_feed.py -- bit for bit."""
    import os
    import numpy as np
    import cs_feed
    from tests.golden.cases import FEED_CASES
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'feed.npz'))
    for k, kw in FEED_CASES.items():
        p, t = cs_feed.generate(g['array'], g['samples_' + k], insolation_array=g['insolation_array'],
                                constants=g['constants'], **kw)
        assert np.array_equal(p, np.concatenate([g['p_' + k], g['const_' + k]], axis=-1))
        assert np.array_equal(t, g['t_' + k])
