"""CPU-only checks of the C-ABI library: it loads, exports every symbol include/dlwpcs.h declares, its host-side halo
table equals the reference-generated golden tables, and the layer classes mirror the reference's constructor behaviour."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    from dlwp_cs_b200 import build, _lib
    build.build()
    return _lib


def test_exports_match_header(lib):
    hdr = open(os.path.join(ROOT, 'include', 'dlwpcs.h')).read()
    declared = set(re.findall(r'\b(dlwpcs_[a-z0-9_]+)\s*\(', hdr))
    assert declared == set(lib.EXPORTED)
    cdll = ctypes.CDLL(lib.LIB_PATH)
    for name in declared:
        assert hasattr(cdll, name), name
    assert lib.load().dlwpcs_version() == 1


def test_desc_struct_matches_header(lib):
    hdr = open(os.path.join(ROOT, 'include', 'dlwpcs.h')).read()
    body = hdr[hdr.index('typedef struct dlwpcs_conv_desc {'):hdr.index('} dlwpcs_conv_desc;')]
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    fields = []
    for decl in re.findall(r'(?:int32_t|float)\s+([^;]+);', body):
        fields += [f.strip() for f in decl.split(',')]
    assert fields == [f[0] for f in lib.ConvDesc._fields_]


@pytest.mark.parametrize('n,p', [(4, 1), (5, 2), (8, 3), (48, 1), (24, 1), (12, 1), (6, 2)])
def test_pad_lut_vs_reference_golden(lib, n, p, golden_dir):
    g = np.load(os.path.join(golden_dir, 'pad_luts.npz'))
    np.testing.assert_array_equal(lib.pad_lut(n, p), g['cl_n%d_p%d' % (n, p)])


@pytest.mark.parametrize('n,p', [(96, 1), (48, 2), (7, 3), (3, 3), (2, 1)])
def test_pad_lut_vs_oracle(lib, n, p):
    import cs_oracle
    np.testing.assert_array_equal(lib.pad_lut(n, p), cs_oracle.pad_lut(n, p))


def test_out_edge_matches_oracle(lib):
    import cs_oracle
    for n in (5, 8, 9, 48):
        for k in (1, 2, 3, 5):
            for s in (1, 2, 3):
                for d in (1, 2):
                    for same in (False, True):
                        if not same and n < (k - 1) * d + 1:
                            continue
                        assert lib.conv_out_edge(n, k, s, d, same) == \
                            cs_oracle.conv_output_length(n, k, 'same' if same else 'valid', s, d)


def test_errors_are_reported(lib):
    with pytest.raises(lib.DlwpcsError):
        lib.pad_lut(4, 9)
    d = lib.make_desc(1, 4, 3, 3, c0=2, c1=0)
    assert lib.load().dlwpcs_packed_weight_bytes(ctypes.byref(d), 0) == -1
    assert b'c0 + c1' in lib.load().dlwpcs_last_error()


def test_no_cpu_fallback(lib):
    from dlwp_cs_b200 import CubeSpherePadding2D
    with pytest.raises(lib.DlwpcsError):
        CubeSpherePadding2D(1, data_format='channels_last')(torch.zeros(1, 6, 4, 4, 2))


def test_layer_constructors_mirror_reference():
    from dlwp_cs_b200 import CubeSphereConv2D, CubeSpherePadding2D
    with pytest.raises(ValueError):          # the reference's declared default (1, 1) is rejected by keras too
        CubeSpherePadding2D((1, 1))
    pad = CubeSpherePadding2D(2, data_format='channels_last')
    assert pad.padding == ((0, 0), (2, 2), (2, 2))
    assert pad.compute_output_shape((None, 6, 48, 48, 7)) == (None, 6, 52, 52, 7)
    assert CubeSpherePadding2D((5, 1, 1)).padding == ((0, 0), (1, 1), (1, 1))
    conv = CubeSphereConv2D(32, 3, data_format='channels_last', independent_north_pole=True, in_channels=18)
    names = [n for n, _ in conv.named_parameters()]
    assert names == ['equatorial_kernel', 'polar_kernel', 'north_pole_kernel', 'equatorial_bias', 'polar_bias',
                     'north_pole_bias']                      # custom.py:882-914 creation order
    assert tuple(conv.equatorial_kernel.shape) == (3, 3, 18, 32)
    limit = np.sqrt(6.0 / (9 * 18 + 9 * 32))
    assert float(conv.equatorial_kernel.abs().max()) <= limit and float(conv.equatorial_bias.abs().max()) == 0.0
    assert conv.compute_output_shape((4, 6, 50, 50, 18)) == (4, 6, 48, 48, 32)
    cf = CubeSphereConv2D(8, (3, 3), strides=2, padding='SAME')
    assert cf.data_format == 'channels_first' and cf.compute_output_shape((4, 5, 6, 9, 9)) == (4, 8, 6, 5, 5)
    cfg = conv.get_config()
    assert cfg['filters'] == 32 and cfg['kernel_size'] == (3, 3) and cfg['flip_north_pole'] is True
    lazy = CubeSphereConv2D(4, 3, data_format='channels_last')
    assert lazy.has_uninitialized_params()
    with pytest.raises(ValueError):
        CubeSphereConv2D(4, 3, padding='full')
    with pytest.raises(TypeError):
        CubeSphereConv2D(4, 3, bogus=1)


def test_wgrad_plan_workspace_host_logic(lib):
    """Host side of the tcgen05 wgrad plan (no GPU needed): every unet2 layer shape at the training batch gets a plan whose
    workspace holds one accumulator image per CTA; shapes outside the kernel's reach (horizontal dilation, windows whose
    accumulators exceed tensor memory) fall back to the float32 kernel's workspace; strides are refused like the
    reference-free backward contract says."""
    import ctypes
    L = lib.load()
    shapes = [(48, 18, 32, 3), (48, 32, 32, 3), (24, 32, 64, 3), (24, 64, 64, 3), (12, 64, 128, 3), (12, 128, 64, 3),
              (24, 128, 64, 3), (24, 64, 32, 3), (48, 64, 32, 3), (48, 32, 14, 1)]
    for n, ci, co, k in shapes:
        d = lib.make_desc(32, n, ci, co, (k, k), (1, 1), (1, 1), 1 if k == 3 else 0, False, True, False, True,
                          lib.ACT_NONE, 0.1, 10.0, lib.BF16, lib.BF16)
        nbytes = L.dlwpcs_wgrad_workspace_bytes(ctypes.byref(d))
        assert nbytes > 0, (n, ci, co, k)
        # 148 CTAs (or fewer) x accumulators x 128 lanes x N fp32 + bias partials: a few tens of MB at most
        assert nbytes < 64 << 20, (n, ci, co, k, nbytes)
        d32 = lib.copy_desc(d, x_dtype=lib.F32, y_dtype=lib.F32)
        assert L.dlwpcs_wgrad_workspace_bytes(ctypes.byref(d32)) > 0
    # horizontal dilation: not stackable along M -> float32 kernel's (different) workspace, still served
    dd = lib.make_desc(2, 12, 16, 16, (3, 3), (1, 1), (1, 2), 0, True, True, False, True, lib.ACT_NONE, 0.1, 10.0, lib.BF16,
                       lib.BF16)
    assert L.dlwpcs_wgrad_workspace_bytes(ctypes.byref(dd)) > 0
    ds = lib.make_desc(2, 12, 16, 16, (3, 3), (2, 2), (1, 1), 0, False, True, False, True, lib.ACT_NONE, 0.1, 10.0, lib.BF16,
                       lib.BF16)
    assert L.dlwpcs_wgrad_workspace_bytes(ctypes.byref(ds)) < 0
    assert b'stride' in L.dlwpcs_last_error()


def test_reference_input_order_and_solar_days_host_logic():
    """reference_input_order: the reference's channel packing (generators.py:880-899) as engine slots; a permutation whose
    inverse puts the prognostic channels first."""
    from dlwp_cs_b200.unet import reference_input_order
    order = reference_input_order(2, 7, 2)
    assert sorted(order) == list(range(18))
    assert order[:7] == list(range(7)) and order[7] == 14 and order[8:15] == list(range(7, 14)) and order[15] == 15
    assert order[16:] == [16, 17]


def test_unet2_keras_weight_order_roundtrip():
    """CubeSphereUNet2.get_weights / set_weights use the reference's Keras ordering (layers in creation order,
    custom.py:882-914 add_weight order inside a layer) -- host logic, parameters on the CPU."""
    from dlwp_cs_b200.unet import CubeSphereUNet2
    torch.manual_seed(0)
    a = CubeSphereUNet2(18, 14, base=8)
    w = a.get_weights()
    assert len(w) == 44
    assert w[0].shape == (3, 3, 18, 8) and w[1].shape == (3, 3, 18, 8) and w[2].shape == (8,) and w[3].shape == (8,)
    assert w[40].shape == (1, 1, 8, 14) and w[43].shape == (14,)
    assert sum(x.size for x in CubeSphereUNet2(18, 14, base=32).get_weights()) == 675932
    b = CubeSphereUNet2(18, 14, base=8)
    b.set_weights([x + 1.0 for x in w])
    for x, y in zip(w, b.get_weights()):
        np.testing.assert_array_equal(x + 1.0, y)
    with pytest.raises(ValueError):
        b.set_weights(w[:-1])
    with pytest.raises(ValueError):
        b.set_weights([w[1].transpose(3, 2, 0, 1)] + w[1:])
    c = CubeSphereUNet2(18, 14, base=8, independent_north_pole=True)
    assert len(c.get_weights()) == 66


@pytest.mark.parametrize('batch,n,cin,cout,grid', [(64, 48, 32, 32, 148), (64, 24, 64, 64, 148), (16, 96, 32, 32, 148),
                                                   (2, 8, 16, 32, 6), (1, 48, 32, 32, 37), (5, 24, 24, 32, 60),
                                                   (64, 12, 64, 64, 126), (3, 48, 64, 32, 160)])
def test_row_streamed_work_cuts(batch, n, cin, cout, grid):
    """Host logic of the row-streamed kernel's work split (dlwpcs_rs_work_cuts, no GPU): the cut points are monotone, start
    at (0, 0), end at (strips, 0), no CTA in front of the end is empty, and the cost (rows + 2 per unit) is balanced to within
    a few rows."""
    from dlwp_cs_b200 import _lib
    d = _lib.make_desc(batch, n, cin, cout, (3, 3), (1, 1), (1, 1), 1, False, True, False, True, _lib.ACT_NONE, 0.1, 10.0,
                       _lib.BF16, _lib.BF16)
    cuts = _lib.rs_work_cuts(d, grid)
    assert cuts is not None and len(cuts) == grid + 1
    wv, hout = n + 2, n
    strips = -(-4 * batch * wv // 128) + 2 * (-(-batch * wv // 128))
    snap = 1
    assert cuts[0] == (0, 0) and cuts[-1] == (strips, 0)
    lin = [s * hout + y for s, y in cuts]
    assert all(0 <= y < hout for _, y in cuts)
    assert all(b >= a for a, b in zip(lin, lin[1:]))
    for s, y in cuts:
        assert y == 0 or (snap <= y <= hout - snap), (s, y)
    sizes = [b - a for a, b in zip(lin, lin[1:])]
    done = [i for i, v in enumerate(lin) if v == lin[-1]][0]          # first cut that has reached the end
    assert all(v > 0 for v in sizes[:done])
    # cost of a CTA: its rows + 2 input rows of overhang per unit (a unit = the part of one strip it covers)
    costs = []
    for a, b in zip(lin, lin[1:]):
        units = 0 if a == b else (b - 1) // hout - a // hout + 1
        costs.append(b - a + 2 * units)
    work = [c for c in costs if c > 0]
    assert max(work) <= sum(work) / len(work) + 6
    # not served by that kernel: 5x5, 128 input channels
    d5 = _lib.make_desc(batch, n, cin, cout, (5, 5), (1, 1), (1, 1), 2, False, True, False, True, _lib.ACT_NONE, 0.1, 10.0,
                        _lib.BF16, _lib.BF16)
    assert _lib.rs_work_cuts(d5, grid) is None


def test_kernel_dispatch_rules_host_side():
    """Which bf16 forward layers the row-streamed kernel serves, and which pairs / layers can be fused -- the host-side
    rules behind dlwpcs_packed_weight_bytes / dlwpcs_conv2d_head_fusable / dlwpcs_conv2d_pool_fusable (no GPU needed)."""
    from dlwp_cs_b200 import _lib
    lib = _lib.load()
    import ctypes

    def desc(cin, cout, k=3, halo=1, n=48, batch=4, dt=_lib.BF16, m0=_lib.SRC_SAME, stride=1, dil=1):
        return _lib.make_desc(batch, n, cin, cout, (k, k), (stride, stride), (dil, dil), halo, False, True, False, True,
                              _lib.ACT_CAPPED_LEAKY_RELU, 0.1, 10.0, dt, dt, cin, m0, 0, _lib.SRC_SAME)

    def nbytes(d, tr):
        return lib.dlwpcs_packed_weight_bytes(ctypes.byref(d), tr)

    # row-streamed image: 3 kernel columns x CinP x (3 x CoutP) bf16 per face group + float32 bias[3][CoutP]
    d = desc(32, 32)
    assert nbytes(d, 0) == 3 * (3 * 32 * 96 * 2) + 3 * 32 * 4
    assert nbytes(d, 2) == 3 * (9 * 32 * 32 * 2) + 3 * 32 * 4          # classic image (chained launches)
    assert nbytes(d, 1) == nbytes(d, 2)                                 # transposed (dgrad): classic layout, cin == cout here
    assert _lib.rs_work_cuts(d, 148) is not None
    # not served by the row-streamed kernel: 128 input channels, 5x5, dilation, stride, a pooled source, a 1x1 layer
    for other in (desc(128, 64), desc(32, 32, k=5, halo=2), desc(32, 32, dil=2, halo=2), desc(32, 32, m0=_lib.SRC_POOL2, n=24),
                  desc(32, 16, k=1, halo=0)):
        assert _lib.rs_work_cuts(other, 148) is None
        assert nbytes(other, 0) == nbytes(other, 2)
    # fused 1x1 head
    head = desc(32, 16, k=1, halo=0)
    assert _lib.conv2d_head_fusable(d, head)
    assert not _lib.conv2d_head_fusable(d, desc(32, 16, k=3, halo=1))           # not a 1x1 layer
    assert not _lib.conv2d_head_fusable(d, desc(40, 16, k=1, halo=0))           # reads something else
    assert not _lib.conv2d_head_fusable(d, desc(32, 16, k=1, halo=0, n=24))     # another resolution
    assert not _lib.conv2d_head_fusable(desc(128, 64), desc(64, 16, k=1, halo=0))   # the 3x3 layer is not row-streamed
    assert _lib.conv2d_head_fusable(desc(64, 64), desc(64, 16, k=1, halo=0))
    # pooled second output: <= 32 output channels, even face edges
    assert _lib.conv2d_pool_fusable(d)
    assert not _lib.conv2d_pool_fusable(desc(64, 64))
    assert not _lib.conv2d_pool_fusable(desc(32, 32, n=9))
    assert not _lib.conv2d_pool_fusable(desc(128, 32))
