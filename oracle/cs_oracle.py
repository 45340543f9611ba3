"""
TEST INFRASTRUCTURE ONLY -- the parity oracle.  Nothing under ``dlwp_cs_b200/`` may import this module; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs do.

A torch-CPU restatement of the one hot path of jweyn/DLWP-CS (reference mounted at /root/reference, file:line below):

  * ``cube_sphere_pad``      <- DLWP/custom.py:1082-1308  CubeSpherePadding2D.call (both data formats)
  * ``cube_sphere_conv2d``   <- DLWP/custom.py:921-1002   CubeSphereConv2D.call   (TF ``K.conv2d`` semantics restated)
  * ``conv_output_length``   <- DLWP/custom.py:1004-1030  compute_output_shape (keras conv_utils rule)
  * ``capped_leaky_relu`` / ``avg_pool_2x2`` / ``upsample_2x2`` / ``unet2`` <- Azure/train_cs.py:196-228, 277-305
  * ``rollout``              <- DLWP/model/models.py:418-460 predict_timeseries step loop (state fed back each step)
  * ``cs_network``           <- Azure/train_cs.py:233-388  basic / unet / unet3 / unet4 (same primitives, different depth)

The arithmetic itself lives in an absent third-party dependency (tensorflow==2.1.0, environment.yml:180); its published
semantics are restated: ``conv2d`` is a cross-correlation with HWIO kernels, 'valid' = no padding, 'same' = asymmetric
zero padding with the extra element at the bottom/right.

PARITY PINNING: the reference ships no tests or golden vectors.  This restatement is pinned
against the reference's *own* ``custom.py`` executed on a numpy TensorFlow shim (oracle/tf_shim.py): the fixtures under
tests/golden/ are produced by tests/golden/make_golden.py from the reference code, and tests/test_oracle.py checks this
module against them bit-for-bit (padding) / to 1e-12 (float64 conv).  What stays unpinned is TensorFlow's own conv2d
binary -- it cannot run here.

Everything is float64-capable and differentiable (torch autograd), which gives the dgrad / wgrad / halo scatter-add
reference for free.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------------------------------
# CubeSpherePadding2D  (custom.py:1082-1308)
# ----------------------------------------------------------------------------------------------------------------------

def _T(a):
    """swap the two face axes of a (B,H,W,C) block -- ``tf.transpose(..., (0, 2, 1, 3))`` at custom.py:1199"""
    return a.transpose(1, 2)


def _fh(a):
    """reverse rows (K.reverse(.., 1) on a channels_last face)"""
    return a.flip(1)


def _fw(a):
    """reverse columns (K.reverse(.., 2) on a channels_last face)"""
    return a.flip(2)


def cube_sphere_pad(x, p, data_format='channels_last'):
    """
    Six-face halo exchange.  x: (B,6,N,N,C) channels_last or (B,C,6,N,N) channels_first; returns the tensor grown by p on
    both face axes.  Two stages exactly like the reference: rows first (custom.py:1200-1251), then columns reading the
    row-padded neighbours (custom.py:1254-1306), so that corners are whatever that composition yields.
    """
    if data_format == 'channels_first':
        # custom.py:1085-1196 is the same geometry with the channel axis in front; restate by moving the axis.
        y = cube_sphere_pad(x.permute(0, 2, 3, 4, 1), p, 'channels_last')
        return y.permute(0, 4, 1, 2, 3)
    if data_format != 'channels_last':
        raise ValueError('Unknown data_format: %r' % (data_format,))
    if x.dim() != 5 or x.shape[1] != 6 or x.shape[2] != x.shape[3]:
        raise ValueError('expected (B,6,N,N,C), got %r' % (tuple(x.shape),))
    if p == 0:
        return x
    f = [x[:, i] for i in range(6)]
    # ---- stage 1: top / bottom halos (custom.py:1203-1249)
    top = [
        f[4][:, -p:, :],                         # face 0 <- south pole, last rows           (1205)
        _T(f[4].flip(1)[:, :, -p:]),             # face 1 <- south pole, last cols, rows reversed, transposed (1213)
        _fh(_fw(f[4][:, :p, :])),                # face 2 <- south pole, first rows, rotated 180 (1221)
        _T(_fw(f[4][:, :, :p])),                 # face 3 <- south pole, first cols, reversed, transposed (1229)
        _fh(_fw(f[2][:, :p, :])),                # face 4 <- face 2 first rows rotated 180    (1237)
        f[0][:, -p:, :],                         # face 5 <- face 0 last rows                 (1245)
    ]
    bottom = [
        f[5][:, :p, :],                          # (1207)
        _T(_fw(f[5][:, :, -p:])),                # (1215)
        _fh(_fw(f[5][:, -p:, :])),               # (1223)
        _T(f[5].flip(1)[:, :, :p]),              # (1231)
        f[0][:, :p, :],                          # (1239)
        _fh(_fw(f[2][:, -p:, :])),               # (1247)
    ]
    s1 = [torch.cat([top[i], f[i], bottom[i]], dim=1) for i in range(6)]       # each (B, N+2p, N, C)
    # ---- stage 2: left / right halos from the row-padded faces (custom.py:1257-1303)
    out = [None] * 6
    for i in range(4):                                                         # equatorial belt is periodic (1259-1286)
        out[i] = torch.cat([s1[(i + 3) % 4][:, :, -p:], s1[i], s1[(i + 1) % 4][:, :, :p]], dim=2)
    # polar faces read rows p:2p / -2p:-p of the FULLY padded faces 3 and 1 (1291-1301)
    out[4] = torch.cat([_T(_fh(out[3][:, p:2 * p, :])), s1[4], _T(_fw(out[1][:, p:2 * p, :]))], dim=2)
    n2 = out[3].shape[1]
    out[5] = torch.cat([_T(_fw(out[3][:, n2 - 2 * p:n2 - p, :])), s1[5], _T(_fh(out[1][:, n2 - 2 * p:n2 - p, :]))], dim=2)
    return torch.stack(out, dim=1)


def pad_lut(n, p):
    """
    Index form of the halo exchange: int64 array (6, n+2p, n+2p) whose entry is the flat (face*n*n + i*n + j) index of the
    source element.  Obtained by pushing an index-encoded tensor through ``cube_sphere_pad``.
    """
    idx = torch.arange(6 * n * n, dtype=torch.float64).reshape(1, 6, n, n, 1)
    return cube_sphere_pad(idx, p).reshape(6, n + 2 * p, n + 2 * p).to(torch.int64).numpy()


# ----------------------------------------------------------------------------------------------------------------------
# CubeSphereConv2D  (custom.py:921-1002)
# ----------------------------------------------------------------------------------------------------------------------

def _pair(v):
    return (int(v), int(v)) if isinstance(v, (int, np.integer)) else (int(v[0]), int(v[1]))


def conv_output_length(n, k, padding, stride, dilation=1):
    """keras conv_utils.conv_output_length as used at custom.py:1010-1015."""
    k_eff = k + (k - 1) * (dilation - 1)
    if padding == 'same':
        out = n
    elif padding == 'valid':
        out = n - k_eff + 1
    else:
        raise ValueError(padding)
    return (out + stride - 1) // stride


def _same_pads(n, k_eff, s):
    total = max((-(-n // s) - 1) * s + k_eff - n, 0)
    return total // 2, total - total // 2


def conv2d_tf(x, w, strides=1, padding='valid', dilation=1, exact=True):
    """
    TensorFlow ``K.conv2d`` on one channels_last face batch.  x: (B,H,W,Cin), w: HWIO (kh,kw,Cin,Cout).
    exact=True: explicit window gather + one matmul in the input dtype (use float64 for the oracle);
    exact=False: torch/oneDNN ``F.conv2d`` -- only for the timed CPU baseline.
    """
    sh, sw = _pair(strides)
    dh, dw = _pair(dilation)
    kh, kw, cin, cout = w.shape
    if x.shape[-1] != cin:
        raise ValueError('input depth %d != kernel depth %d' % (x.shape[-1], cin))
    padding = padding.lower()
    if padding == 'same':
        pt, pb = _same_pads(x.shape[1], (kh - 1) * dh + 1, sh)
        pl, pr = _same_pads(x.shape[2], (kw - 1) * dw + 1, sw)
        x = F.pad(x, (0, 0, pl, pr, pt, pb))
    elif padding != 'valid':
        raise ValueError('Unknown padding: %r' % (padding,))
    b, h, wd, _ = x.shape
    ho = (h - ((kh - 1) * dh + 1)) // sh + 1
    wo = (wd - ((kw - 1) * dw + 1)) // sw + 1
    if not exact:
        y = F.conv2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), stride=(sh, sw), dilation=(dh, dw))
        return y.permute(0, 2, 3, 1)
    cols = []
    for u in range(kh):
        for v in range(kw):
            cols.append(x[:, u * dh: u * dh + (ho - 1) * sh + 1: sh, v * dw: v * dw + (wo - 1) * sw + 1: sw, :])
    a = torch.cat(cols, dim=-1)                                   # (B,ho,wo,kh*kw*Cin), tap-major then channel
    return a.reshape(-1, kh * kw * cin).matmul(w.reshape(kh * kw * cin, cout)).reshape(b, ho, wo, cout)


def cube_sphere_conv2d(x, w_eq, w_pol, w_np=None, b_eq=None, b_pol=None, b_np=None, strides=1, padding='valid',
                       dilation=1, flip_north_pole=True, data_format='channels_last', exact=True):
    """
    Per-face convolution with the equatorial kernel on faces 0-3 (custom.py:926-943), the polar kernel on face 4
    (946-962) and -- with the rows reversed before and after when flip_north_pole -- the polar or independent north-pole
    kernel on face 5 (965-996); faces re-stacked (998).  Activation (1000-1002) is left to the caller.
    """
    if data_format == 'channels_first':
        y = cube_sphere_conv2d(x.permute(0, 2, 3, 4, 1), w_eq, w_pol, w_np, b_eq, b_pol, b_np, strides, padding,
                               dilation, flip_north_pole, 'channels_last', exact)
        return y.permute(0, 4, 1, 2, 3)
    outs = []
    for f in range(4):
        y = conv2d_tf(x[:, f], w_eq, strides, padding, dilation, exact)
        outs.append(y + b_eq if b_eq is not None else y)
    y = conv2d_tf(x[:, 4], w_pol, strides, padding, dilation, exact)
    outs.append(y + b_pol if b_pol is not None else y)
    w5 = w_np if w_np is not None else w_pol
    b5 = b_np if w_np is not None else b_pol
    x5 = x[:, 5].flip(1) if flip_north_pole else x[:, 5]
    y = conv2d_tf(x5, w5, strides, padding, dilation, exact)
    if b5 is not None:
        y = y + b5
    outs.append(y.flip(1) if flip_north_pole else y)
    return torch.stack(outs, dim=1)


# ----------------------------------------------------------------------------------------------------------------------
# the steps either side of the conv in the Weyn-2020 U-Net (Azure/train_cs.py:196-228)
# ----------------------------------------------------------------------------------------------------------------------

def capped_leaky_relu(x, negative_slope=0.1, max_value=10.0):
    """keras ReLU(negative_slope=0.1, max_value=10.) at train_cs.py:199: x<0 -> slope*x; 0<=x<=max -> x; else max."""
    return torch.where(x < 0, negative_slope * x, torch.clamp(x, max=max_value))


def avg_pool_2x2(x):
    """AveragePooling3D((1,2,2)) channels_last, train_cs.py:197."""
    b, f, h, w, c = x.shape
    return x.reshape(b, f, h // 2, 2, w // 2, 2, c).mean(dim=(3, 5))


def upsample_2x2(x):
    """UpSampling3D((1,2,2)) nearest neighbour, train_cs.py:198."""
    return x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)


UNET2_LAYERS = ('conv_2d_1', 'conv_2d_1_2', 'conv_2d_2', 'conv_2d_2_2', 'conv_2d_5_2', 'conv_2d_5', 'conv_2d_6_2',
                'conv_2d_6', 'conv_2d_7', 'conv_2d_7_2', 'conv_2d_8')


def unet2_shapes(cin, cout, base=32):
    """(name, kernel, Cin, Cout) for every CubeSphereConv2D of ``unet2`` (train_cs.py:209-228, 277-305)."""
    b = base
    return [('conv_2d_1', 3, cin, b), ('conv_2d_1_2', 3, b, b), ('conv_2d_2', 3, b, 2 * b), ('conv_2d_2_2', 3, 2 * b, 2 * b),
            ('conv_2d_5_2', 3, 2 * b, 4 * b), ('conv_2d_5', 3, 4 * b, 2 * b), ('conv_2d_6_2', 3, 4 * b, 2 * b),
            ('conv_2d_6', 3, 2 * b, b), ('conv_2d_7', 3, 2 * b, b), ('conv_2d_7_2', 3, b, b), ('conv_2d_8', 1, b, cout)]


def make_unet2_params(cin, cout, base=32, seed=1, dtype=torch.float32, bias_scale=0.05):
    """glorot-uniform kernels (custom.py:835 default) and small random biases, deterministic in ``seed``."""
    g = torch.Generator().manual_seed(seed)
    params = {}
    for name, k, ci, co in unet2_shapes(cin, cout, base):
        limit = math.sqrt(6.0 / (k * k * ci + k * k * co))
        for kind in ('equatorial', 'polar'):
            params['%s.%s_kernel' % (name, kind)] = ((torch.rand(k, k, ci, co, generator=g, dtype=torch.float64) * 2 - 1)
                                                     * limit).to(dtype)
            params['%s.%s_bias' % (name, kind)] = ((torch.rand(co, generator=g, dtype=torch.float64) * 2 - 1)
                                                   * bias_scale).to(dtype)
    return params


def unet2(params, x, exact=True):
    """The ``unet2`` function of Azure/train_cs.py:277-305, channels_last, flip_north_pole=True, shared polar kernel."""
    def cs(name, t, pad=1, act=True):
        if pad:
            t = cube_sphere_pad(t, pad)
        t = cube_sphere_conv2d(t, params[name + '.equatorial_kernel'], params[name + '.polar_kernel'], None,
                               params[name + '.equatorial_bias'], params[name + '.polar_bias'], None, exact=exact)
        return capped_leaky_relu(t) if act else t
    x0 = cs('conv_2d_1_2', cs('conv_2d_1', x))
    x1 = cs('conv_2d_2_2', cs('conv_2d_2', avg_pool_2x2(x0)))
    x2 = cs('conv_2d_5', cs('conv_2d_5_2', avg_pool_2x2(x1)))
    t = torch.cat([upsample_2x2(x2), x1], dim=-1)                                 # train_cs.py:292-293
    t = cs('conv_2d_6', cs('conv_2d_6_2', t))
    t = torch.cat([upsample_2x2(t), x0], dim=-1)                                  # train_cs.py:298-299
    t = cs('conv_2d_7_2', cs('conv_2d_7', t))
    return cs('conv_2d_8', t, pad=0, act=False)                                  # train_cs.py:304, 1x1 'output'


# ----------------------------------------------------------------------------------------------------------------------
# the other architectures of Azure/train_cs.py (233-251 basic, 254-274 unet, 308-346 unet3, 349-388 unet4), restated
# statement by statement like ``unet2`` above.  Filter counts: train_cs.py:208-228 with skip_connections = 'unet' in name.
# ----------------------------------------------------------------------------------------------------------------------

def arch_shapes(arch, cin, cout, base=32):
    """(name, kernel, Cin, Cout) of every CubeSphereConv2D the architecture uses, in execution order."""
    b = base
    if arch == 'unet2':
        return unet2_shapes(cin, cout, base)
    if arch == 'basic':       # no skip connections: conv_2d_6 has 2b filters (train_cs.py:222)
        return [('conv_2d_1', 3, cin, b), ('conv_2d_2', 3, b, 2 * b), ('conv_2d_3', 3, 2 * b, 4 * b),
                ('conv_2d_6', 3, 4 * b, 2 * b), ('conv_2d_7', 3, 2 * b, b), ('conv_2d_7_2', 3, b, b), ('conv_2d_8', 1, b, cout)]
    if arch == 'unet':
        return [('conv_2d_1', 3, cin, b), ('conv_2d_2', 3, b, 2 * b), ('conv_2d_3', 3, 2 * b, 4 * b),
                ('conv_2d_6', 3, 6 * b, b), ('conv_2d_7', 3, 2 * b, b), ('conv_2d_7_2', 3, b, b), ('conv_2d_8', 1, b, cout)]
    if arch == 'unet3':
        return [('conv_2d_1', 3, cin, b), ('conv_2d_1_2', 3, b, b), ('conv_2d_1_3', 3, b, b),
                ('conv_2d_2', 3, b, 2 * b), ('conv_2d_2_2', 3, 2 * b, 2 * b), ('conv_2d_2_3', 3, 2 * b, 2 * b),
                ('conv_2d_5_3', 3, 2 * b, 4 * b), ('conv_2d_5_2', 3, 4 * b, 4 * b), ('conv_2d_5', 3, 4 * b, 2 * b),
                ('conv_2d_6_3', 3, 4 * b, 2 * b), ('conv_2d_6_2', 3, 2 * b, 2 * b), ('conv_2d_6', 3, 2 * b, b),
                ('conv_2d_7', 3, 2 * b, b), ('conv_2d_7_2', 3, b, b), ('conv_2d_7_3', 3, b, b), ('conv_2d_8', 1, b, cout)]
    if arch == 'unet4':
        return [('conv_2d_1', 3, cin, b), ('conv_2d_1_2', 3, b, b), ('conv_2d_2', 3, b, 2 * b), ('conv_2d_2_2', 3, 2 * b, 2 * b),
                ('conv_2d_3_2', 3, 2 * b, 4 * b), ('conv_2d_3', 3, 4 * b, 4 * b), ('conv_2d_4_2', 3, 4 * b, 8 * b),
                ('conv_2d_4', 3, 8 * b, 4 * b), ('conv_2d_5_2', 3, 8 * b, 4 * b), ('conv_2d_5', 3, 4 * b, 2 * b),
                ('conv_2d_6_2', 3, 4 * b, 2 * b), ('conv_2d_6', 3, 2 * b, b), ('conv_2d_7', 3, 2 * b, b),
                ('conv_2d_7_2', 3, b, b), ('conv_2d_8', 1, b, cout)]
    raise ValueError(arch)


def make_arch_params(arch, cin, cout, base=32, seed=1, dtype=torch.float32, bias_scale=0.05):
    g = torch.Generator().manual_seed(seed)
    params = {}
    for name, k, ci, co in arch_shapes(arch, cin, cout, base):
        limit = math.sqrt(6.0 / (k * k * ci + k * k * co))
        for kind in ('equatorial', 'polar'):
            params['%s.%s_kernel' % (name, kind)] = ((torch.rand(k, k, ci, co, generator=g, dtype=torch.float64) * 2 - 1)
                                                     * limit).to(dtype)
            params['%s.%s_bias' % (name, kind)] = ((torch.rand(co, generator=g, dtype=torch.float64) * 2 - 1)
                                                   * bias_scale).to(dtype)
    return params


def cs_network(arch, params, x, exact=True):
    """``basic`` / ``unet`` / ``unet2`` / ``unet3`` / ``unet4`` of Azure/train_cs.py, channels_last."""
    def cs(name, t, pad=1, act=True):
        if pad:
            t = cube_sphere_pad(t, pad)
        t = cube_sphere_conv2d(t, params[name + '.equatorial_kernel'], params[name + '.polar_kernel'], None,
                               params[name + '.equatorial_bias'], params[name + '.polar_bias'], None, exact=exact)
        return capped_leaky_relu(t) if act else t
    pool, up = avg_pool_2x2, upsample_2x2
    cat = lambda a, b: torch.cat([a, b], dim=-1)
    if arch == 'unet2':
        return unet2(params, x, exact)
    if arch == 'basic':                                                         # train_cs.py:233-251
        t = cs('conv_2d_1', x)
        t = cs('conv_2d_2', pool(t))
        t = cs('conv_2d_3', pool(t))
        t = cs('conv_2d_6', up(t))
        t = cs('conv_2d_7', up(t))
        t = cs('conv_2d_7_2', t)
        return cs('conv_2d_8', t, pad=0, act=False)
    if arch == 'unet':                                                          # train_cs.py:254-274
        x0 = cs('conv_2d_1', x)
        x1 = cs('conv_2d_2', pool(x0))
        x2 = cs('conv_2d_3', pool(x1))
        t = cs('conv_2d_6', cat(up(x2), x1))
        t = cs('conv_2d_7', cat(up(t), x0))
        t = cs('conv_2d_7_2', t)
        return cs('conv_2d_8', t, pad=0, act=False)
    if arch == 'unet3':                                                         # train_cs.py:308-346
        x0 = cs('conv_2d_1_3', cs('conv_2d_1_2', cs('conv_2d_1', x)))
        x1 = cs('conv_2d_2_3', cs('conv_2d_2_2', cs('conv_2d_2', pool(x0))))
        x2 = cs('conv_2d_5', cs('conv_2d_5_2', cs('conv_2d_5_3', pool(x1))))
        t = cs('conv_2d_6', cs('conv_2d_6_2', cs('conv_2d_6_3', cat(up(x2), x1))))
        t = cs('conv_2d_7_3', cs('conv_2d_7_2', cs('conv_2d_7', cat(up(t), x0))))
        return cs('conv_2d_8', t, pad=0, act=False)
    if arch == 'unet4':                                                         # train_cs.py:349-388
        x0 = cs('conv_2d_1_2', cs('conv_2d_1', x))
        x1 = cs('conv_2d_2_2', cs('conv_2d_2', pool(x0)))
        x2 = cs('conv_2d_3', cs('conv_2d_3_2', pool(x1)))
        x3 = cs('conv_2d_4', cs('conv_2d_4_2', pool(x2)))
        t = cs('conv_2d_5', cs('conv_2d_5_2', cat(up(x3), x2)))
        t = cs('conv_2d_6', cs('conv_2d_6_2', cat(up(t), x1)))
        t = cs('conv_2d_7_2', cs('conv_2d_7', cat(up(t), x0)))
        return cs('conv_2d_8', t, pad=0, act=False)
    raise ValueError(arch)


def rollout(params, state, forcing, steps, exact=True, host_hop=False):
    """
    Autoregressive forecast, the loop of DLWP/model/models.py:446-454: every step the network output replaces the
    prognostic channels of the input.  state: (B,6,N,N,Cp) prognostic channels; forcing: (B,6,N,N,Cf) channels appended
    unchanged every step (insolation + constants stand-in: Azure/train_cs.py:396-407) or None.
    Returns (steps,B,6,N,N,Cp).  host_hop=True round-trips through numpy each step like ``keras.Model.predict`` does.
    """
    outs = []
    p = state
    for _ in range(steps):
        xin = p if forcing is None else torch.cat([p, forcing], dim=-1)
        p = unet2(params, xin, exact=exact)
        if host_hop:
            p = torch.from_numpy(np.array(p.numpy(), copy=True))
        outs.append(p)
    return torch.stack(outs, dim=0)


# ----------------------------------------------------------------------------------------------------------------------
# Timed CPU baseline: the same arithmetic arranged the way a CPU likes it (one LUT gather for the halo exchange instead
# of ~70 strided slices/concats, the four equatorial faces batched into one oneDNN convolution).  Checked against the
# literal restatement above in tests/test_oracle.py; used only by bench.py's cpu_baseline / --impl reference legs.
# ----------------------------------------------------------------------------------------------------------------------
_LUT_CACHE = {}


def cube_sphere_pad_fast(x, p):
    b, _, n, _, c = x.shape
    key = (n, p)
    if key not in _LUT_CACHE:
        _LUT_CACHE[key] = torch.from_numpy(pad_lut(n, p).reshape(-1))
    h = n + 2 * p
    return x.reshape(b, 6 * n * n, c).index_select(1, _LUT_CACHE[key]).reshape(b, 6, h, h, c)


def _conv_nhwc(x, w):
    # x (B',H,W,C) contiguous == NCHW tensor in channels_last memory format; w HWIO
    y = F.conv2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1))
    return y.permute(0, 2, 3, 1)


def cube_sphere_conv2d_fast(x, w_eq, w_pol, b_eq, b_pol):
    """stride 1 / 'valid' / shared polar kernel / flip_north_pole=True only (the U-Net's configuration)."""
    b, _, h, w, c = x.shape
    y_eq = _conv_nhwc(x[:, :4].reshape(b * 4, h, w, c), w_eq) + b_eq
    ho, wo, co = y_eq.shape[1:]
    y4 = _conv_nhwc(x[:, 4], w_pol) + b_pol
    y5 = (_conv_nhwc(x[:, 5].flip(1), w_pol) + b_pol).flip(1)
    return torch.cat([y_eq.reshape(b, 4, ho, wo, co), y4.unsqueeze(1), y5.unsqueeze(1)], dim=1)


def unet2_fast(params, x):
    def cs(name, t, pad=1, act=True):
        if pad:
            t = cube_sphere_pad_fast(t, pad)
        t = cube_sphere_conv2d_fast(t, params[name + '.equatorial_kernel'], params[name + '.polar_kernel'],
                                    params[name + '.equatorial_bias'], params[name + '.polar_bias'])
        return capped_leaky_relu(t) if act else t
    x0 = cs('conv_2d_1_2', cs('conv_2d_1', x))
    x1 = cs('conv_2d_2_2', cs('conv_2d_2', avg_pool_2x2(x0)))
    x2 = cs('conv_2d_5', cs('conv_2d_5_2', avg_pool_2x2(x1)))
    t = torch.cat([upsample_2x2(x2), x1], dim=-1)
    t = cs('conv_2d_6', cs('conv_2d_6_2', t))
    t = torch.cat([upsample_2x2(t), x0], dim=-1)
    t = cs('conv_2d_7_2', cs('conv_2d_7', t))
    return cs('conv_2d_8', t, pad=0, act=False)


def rollout_fast(params, state, forcing, steps, host_hop=True):
    """``rollout`` on the fast arrangement, with the numpy hop per step of keras ``Model.predict`` (models.py:446-454)."""
    outs = []
    p = state
    for _ in range(steps):
        xin = p if forcing is None else torch.cat([p, forcing], dim=-1)
        p = unet2_fast(params, xin)
        if host_hop:
            p = torch.from_numpy(np.array(p.numpy(), copy=True))
        outs.append(p)
    return torch.stack(outs, dim=0)
