"""Differentiable operators over libdlwpcs: the halo exchange and the (optionally halo-fused) per-face convolution.

Forward semantics follow the reference's DLWP/custom.py:1082-1308 (CubeSpherePadding2D.call) and 921-1002
(CubeSphereConv2D.call); backward is what TensorFlow autodiff derives from those: dgrad with the halo
scatter-add, wgrad reduced over the faces sharing a kernel.  All tensors here are channels_last (B,6,H,W,C).
"""
import torch

from . import _lib

ACTIVATIONS = {None: (_lib.ACT_NONE, 0.0, 0.0), 'linear': (_lib.ACT_NONE, 0.0, 0.0),
               'relu': (_lib.ACT_CAPPED_LEAKY_RELU, 0.0, float('inf'))}


def resolve_activation(activation):
    """-> (act code, slope, max) for activations the kernels fuse, or None if it has to be applied separately.
    ('capped_leaky_relu', slope, max) is keras ReLU(negative_slope, max_value) of Azure/train_cs.py:199."""
    if isinstance(activation, (tuple, list)) and len(activation) == 3 and activation[0] == 'capped_leaky_relu':
        return (_lib.ACT_CAPPED_LEAKY_RELU, float(activation[1]), float(activation[2]))
    try:
        return ACTIVATIONS.get(activation)
    except TypeError:
        return None


def _check_cl(x):
    if x.dim() != 5 or x.shape[1] != 6:
        raise ValueError('expected a (batch, 6, height, width, channels) tensor, got %r' % (tuple(x.shape),))
    if x.shape[2] != x.shape[3]:
        raise ValueError('cubed-sphere faces must be square, got %dx%d' % (x.shape[2], x.shape[3]))


class _CubeSpherePad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p):
        ctx.p = p
        return _lib.pad_fwd(x.contiguous(), p)

    @staticmethod
    def backward(ctx, dy):
        return _lib.pad_bwd(dy.contiguous(), ctx.p), None


def cube_sphere_pad(x, p):
    """CubeSpherePadding2D.call on a channels_last tensor (float32 or bfloat16)."""
    _check_cl(x)
    if p == 0:
        return x
    if p < 0 or p > x.shape[2]:
        raise ValueError('padding %d out of range for face edge %d' % (p, x.shape[2]))
    return _CubeSpherePad.apply(x, int(p))


# Gradient sink (set by the data-parallel trainer): an object with ``views(params)`` -> the six destination tensors for the
# weight / bias gradients of a layer whose parameter tensors are `params` (or None if it does not own them) and
# ``delivered(params)``.  With a sink the weight gradients are written straight into the trainer's flat gradient buffer by
# the wgrad kernel's second pass -- no per-parameter ``.grad`` tensors, no accumulation or gather kernels -- and the trainer
# learns when a layer's gradients are complete (to start their all-reduce while backward continues).
_GRAD_SINK = None


def set_grad_sink(sink):
    global _GRAD_SINK
    prev, _GRAD_SINK = _GRAD_SINK, sink
    return prev


# Pre-packed weight images (set by the data-parallel trainer for the duration of a step): data_ptr of a layer's equatorial
# kernel -> (layout key, packed, packed_t).  The trainer packs every layer of the model in ONE launch at the start of the step
# (dlwpcs_pack_weights_batch) instead of one or two launches per layer inside the layers' forward.
_PREPACKED = {}


def _pack_key(d):
    return (d.cin, d.cout, d.kh, d.kw, d.stride_h, d.stride_w, d.dil_h, d.dil_w, d.same, d.flip_north_pole,
            d.independent_north_pole, d.use_bias, d.x_dtype, d.y_dtype)


def prepack_model(model, dtype=torch.bfloat16, state=None):
    """Pack the forward and transposed (dgrad) bf16 weight images of every CubeSphereConv2D of `model` in one launch and
    publish them for the layers' forward (until ``clear_prepacked``).  state: what a previous call returned (buffers are
    reused -- CUDA-graph friendly).  The first layer of a layer program needs no transposed image (its input has no
    gradient)."""
    from .custom import CubeSphereConv2D
    layers = [m for m in model.modules() if isinstance(m, CubeSphereConv2D) and not m.has_uninitialized_params()]
    program = getattr(model, 'program', None)
    first = getattr(model, program[0]['name']) if program else None
    entries, keys = [], []
    for layer in layers:
        d, ws = layer.pack_descriptor(dtype)
        entries.append((d, ws, True, layer is not first))
        keys.append(_pack_key(d))
    res = _lib.pack_weights_batch(entries, out=state)
    for layer, key, (packed, packed_t) in zip(layers, keys, res):
        _PREPACKED[layer.equatorial_kernel.data_ptr()] = (key, packed, packed_t)
    return res


def clear_prepacked():
    _PREPACKED.clear()


def _pad_w(w, pad_in, pad_out):
    return w if (w is None or not (pad_in or pad_out)) else torch.nn.functional.pad(w, (0, pad_out, 0, pad_in))


def _pad_b(b, pad_out):
    return b if (b is None or not pad_out) else torch.nn.functional.pad(b, (0, pad_out))


class _CubeSphereConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w_eq, w_pol, w_np, b_eq, b_pol, b_np, cfg):
        x = x.contiguous()
        b, _, n, _, cin = x.shape
        # zero-weight channel padding of the bf16 path (multiples of 8 channels for 16-byte gathers / stores) happens in
        # here, on the packed copies only: x arrives padded, the parameters keep their shape
        pad_in, pad_out = cfg.get('pad_in', 0), cfg.get('pad_out', 0)
        kh, kw, wcin, cout = w_eq.shape
        if wcin + pad_in != cin:
            raise ValueError('input has %d channels, kernel expects %d' % (cin, wcin + pad_in))
        cout_p = cout + pad_out
        d = _lib.make_desc(b, n, cin, cout_p, (kh, kw), cfg['strides'], cfg['dilation'], cfg['halo'], cfg['same'],
                           cfg['flip_north_pole'], w_np is not None, b_eq is not None, cfg['act'][0], cfg['act'][1],
                           cfg['act'][2], _lib.dtype_code(x.dtype), _lib.dtype_code(cfg.get('out_dtype', x.dtype)))
        ctx.packed_t = None
        pre = _PREPACKED.get(w_eq.data_ptr()) if d.x_dtype == _lib.BF16 else None
        if pre is not None and pre[0] == _pack_key(d) and (pre[2] is not None or not ctx.needs_input_grad[0]):
            packed, ctx.packed_t = pre[1], pre[2]
        elif d.x_dtype == _lib.BF16:
            # one launch packs the forward image and -- when the input needs a gradient -- the transposed one for dgrad;
            # the kernels are zero-extended to the padded channel counts inside the pack kernel
            packed, ctx.packed_t = _lib.pack_weights2(d, w_eq, w_pol, w_np, b_eq, b_pol, b_np, forward=True,
                                                      transposed=bool(ctx.needs_input_grad[0]))
        else:
            packed = _lib.pack_weights(d, _pad_w(w_eq, pad_in, pad_out), _pad_w(w_pol, pad_in, pad_out),
                                       _pad_w(w_np, pad_in, pad_out), _pad_b(b_eq, pad_out), _pad_b(b_pol, pad_out),
                                       _pad_b(b_np, pad_out))
        y = _lib.conv2d_fwd(d, x, None, packed)
        ctx.d = d
        ctx.pads = (pad_in, pad_out)
        ctx.in_act = cfg.get('in_act')
        ctx.dy_premasked = bool(cfg.get('dy_premasked', False))
        ctx.save_for_backward(x, y, w_eq, w_pol, w_np, b_eq, b_pol, b_np)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, w_eq, w_pol, w_np, b_eq, b_pol, b_np = ctx.saved_tensors
        d = ctx.d
        pad_in, pad_out = ctx.pads
        if d.x_dtype != d.y_dtype:
            raise _lib.DlwpcsError('backward needs the output dtype to equal the input dtype')
        dy = dy.contiguous()
        if ctx.dy_premasked:
            # the consumer's dgrad already multiplied by this layer's activation derivative (dlwpcs_conv2d_dgrad_act)
            d = _lib.copy_desc(d, act=_lib.ACT_NONE)
            y = None
        elif d.act != _lib.ACT_NONE and d.x_dtype == _lib.BF16:
            # bf16 path: the activation derivative is applied once (dgrad and wgrad both need dy * act'(y)); both kernels
            # then stream the product with plain asynchronous 16-byte copies instead of masking in registers
            dy = _lib.act_bwd(dy, y, d.act, d.act_slope, d.act_max)
            d = _lib.copy_desc(d, act=_lib.ACT_NONE)
            y = None
        dx = None
        if ctx.needs_input_grad[0]:
            packed_t = ctx.packed_t
            if packed_t is None:
                packed_t = _lib.pack_weights(d, _pad_w(w_eq, pad_in, pad_out), _pad_w(w_pol, pad_in, pad_out),
                                             _pad_w(w_np, pad_in, pad_out), transposed=True)
            # in_act: x is the output of an activated layer that nothing else reads -- hand that layer
            # dL/d(pre-activation) directly (the mask is applied inside the halo scatter-add)
            dx = _lib.conv2d_dgrad(d, dy, y, packed_t, x_in=x, in_act=ctx.in_act)
        grads = [None] * 6
        if any(ctx.needs_input_grad[1:7]):
            params = (w_eq, w_pol, w_np, b_eq, b_pol, b_np)
            sink = _GRAD_SINK
            views = sink.views(params) if sink is not None else None
            if views is not None and not (pad_in or pad_out):
                _lib.conv2d_wgrad(d, x, dy, y, outs=views)          # second pass writes the flat buffer directly
                sink.delivered(params)
            else:
                got = _lib.conv2d_wgrad(d, x, dy, y)
                if pad_in or pad_out:                                # drop the pad channels' (exactly zero) gradients
                    cin0, cout0 = w_eq.shape[2], w_eq.shape[3]
                    got = tuple(None if g is None else (g[:, :, :cin0, :cout0] if g.dim() == 4 else g[:cout0]) for g in got)
                if views is not None:
                    for v, g in zip(views, got):
                        if v is not None:
                            v.copy_(g)
                    sink.delivered(params)
                else:
                    grads = [g if (g is None or g.is_contiguous()) else g.contiguous() for g in got]
        return (dx, *grads, None)


def cube_sphere_conv2d(x, equatorial_kernel, polar_kernel, north_pole_kernel=None, equatorial_bias=None,
                       polar_bias=None, north_pole_bias=None, strides=(1, 1), padding='valid', dilation_rate=(1, 1),
                       flip_north_pole=True, halo=0, activation=None, out_dtype=None, pad_in=0, pad_out=0, in_act=None,
                       dy_premasked=False):
    """
    CubeSphereConv2D.call on a channels_last tensor.  `halo` > 0 additionally performs CubeSpherePadding2D(halo) inside
    the kernel's load stage (x is then the un-padded tensor).  `activation`: None / 'linear' / 'relu' /
    ('capped_leaky_relu', slope, max) are fused into the epilogue.
    """
    _check_cl(x)
    act = resolve_activation(activation)
    if act is None:
        raise ValueError('activation %r cannot be fused; apply it to the output instead' % (activation,))
    padding = padding.lower()
    if padding not in ('valid', 'same'):
        raise ValueError('The `padding` argument must be one of "valid", "same". Received: %s' % padding)
    cfg = dict(strides=tuple(strides), dilation=tuple(dilation_rate), halo=int(halo), same=padding == 'same',
               flip_north_pole=bool(flip_north_pole), act=act)
    if out_dtype is not None:
        cfg['out_dtype'] = out_dtype
    # pad_in / pad_out: x carries pad_in extra zero channels and the result pad_out extra (exactly zero) channels; the
    # kernels / biases are given un-padded and padded with zeros on their packed copies only
    cfg['pad_in'], cfg['pad_out'] = int(pad_in), int(pad_out)
    # in_act / dy_premasked: the producer / consumer halves of the fused activation backward (used by CubeSphereCNN for
    # directly chained layers): the consumer (in_act = the producer's activation) returns dL/d(pre-activation of the
    # producer) as its input gradient, the producer (dy_premasked) skips its own derivative
    if in_act is not None:
        cfg['in_act'] = resolve_activation(in_act)
        if pad_in:
            raise ValueError('in_act cannot be combined with channel padding of the input')
    cfg['dy_premasked'] = bool(dy_premasked)
    return _CubeSphereConv.apply(x, equatorial_kernel, polar_kernel, north_pole_kernel, equatorial_bias, polar_bias,
                                 north_pole_bias, cfg)


def cube_sphere_conv2d_tc32(x, equatorial_kernel, polar_kernel, north_pole_kernel=None, equatorial_bias=None,
                            polar_bias=None, north_pole_bias=None, dilation_rate=(1, 1), flip_north_pole=True, halo=0,
                            activation=None, padding='valid'):
    """
    CubeSphereConv2D.call (stride 1) on float32 tensors with float32 ACCURACY on the tensor cores (forward only): x is split
    into bf16 hi / lo parts laid out [hi | lo | hi] (dlwpcs_split3), the kernels into [w_hi ; w_hi ; w_lo], and the bf16
    tcgen05 kernel accumulates x_hi*w_hi + x_lo*w_hi + x_hi*w_lo in float32 -- 16 mantissa bits per operand instead of the
    8 of the plain bf16 path; the result is float32.  Channel counts are padded to multiples of 8 internally.
    """
    _check_cl(x)
    if x.dtype != torch.float32:
        raise ValueError('cube_sphere_conv2d_tc32 takes float32 tensors')
    act = resolve_activation(activation)
    if act is None:
        raise ValueError('activation %r cannot be fused; apply it to the output instead' % (activation,))
    b, _, n, _, cin = x.shape
    kh, kw, wcin, cout = equatorial_kernel.shape
    if wcin != cin:
        raise ValueError('input has %d channels, kernel expects %d' % (cin, wcin))
    pad_in = (-cin) % 8
    xp = torch.nn.functional.pad(x, (0, pad_in)) if pad_in else x
    xs = _lib.split3(xp)
    indep = north_pole_kernel is not None
    d = _lib.make_desc(b, n, 3 * (cin + pad_in), cout, (kh, kw), (1, 1), tuple(dilation_rate), int(halo),
                       padding.lower() == 'same', bool(flip_north_pole), indep, equatorial_bias is not None, act[0], act[1],
                       act[2], _lib.BF16, _lib.F32)
    ws = [None if w is None else _lib.split3_weights(w, pad_in) for w in (equatorial_kernel, polar_kernel, north_pole_kernel)]
    packed = _lib.pack_weights(d, ws[0], ws[1], ws[2], equatorial_bias, polar_bias, north_pole_bias)
    return _lib.conv2d_fwd(d, xs, None, packed)


class _Act(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, act):
        y = _lib.act_fwd(x.contiguous(), *act)
        ctx.act = act
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        return _lib.act_bwd(dy.contiguous(), y, *ctx.act), None


def capped_leaky_relu(x, negative_slope=0.1, max_value=10.0):
    """keras ReLU(negative_slope, max_value) (Azure/train_cs.py:199) as a standalone op."""
    return _Act.apply(x, (_lib.ACT_CAPPED_LEAKY_RELU, float(negative_slope), float(max_value)))


class _Pool2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return _lib.pool2(x, backward=False)

    @staticmethod
    def backward(ctx, dy):
        return _lib.pool2(dy.contiguous(), backward=True)


class _Up2Cat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        ctx.ca = a.shape[-1]
        return _lib.up2cat_fwd(a, b)

    @staticmethod
    def backward(ctx, dt):
        return _lib.up2cat_bwd(dt.contiguous(), ctx.ca)


def avg_pool_2x2(x):
    """AveragePooling3D((1,2,2)) (Azure/train_cs.py:197), channels_last; one 16-byte kernel each way."""
    _check_cl(x)
    if x.shape[2] % 2:
        raise ValueError('face edge %d is not divisible by 2' % x.shape[2])
    return _Pool2.apply(x.contiguous())


def upsample_concat(a, b):
    """concatenate([UpSampling3D((1,2,2))(a), b]) (Azure/train_cs.py:198, 293, 299): a (B,6,n/2,n/2,Ca), b (B,6,n,n,Cb)."""
    _check_cl(a)
    _check_cl(b)
    if 2 * a.shape[2] != b.shape[2] or a.shape[0] != b.shape[0] or a.dtype != b.dtype:
        raise ValueError('upsample_concat: incompatible tensors %r and %r' % (tuple(a.shape), tuple(b.shape)))
    return _Up2Cat.apply(a.contiguous(), b.contiguous())
