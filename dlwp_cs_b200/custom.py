"""
Drop-in layer classes for the cubed-sphere hot path, with the constructor / call signatures of the reference's
``DLWP.custom`` (CubeSpherePadding2D: custom.py:1057-1308, CubeSphereConv2D: custom.py:755-1054), as ``torch.nn.Module``s
backed by libdlwpcs (hand-written sm_100a kernels).  Resolvable by name the way ``DLWPTorchNN.build_model`` does it
(models_torch.py:131-135 -> ``util.get_from_class('DLWP.custom', name)``).

Differences, all deliberate:
  * ``CubeSpherePadding2D(padding=(1, 1))`` -- the reference's declared default is rejected by keras ``ZeroPadding3D``
    (custom.py:1073-1080), so the reference can only be called with an int or a 3-tuple; here the 2-tuple raises the
    same ValueError, and the default is the int 1 that every call site passes (Azure/train_cs.py:196).
  * the input depth is inferred at the first call like keras ``build`` (custom.py:871-919) unless ``in_channels=`` is
    given; parameters carry the reference's names and HWIO layout.
  * ``fuse_padding=p`` (extension) folds a preceding CubeSpherePadding2D(p) into the convolution kernel's load stage.
"""
import math

import torch
import torch.nn as nn
from torch.nn.parameter import UninitializedParameter

from . import _lib
from . import functional as F_cs


def _normalize_tuple(value, n, name):
    if isinstance(value, int):
        return (value,) * n
    try:
        t = tuple(value)
    except TypeError:
        raise ValueError('The `%s` argument must be a tuple of %d integers. Received: %s' % (name, n, value))
    if len(t) != n or not all(isinstance(v, int) for v in t):
        raise ValueError('The `%s` argument must be a tuple of %d integers. Received: %s' % (name, n, value))
    return t


def _normalize_data_format(value):
    if value is None:
        value = 'channels_last'
    v = str(value).lower()
    if v not in ('channels_first', 'channels_last'):
        raise ValueError('The `data_format` argument must be one of "channels_first", "channels_last". '
                         'Received: %s' % value)
    return v


def _to_channels_last(x, data_format):
    return x.permute(0, 2, 3, 4, 1) if data_format == 'channels_first' else x


def _from_channels_last(y, data_format):
    return y.permute(0, 4, 1, 2, 3) if data_format == 'channels_first' else y


class CubeSpherePadding2D(nn.Module):
    """
    Padding layer for 2D data on a cubed sphere (reference custom.py:1057-1308): pads both face axes by ``padding``
    with data from the neighbouring faces.  Input (batch, channels, 6, h, w) for channels_first or
    (batch, 6, h, w, channels) for channels_last; faces 4 and 5 are the south and north polar faces.
    """

    def __init__(self, padding=1, data_format='channels_first', **kwargs):
        super().__init__()
        if kwargs:
            raise TypeError('Keyword argument not understood: %s' % sorted(kwargs))
        self.data_format = _normalize_data_format(data_format)
        # keras ZeroPadding3D normalisation (custom.py:1077-1080), then the face axis is forced to (0, 0)
        if isinstance(padding, int):
            pads = ((padding, padding),) * 3
        elif hasattr(padding, '__len__'):
            if len(padding) != 3:
                raise ValueError('`padding` should have 3 elements. Found: ' + str(padding))
            pads = tuple(_normalize_tuple(p, 2, '%dst entry of padding' % (i + 1)) for i, p in enumerate(padding))
        else:
            raise ValueError('`padding` should be either an int, a tuple of 3 ints, or a tuple of 3 tuples of 2 ints.')
        self.padding = ((0, 0),) + tuple(pads[1:])
        if self.padding[1][0] < 0:
            raise ValueError('padding must be non-negative')

    def forward(self, inputs):
        p = self.padding[1][0]          # the only entry the reference reads (custom.py:1083)
        y = F_cs.cube_sphere_pad(_to_channels_last(inputs, self.data_format), p)
        return _from_channels_last(y, self.data_format)

    def compute_output_shape(self, input_shape):
        p = self.padding[1][0]
        s = list(input_shape)
        ax = (3, 4) if self.data_format == 'channels_first' else (2, 3)
        for a in ax:
            s[a] = None if s[a] is None else s[a] + 2 * p
        return tuple(s)

    def get_config(self):
        return {'padding': self.padding, 'data_format': self.data_format}

    def extra_repr(self):
        return 'padding=%d, data_format=%s' % (self.padding[1][0], self.data_format)


_INITIALIZERS = ('glorot_uniform', 'zeros', 'ones')


def _initialize(t, how):
    if callable(how):
        with torch.no_grad():
            t.copy_(torch.as_tensor(how(tuple(t.shape)), dtype=t.dtype))
    elif how == 'glorot_uniform':
        if t.dim() >= 2:                                   # keras VarianceScaling fans for an HWIO kernel
            rf = int(math.prod(t.shape[:-2]))
            fan_in, fan_out = t.shape[-2] * rf, t.shape[-1] * rf
        else:
            fan_in = fan_out = t.shape[0]
        limit = math.sqrt(6.0 / (fan_in + fan_out))
        nn.init.uniform_(t, -limit, limit)
    elif how == 'zeros':
        nn.init.zeros_(t)
    elif how == 'ones':
        nn.init.ones_(t)
    else:
        raise ValueError('unknown initializer %r (supported: %s or a callable shape -> array)' % (how, _INITIALIZERS))


class CubeSphereConv2D(nn.modules.lazy.LazyModuleMixin, nn.Module):
    """
    2D convolution for data on a cubed sphere (reference custom.py:755-1054).  Learns one kernel/bias for the four
    equatorial faces and one for the polar faces (optionally a separate one for the north pole); the north-pole face can
    be row-reversed around the convolution (``flip_north_pole``).  Should be preceded by ``CubeSpherePadding2D`` (or use
    ``fuse_padding``), otherwise the faces are not connected.
    """
    cls_to_become = None

    def __init__(self, filters, kernel_size, strides=1, padding='valid', data_format='channels_first', dilation_rate=1,
                 activation=None, use_bias=True, flip_north_pole=True, independent_north_pole=False,
                 kernel_initializer='glorot_uniform', bias_initializer='zeros', kernel_regularizer=None,
                 bias_regularizer=None, activity_regularizer=None, kernel_constraint=None, bias_constraint=None,
                 in_channels=None, fuse_padding=0, name=None, **kwargs):
        super().__init__()
        if kwargs:
            raise TypeError('Keyword argument not understood: %s' % sorted(kwargs))
        self.filters = int(filters)
        self.kernel_size = _normalize_tuple(kernel_size, 2, 'kernel_size')
        self.strides = _normalize_tuple(strides, 2, 'strides')
        padding = str(padding).lower()
        if padding not in ('valid', 'same'):
            raise ValueError('The `padding` argument must be one of "valid", "same". Received: %s' % padding)
        self.padding = padding
        self.data_format = _normalize_data_format(data_format)
        self.dilation_rate = _normalize_tuple(dilation_rate, 2, 'dilation_rate')
        self.activation = activation
        self._fused_act = F_cs.resolve_activation(activation)
        if self._fused_act is None and not callable(activation):
            raise ValueError('unknown activation %r' % (activation,))
        self.use_bias = bool(use_bias)
        self.flip_north_pole = bool(flip_north_pole)
        self.independent_north_pole = bool(independent_north_pole)
        self.kernel_initializer = kernel_initializer
        self.bias_initializer = bias_initializer
        for nm, v in (('kernel_regularizer', kernel_regularizer), ('bias_regularizer', bias_regularizer),
                      ('activity_regularizer', activity_regularizer), ('kernel_constraint', kernel_constraint),
                      ('bias_constraint', bias_constraint)):
            if v is not None:
                raise NotImplementedError('%s is not supported (no cubed-sphere model in the reference sets it)' % nm)
        self.fuse_padding = int(fuse_padding)
        self.name = name
        self.rank = 3
        self.equatorial_kernel = UninitializedParameter()
        self.polar_kernel = UninitializedParameter()
        self.north_pole_kernel = UninitializedParameter() if self.independent_north_pole else None
        self.equatorial_bias = UninitializedParameter() if self.use_bias else None
        self.polar_bias = UninitializedParameter() if self.use_bias else None
        self.north_pole_bias = UninitializedParameter() if (self.use_bias and self.independent_north_pole) else None
        if in_channels is not None:
            self.build_with_channels(int(in_channels))

    # keras `build` (custom.py:871-919): parameters in the reference's creation order
    def build_with_channels(self, input_dim):
        kernel_shape = self.kernel_size + (input_dim, self.filters)
        for nm in ('equatorial_kernel', 'polar_kernel', 'north_pole_kernel'):
            p = getattr(self, nm)
            if p is not None and isinstance(p, UninitializedParameter):
                p.materialize(kernel_shape)
                _initialize(p.data, self.kernel_initializer)
        for nm in ('equatorial_bias', 'polar_bias', 'north_pole_bias'):
            p = getattr(self, nm)
            if p is not None and isinstance(p, UninitializedParameter):
                p.materialize((self.filters,))
                _initialize(p.data, self.bias_initializer)

    def initialize_parameters(self, inputs, **_):     # LazyModuleMixin hook, first call only
        if self.has_uninitialized_params():
            ch_axis = 1 if self.data_format == 'channels_first' else -1
            if inputs.dim() != self.rank + 2:
                raise ValueError('expected a 5-dimensional input, got %d dimensions' % inputs.dim())
            if inputs.shape[ch_axis] is None:
                raise ValueError('The channel dimension of the inputs should be defined. Found `None`.')
            self.build_with_channels(int(inputs.shape[ch_axis]))

    def reset_parameters(self):
        if not self.has_uninitialized_params():
            for nm in ('equatorial_kernel', 'polar_kernel', 'north_pole_kernel'):
                if getattr(self, nm) is not None:
                    _initialize(getattr(self, nm).data, self.kernel_initializer)
            for nm in ('equatorial_bias', 'polar_bias', 'north_pole_bias'):
                if getattr(self, nm) is not None:
                    _initialize(getattr(self, nm).data, self.bias_initializer)

    def forward(self, inputs, in_act=None, dy_premasked=False, keep_pad=False):
        """in_act / dy_premasked (used by CubeSphereCNN between directly chained layers, see functional.cube_sphere_conv2d):
        fuse the previous layer's activation derivative into this layer's input gradient."""
        x = _to_channels_last(inputs, self.data_format)
        fused = self._fused_act is not None
        kernels = (self.equatorial_kernel, self.polar_kernel, self.north_pole_kernel)
        biases = (self.equatorial_bias, self.polar_bias, self.north_pole_bias)
        pad_in = pad_out = 0
        if x.dtype == torch.bfloat16:
            # bf16 tensor-core path: 16-byte gathers / stores need channel counts that are multiples of 8 -- zero input
            # channels with zero weights and zero-weight output channels leave the result unchanged (the operator pads
            # the packed weights only; the parameters and their gradients keep their shape)
            pad_in, pad_out = (-x.shape[-1]) % 8, (-self.filters) % 8
            if pad_in:
                x = torch.nn.functional.pad(x, (0, pad_in))
        y = F_cs.cube_sphere_conv2d(x, kernels[0], kernels[1], kernels[2], biases[0], biases[1], biases[2], self.strides,
                                    self.padding, self.dilation_rate, self.flip_north_pole, self.fuse_padding,
                                    self.activation if fused else None, pad_in=pad_in, pad_out=pad_out,
                                    in_act=in_act if not pad_in else None,
                                    dy_premasked=dy_premasked and fused)
        if pad_out and not keep_pad:       # keep_pad: hand out the zero pad channels too (the trainer's loss ignores them)
            y = y[..., :self.filters]
        y = _from_channels_last(y, self.data_format)
        if not fused:
            y = self.activation(y)
        return y

    def pack_descriptor(self, dtype=torch.bfloat16):
        """(descriptor, parameter tuple) of this layer's packed weight images for activations of `dtype`, as ``forward``
        builds them (channel counts padded to multiples of 8 on the bf16 path; batch / face edge are placeholders: the
        images do not depend on them)."""
        kh, kw, wcin, cout = self.equatorial_kernel.shape
        bf = dtype == torch.bfloat16
        pad_in, pad_out = ((-wcin) % 8, (-cout) % 8) if bf else (0, 0)
        act = F_cs.resolve_activation(self.activation) if self._fused_act is not None else F_cs.resolve_activation(None)
        n = max(8, (kh - 1) * self.dilation_rate[0] + 1, (kw - 1) * self.dilation_rate[1] + 1)
        code = _lib.dtype_code(dtype)
        d = _lib.make_desc(1, n, wcin + pad_in, cout + pad_out, (kh, kw), tuple(self.strides), tuple(self.dilation_rate),
                           self.fuse_padding, self.padding == 'same', self.flip_north_pole,
                           self.north_pole_kernel is not None, self.equatorial_bias is not None, act[0], act[1], act[2],
                           code, code)
        ws = (self.equatorial_kernel, self.polar_kernel, self.north_pole_kernel, self.equatorial_bias, self.polar_bias,
              self.north_pole_bias)
        return d, ws

    def compute_output_shape(self, input_shape):
        from ._lib import conv_out_edge
        sp = input_shape[2:4] if self.data_format == 'channels_last' else input_shape[-2:]
        new = tuple(conv_out_edge(sp[i] + 2 * self.fuse_padding, self.kernel_size[i], self.strides[i],
                                  self.dilation_rate[i], self.padding == 'same') for i in range(2))
        if self.data_format == 'channels_last':
            return (input_shape[0], 6) + new + (self.filters,)
        return (input_shape[0], self.filters, 6) + new

    def get_config(self):
        act = self.activation
        return {'filters': self.filters, 'kernel_size': self.kernel_size, 'strides': self.strides,
                'padding': self.padding, 'data_format': self.data_format, 'dilation_rate': self.dilation_rate,
                'activation': act if (act is None or isinstance(act, (str, tuple, list))) else getattr(act, '__name__', None),
                'use_bias': self.use_bias, 'flip_north_pole': self.flip_north_pole,
                'independent_north_pole': self.independent_north_pole, 'kernel_initializer': self.kernel_initializer,
                'bias_initializer': self.bias_initializer, 'kernel_regularizer': None, 'bias_regularizer': None,
                'activity_regularizer': None, 'kernel_constraint': None, 'bias_constraint': None}

    def get_weights(self):
        """numpy copies in the reference's ``add_weight`` order (custom.py:882-914)."""
        out = []
        for nm in ('equatorial_kernel', 'polar_kernel', 'north_pole_kernel', 'equatorial_bias', 'polar_bias',
                   'north_pole_bias'):
            p = getattr(self, nm)
            if p is not None:
                out.append(p.detach().cpu().numpy())
        return out

    def set_weights(self, weights):
        names = [nm for nm in ('equatorial_kernel', 'polar_kernel', 'north_pole_kernel', 'equatorial_bias',
                               'polar_bias', 'north_pole_bias') if getattr(self, nm) is not None]
        if len(weights) != len(names):
            raise ValueError('expected %d weight arrays, got %d' % (len(names), len(weights)))
        if self.has_uninitialized_params():
            self.build_with_channels(int(weights[0].shape[2]))
        with torch.no_grad():
            for nm, w in zip(names, weights):
                p = getattr(self, nm)
                p.copy_(torch.as_tensor(w, dtype=p.dtype).reshape(p.shape))

    def extra_repr(self):
        return 'filters=%d, kernel_size=%s, padding=%s, data_format=%s, fuse_padding=%d' % (
            self.filters, self.kernel_size, self.padding, self.data_format, self.fuse_padding)
