"""
TEST INFRASTRUCTURE ONLY -- never imported by the product path (dlwp_cs_b200/).

A numpy stand-in for the handful of TensorFlow 2.1 / Keras symbols that the reference module
``/root/reference/DLWP/custom.py`` touches, so that the reference's *own* ``CubeSpherePadding2D.call``
(custom.py:1082-1308) and ``CubeSphereConv2D.call`` (custom.py:921-1002) bodies can be executed unmodified in this
container (TensorFlow itself is not installable here: SURVEY.md section 8c).  ``tests/golden/make_golden.py`` uses it to
produce the committed golden vectors; nothing on the GPU box needs it.

What is *real reference code* under this shim: every slice / reverse / transpose / concatenate of the halo exchange,
the face -> kernel dispatch, the north-pole flip, bias add and the face re-stacking.
What is *restated* (TensorFlow semantics re-implemented here because TF is an absent third-party dependency,
tensorflow==2.1.0 per environment.yml:180):
  * ``K.conv2d``   -- cross-correlation (no kernel flip), HWIO kernel, 'valid' = no padding, 'same' = TF's asymmetric
                      zero padding (extra element at the bottom/right), strides, dilation.  Implemented as a float64
                      direct sum, independent of torch/oneDNN.
  * ``K.bias_add``, ``K.concatenate``, ``K.reverse``, ``K.expand_dims``, ``tf.transpose`` -- numpy one-liners.
  * ``keras.layers.ZeroPadding3D.__init__`` padding normalisation (int -> 3 symmetric pairs, 3-tuple of ints or pairs);
    a 2-tuple such as the reference's declared default ``(1, 1)`` raises ValueError exactly like Keras does.
"""
import sys
import types

import numpy as np


# --------------------------------------------------------------------------------------------------------------------
# backend ops
# --------------------------------------------------------------------------------------------------------------------

def _concatenate(tensors, axis=-1):
    return np.concatenate([np.asarray(t) for t in tensors], axis=axis)


def _reverse(x, axes):
    if isinstance(axes, int):
        axes = [axes]
    return np.flip(np.asarray(x), axis=tuple(axes))


def _expand_dims(x, axis=-1):
    return np.expand_dims(np.asarray(x), axis)


def _transpose(a, perm=None):
    return np.transpose(np.asarray(a), perm)


def _same_pads(size, k_eff, stride):
    """TensorFlow 'SAME' padding: total = max((ceil(size/stride)-1)*stride + k_eff - size, 0); extra goes after."""
    out = -(-size // stride)
    total = max((out - 1) * stride + k_eff - size, 0)
    return total // 2, total - total // 2


def conv2d_tf(x, kernel, strides=(1, 1), padding='valid', data_format=None, dilation_rate=(1, 1)):
    """
    ``tf.keras.backend.conv2d`` semantics on numpy arrays, float64 direct sum.
    x: (B,H,W,C) for channels_last / (B,C,H,W) for channels_first; kernel HWIO (kh,kw,Cin,Cout).
    """
    x = np.asarray(x)
    kernel = np.asarray(kernel)
    out_dtype = np.result_type(x.dtype, kernel.dtype)
    if data_format in (None, 'channels_last'):
        cl = True
    elif data_format == 'channels_first':
        cl = False
        x = np.transpose(x, (0, 2, 3, 1))
    else:
        raise ValueError('Unknown data_format: %r' % (data_format,))
    if isinstance(strides, int):
        strides = (strides, strides)
    if isinstance(dilation_rate, int):
        dilation_rate = (dilation_rate, dilation_rate)
    sh, sw = strides
    dh, dw = dilation_rate
    kh, kw, cin, cout = kernel.shape
    if x.shape[-1] != cin:
        raise ValueError('input depth %d != kernel depth %d' % (x.shape[-1], cin))
    x64 = x.astype(np.float64)
    k64 = kernel.astype(np.float64)
    padding = padding.lower()
    if padding == 'same':
        pt, pb = _same_pads(x64.shape[1], (kh - 1) * dh + 1, sh)
        pl, pr = _same_pads(x64.shape[2], (kw - 1) * dw + 1, sw)
        x64 = np.pad(x64, ((0, 0), (pt, pb), (pl, pr), (0, 0)))
    elif padding != 'valid':
        raise ValueError('Unknown padding: %r' % (padding,))
    b, h, w, _ = x64.shape
    ho = (h - ((kh - 1) * dh + 1)) // sh + 1
    wo = (w - ((kw - 1) * dw + 1)) // sw + 1
    y = np.zeros((b, ho, wo, cout), dtype=np.float64)
    for u in range(kh):
        for v in range(kw):
            patch = x64[:, u * dh: u * dh + (ho - 1) * sh + 1: sh, v * dw: v * dw + (wo - 1) * sw + 1: sw, :]
            y += np.tensordot(patch, k64[u, v], axes=([3], [0]))
    y = y.astype(out_dtype)
    if not cl:
        y = np.transpose(y, (0, 3, 1, 2))
    return y


def _bias_add(x, bias, data_format=None):
    x = np.asarray(x)
    bias = np.asarray(bias)
    if data_format in (None, 'channels_last'):
        return x + bias.reshape((1,) * (x.ndim - 1) + (-1,))
    if data_format == 'channels_first':
        return x + bias.reshape((1, -1) + (1,) * (x.ndim - 2))
    raise ValueError('Unknown data_format: %r' % (data_format,))


# --------------------------------------------------------------------------------------------------------------------
# keras bits
# --------------------------------------------------------------------------------------------------------------------

class _Layer(object):
    """Just enough of keras.layers.Layer: attribute bag + add_weight that draws from a seeded numpy generator."""
    _rng = np.random.default_rng(0)

    def __init__(self, name=None, dtype=None, **kwargs):
        if kwargs:
            raise TypeError('Keyword argument not understood: %s' % sorted(kwargs))
        self.name = name
        self.built = False
        self.weights = []

    def add_weight(self, shape=None, initializer=None, name=None, regularizer=None, constraint=None, **kwargs):
        w = initializer(tuple(shape)) if callable(initializer) else np.zeros(tuple(shape))
        self.weights.append((name, w))
        return w

    def get_config(self):
        return {'name': self.name}

    def __call__(self, inputs, **kwargs):
        if not self.built and hasattr(self, 'build'):
            self.build(np.shape(inputs))
            self.built = True
        return self.call(inputs, **kwargs)


def _normalize_tuple(value, n, name):
    if isinstance(value, int):
        return (value,) * n
    try:
        value_tuple = tuple(value)
    except TypeError:
        raise ValueError('The `%s` argument must be a tuple of %d integers. Received: %s' % (name, n, value))
    if len(value_tuple) != n:
        raise ValueError('The `%s` argument must be a tuple of %d integers. Received: %s' % (name, n, value))
    for v in value_tuple:
        if not isinstance(v, (int, np.integer)):
            raise ValueError('The `%s` argument must be a tuple of %d integers. Received: %s' % (name, n, value))
    return tuple(int(v) for v in value_tuple)


def _normalize_padding(value):
    if isinstance(value, (list, tuple)):
        return value
    padding = value.lower()
    if padding not in {'valid', 'same', 'causal'}:
        raise ValueError('The `padding` argument must be one of "valid", "same" (or "causal"). Received: %s' % padding)
    return padding


def _normalize_data_format(value):
    if value is None:
        value = 'channels_last'
    data_format = value.lower()
    if data_format not in {'channels_first', 'channels_last'}:
        raise ValueError('The `data_format` argument must be one of "channels_first", "channels_last". '
                         'Received: %s' % value)
    return data_format


def _conv_output_length(input_length, filter_size, padding, stride, dilation=1):
    if input_length is None:
        return None
    dilated = filter_size + (filter_size - 1) * (dilation - 1)
    if padding in ('same', 'causal'):
        output_length = input_length
    elif padding == 'valid':
        output_length = input_length - dilated + 1
    elif padding == 'full':
        output_length = input_length + dilated - 1
    else:
        raise ValueError(padding)
    return (output_length + stride - 1) // stride


class _ZeroPadding3D(_Layer):
    def __init__(self, padding=(1, 1, 1), data_format=None, **kwargs):
        super(_ZeroPadding3D, self).__init__(**kwargs)
        self.data_format = _normalize_data_format(data_format)
        if isinstance(padding, int):
            self.padding = ((padding, padding), (padding, padding), (padding, padding))
        elif hasattr(padding, '__len__'):
            if len(padding) != 3:
                raise ValueError('`padding` should have 3 elements. Found: ' + str(padding))
            self.padding = tuple(_normalize_tuple(pd, 2, '%dst entry of padding' % (i + 1))
                                 for i, pd in enumerate(padding))
        else:
            raise ValueError('`padding` should be either an int, a tuple of 3 ints, or a tuple of 3 tuples of 2 ints.')


class _ZeroPadding2D(_Layer):
    def __init__(self, padding=(1, 1), data_format=None, **kwargs):
        super(_ZeroPadding2D, self).__init__(**kwargs)
        self.data_format = _normalize_data_format(data_format)
        self.padding = padding


class _Getter(types.ModuleType):
    """activations / initializers / regularizers / constraints: get() passes callables and None through."""

    def __init__(self, name, table=None):
        super(_Getter, self).__init__(name)
        self._table = table or {}

    def get(self, identifier):
        if identifier is None or callable(identifier):
            return identifier
        if identifier in self._table:
            return self._table[identifier]
        raise ValueError('unknown identifier for %s: %r' % (self.__name__, identifier))

    def serialize(self, obj):
        return getattr(obj, '__name__', None)


def _glorot_uniform(shape):
    shape = tuple(shape)
    if len(shape) < 2:
        fan_in = fan_out = shape[0]
    else:
        receptive = int(np.prod(shape[:-2]))
        fan_in, fan_out = shape[-2] * receptive, shape[-1] * receptive
    limit = np.sqrt(6.0 / (fan_in + fan_out))
    return _Layer._rng.uniform(-limit, limit, size=shape)


def _zeros(shape):
    return np.zeros(tuple(shape))


def _linear(x):
    return x


def _relu(x):
    return np.maximum(x, 0)


def install():
    """Register the fake ``tensorflow`` package tree in sys.modules (idempotent). Returns the fake tf module."""
    if 'tensorflow' in sys.modules and getattr(sys.modules['tensorflow'], '_dlwp_shim', False):
        return sys.modules['tensorflow']
    if 'tensorflow' in sys.modules:
        raise RuntimeError('a real tensorflow is already imported; the shim must not shadow it')

    def mod(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m

    tf = mod('tensorflow')
    tf._dlwp_shim = True
    tf.transpose = _transpose

    compat = mod('tensorflow.compat')
    v1 = mod('tensorflow.compat.v1')
    tfk = mod('tensorflow.compat.v1.keras')
    backend = mod('tensorflow.compat.v1.keras.backend')
    backend.concatenate = _concatenate
    backend.reverse = _reverse
    backend.expand_dims = _expand_dims
    backend.conv2d = conv2d_tf
    backend.bias_add = _bias_add
    tfk.backend = backend
    v1.keras = tfk
    compat.v1 = v1
    tf.compat = compat

    keras = mod('tensorflow.keras')
    tf.keras = keras
    callbacks = mod('tensorflow.keras.callbacks')
    callbacks.Callback = type('Callback', (object,), {})
    callbacks.EarlyStopping = type('EarlyStopping', (callbacks.Callback,), {})
    layers = mod('tensorflow.keras.layers')
    layers.Layer = _Layer
    layers.ZeroPadding2D = _ZeroPadding2D
    layers.ZeroPadding3D = _ZeroPadding3D
    layers.LocallyConnected2D = type('LocallyConnected2D', (_Layer,), {})
    layers.Lambda = type('Lambda', (_Layer,), {})
    losses = mod('tensorflow.keras.losses')
    losses.mean_absolute_error = lambda a, b: np.mean(np.abs(a - b), axis=-1)
    losses.mean_squared_error = lambda a, b: np.mean((a - b) ** 2, axis=-1)
    keras.callbacks, keras.layers, keras.losses = callbacks, layers, losses
    kmodels = mod('tensorflow.keras.models')              # DLWP/util.py imports these two names at module level
    kutils = mod('tensorflow.keras.utils')
    kutils.multi_gpu_model = lambda model, gpus=None: model
    kutils.Sequence = type('Sequence', (object,), {})        # base class of the reference's data generators
    keras.models, keras.utils = kmodels, kutils
    for nm, table in (('activations', {'linear': _linear, 'relu': _relu}),
                      ('initializers', {'glorot_uniform': _glorot_uniform, 'zeros': _zeros}),
                      ('regularizers', {}), ('constraints', {})):
        g = _Getter('tensorflow.keras.' + nm, table)
        sys.modules['tensorflow.keras.' + nm] = g
        setattr(keras, nm, g)

    python = mod('tensorflow.python')
    pkeras = mod('tensorflow.python.keras')
    putils = mod('tensorflow.python.keras.utils')
    conv_utils = mod('tensorflow.python.keras.utils.conv_utils')
    conv_utils.normalize_tuple = _normalize_tuple
    conv_utils.normalize_padding = _normalize_padding
    conv_utils.normalize_data_format = _normalize_data_format
    conv_utils.conv_output_length = _conv_output_length
    putils.conv_utils = conv_utils
    engine = mod('tensorflow.python.keras.engine')
    base_layer = mod('tensorflow.python.keras.engine.base_layer')
    base_layer.InputSpec = lambda **kw: dict(kw)
    engine.base_layer = base_layer
    pkeras.utils, pkeras.engine = putils, engine
    python.keras = pkeras
    tf.python = python
    return tf


def load_reference_custom(path='/root/reference/DLWP/custom.py'):
    """Import the reference's custom.py *by file path* (bypassing DLWP/__init__) on top of the shim."""
    import importlib.util
    install()
    spec = importlib.util.spec_from_file_location('_dlwp_reference_custom', path)
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    return module


def load_reference_util(path='/root/reference/DLWP/util.py'):
    """Import the reference's util.py by file path on top of the shim (for `insolation`, util.py:306-364)."""
    import importlib.util
    install()
    spec = importlib.util.spec_from_file_location('_dlwp_reference_util', path)
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    return module


def load_reference_generators(root='/root/reference'):
    """Import the reference's DLWP/model/generators.py (ArrayDataGenerator, generators.py:636-1011) on top of the shim.
    The module uses a relative import of ..util, so stub package objects for ``DLWP`` and ``DLWP.model`` are registered
    without running their __init__ (which would pull in Keras models); xarray is only needed by the other generator
    classes and is replaced by an empty module; numpy >= 1.24 dropped the ``np.int`` alias generators.py:874 still uses."""
    import importlib
    import os
    import numpy
    install()
    if not hasattr(numpy, 'int'):
        numpy.int = int
    if 'xarray' not in sys.modules:
        sys.modules['xarray'] = types.ModuleType('xarray')
    for name, sub in (('DLWP', 'DLWP'), ('DLWP.model', os.path.join('DLWP', 'model'))):
        if name not in sys.modules:
            pkg = types.ModuleType(name)
            pkg.__path__ = [os.path.join(root, sub)]
            sys.modules[name] = pkg
    importlib.import_module('DLWP.util')
    return importlib.import_module('DLWP.model.generators')
