"""
Generate the golden vectors under tests/golden/ by running the REFERENCE'S OWN code
(/root/reference/DLWP/custom.py: CubeSpherePadding2D.call 1082-1308, CubeSphereConv2D.build/call 871-1002) on the numpy
TensorFlow shim in oracle/tf_shim.py.  Run in the build container only (the reference is not on the GPU box):

    python tests/golden/make_golden.py

Outputs (committed):
    pad_luts.npz      index-encoded halo exchange for (N,p) in {(4,1),(5,2),(8,3),(48,1),(24,1),(12,1),(6,2)}, both data formats
    conv_cases.npz    CubeSphereConv2D outputs for a set of constructor variants on seeded inputs (float64)
    padconv_cfg1.npz  BASELINE config 1: pad(1) -> conv 3->3, C48, batch 1, seed 0 (float32 inputs, float64 result)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..', 'oracle'))
sys.path.insert(0, HERE)
import tf_shim  # noqa: E402
from cases import CONV_CASES, PAD_CASES  # noqa: E402

ref = tf_shim.load_reference_custom()


def reference_pad(x, p, data_format):
    return np.asarray(ref.CubeSpherePadding2D(p, data_format=data_format)(x))


def make_pad_luts():
    out = {}
    for n, p in PAD_CASES:
        idx = np.arange(6 * n * n, dtype=np.int64).reshape(1, 6, n, n, 1)
        cl = reference_pad(idx, p, 'channels_last')
        cf = reference_pad(np.transpose(idx, (0, 4, 1, 2, 3)), p, 'channels_first')
        assert cl.shape == (1, 6, n + 2 * p, n + 2 * p, 1) and cf.shape == (1, 1, 6, n + 2 * p, n + 2 * p)
        out['cl_n%d_p%d' % (n, p)] = cl.reshape(6, n + 2 * p, n + 2 * p).astype(np.int32)
        out['cf_n%d_p%d' % (n, p)] = cf.reshape(6, n + 2 * p, n + 2 * p).astype(np.int32)
    np.savez_compressed(os.path.join(HERE, 'pad_luts.npz'), **out)
    return out


def build_layer(kwargs, rng, data_format):
    """Instantiate the reference layer; replace its weights by seeded values (biases non-zero to exercise the path)."""
    layer = ref.CubeSphereConv2D(data_format=data_format, **kwargs)
    return layer


def make_conv_cases():
    out = {}
    for name, kwargs, b, h, cin in CONV_CASES:
        rng = np.random.default_rng(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)))
        x = rng.standard_normal((b, 6, h, h, cin))
        for fmt in ('channels_last', 'channels_first'):
            layer = ref.CubeSphereConv2D(data_format=fmt, **kwargs)
            xin = x if fmt == 'channels_last' else np.transpose(x, (0, 4, 1, 2, 3))
            layer.build(xin.shape)                       # custom.py:871-919 (allocates in reference order)
            wrng = np.random.default_rng(1234 + sum(map(ord, name)))
            for attr in ('equatorial_kernel', 'polar_kernel', 'north_pole_kernel',
                         'equatorial_bias', 'polar_bias', 'north_pole_bias'):
                cur = getattr(layer, attr)
                if cur is not None:
                    setattr(layer, attr, wrng.uniform(-0.5, 0.5, size=np.shape(cur)))
            y = np.asarray(layer.call(xin))
            assert tuple(y.shape) == tuple(layer.compute_output_shape(xin.shape)), (name, fmt, y.shape)
            if fmt == 'channels_last':
                out[name + '.x'] = x
                for attr in ('equatorial_kernel', 'polar_kernel', 'north_pole_kernel',
                             'equatorial_bias', 'polar_bias', 'north_pole_bias'):
                    if getattr(layer, attr) is not None:
                        out[name + '.' + attr] = np.asarray(getattr(layer, attr))
                out[name + '.y_cl'] = y
            else:
                out[name + '.y_cf'] = y
    np.savez_compressed(os.path.join(HERE, 'conv_cases.npz'), **out)
    return out


def make_cfg1():
    """BASELINE.json configs[0]: single CubeSphereConv2D fwd, C48, 3->3 channels, batch 1, preceded by the padding."""
    rng = np.random.default_rng(0)
    x = rng.standard_normal((1, 6, 48, 48, 3)).astype(np.float32)
    limit = np.sqrt(6.0 / (9 * 3 + 9 * 3))
    w_eq = rng.uniform(-limit, limit, size=(3, 3, 3, 3)).astype(np.float32)
    w_pol = rng.uniform(-limit, limit, size=(3, 3, 3, 3)).astype(np.float32)
    b_eq = rng.uniform(-0.1, 0.1, size=(3,)).astype(np.float32)
    b_pol = rng.uniform(-0.1, 0.1, size=(3,)).astype(np.float32)
    pad = ref.CubeSpherePadding2D(1, data_format='channels_last')
    conv = ref.CubeSphereConv2D(3, 3, padding='valid', data_format='channels_last')
    xp = np.asarray(pad(x))
    conv.build(xp.shape)
    conv.equatorial_kernel, conv.polar_kernel = w_eq.astype(np.float64), w_pol.astype(np.float64)
    conv.equatorial_bias, conv.polar_bias = b_eq.astype(np.float64), b_pol.astype(np.float64)
    y = np.asarray(conv.call(xp.astype(np.float64)))
    np.savez_compressed(os.path.join(HERE, 'padconv_cfg1.npz'), x=x, w_eq=w_eq, w_pol=w_pol, b_eq=b_eq, b_pol=b_pol,
                        xp=xp, y=y)


def check_padding_default_quirk():
    """custom.py:1073 declares padding=(1, 1), which keras ZeroPadding3D rejects; every call site passes an int."""
    try:
        ref.CubeSpherePadding2D()
    except ValueError:
        return True
    raise AssertionError('reference default padding unexpectedly accepted')


if __name__ == '__main__':
    check_padding_default_quirk()
    luts = make_pad_luts()
    cases = make_conv_cases()
    make_cfg1()
    print('wrote %d LUT arrays, %d conv arrays, cfg1' % (len(luts), len(cases)))
