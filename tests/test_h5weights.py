"""f4: reading the reference's saved models (DLWP/util.py:127-193 -> Keras HDF5) without h5py.

The reader (dlwp_cs_b200/h5weights.py) is checked against files produced by an independent writer of the same HDF5 subset
(tests/h5_writer.py) -- no HDF5 library exists in the image -- in the layout Keras uses:
``model_weights/<layer>/<layer>/<weight>:0`` with the layer objects' Keras names (creation order of
Azure/train_cs.py:209-228) and the ``add_weight`` names of custom.py:882-914."""
import numpy as np
import pytest

from dlwp_cs_b200 import h5weights
from dlwp_cs_b200.unet import CubeSphereCNN, CubeSphereUNet2, keras_layer_name
from tests import h5_writer


def keras_tree(model, rng, extra_optimizer=True):
    tree = {}
    for s in model.program:
        layer = getattr(model, s['name'])
        kname = keras_layer_name(s['name'])
        names = [nm for nm in ('equatorial_kernel', 'polar_kernel', 'north_pole_kernel', 'equatorial_bias', 'polar_bias',
                               'north_pole_bias') if getattr(layer, nm) is not None]
        tree[kname] = {kname: {nm + ':0': rng.standard_normal(tuple(getattr(layer, nm).shape)).astype(np.float32)
                               for nm in names}}
    root = {'model_weights': tree}
    if extra_optimizer:
        root['optimizer_weights'] = {'training': {'Adam': {'iter:0': np.array(7, dtype=np.int64)}}}
    return root


@pytest.mark.parametrize('flavour', ['earliest', 'earliest_split', 'earliest_userblock', 'latest'])
@pytest.mark.parametrize('arch,indep', [('unet2', False), ('unet3', True)])
def test_load_keras_weights_roundtrip(tmp_path, flavour, arch, indep):
    rng = np.random.default_rng(3)
    model = CubeSphereCNN(arch, 18, 14, base=8, independent_north_pole=indep)
    tree = keras_tree(model, rng)
    path = str(tmp_path / 'model.keras')
    if flavour == 'latest':
        h5_writer.write_latest(path, tree)
    else:
        h5_writer.write_earliest(path, tree, split_headers=flavour == 'earliest_split',
                                 userblock=512 if flavour == 'earliest_userblock' else 0)
    table = h5weights.read_keras_weights(path)
    assert 'optimizer_weights' not in table and len(table) == len(model.program)       # 16 layer groups -> several SNODs
    fresh = CubeSphereCNN(arch, 18, 14, base=8, independent_north_pole=indep)
    fresh.load_keras_weights(path)
    for s in model.program:
        kname, layer = keras_layer_name(s['name']), getattr(fresh, s['name'])
        for nm, arr in tree['model_weights'][kname][kname].items():
            np.testing.assert_array_equal(getattr(layer, nm[:-2]).detach().numpy(), arr)
    # == the Keras ordering of get_weights()
    flat = fresh.get_weights()
    per = 6 if indep else 4
    assert len(flat) == per * len(model.program)
    first = tree['model_weights']['cube_sphere_conv2d']['cube_sphere_conv2d']
    np.testing.assert_array_equal(flat[0], first['equatorial_kernel:0'])
    np.testing.assert_array_equal(flat[1], first['polar_kernel:0'])


def test_reader_dtypes_scalars_and_paths(tmp_path):
    tree = {'a': {'be64': np.arange(6, dtype='>f8').reshape(2, 3), 'i32': np.arange(5, dtype='<i4'), 'scalar': np.array(2.5, dtype='<f4')},
            'top': np.linspace(0, 1, 7, dtype='<f4')}
    for writer in (h5_writer.write_earliest, h5_writer.write_latest):
        path = str(tmp_path / 'x.h5')
        writer(path, tree)
        got = h5weights.read_datasets(path)
        assert sorted(got) == ['a/be64', 'a/i32', 'a/scalar', 'top']
        np.testing.assert_array_equal(got['a/be64'], np.arange(6, dtype=np.float64).reshape(2, 3))
        assert got['a/be64'].dtype == np.float64 and got['a/be64'].dtype.isnative
        np.testing.assert_array_equal(got['a/i32'], np.arange(5))
        assert got['a/scalar'].shape == () and float(got['a/scalar']) == 2.5
        np.testing.assert_array_equal(got['top'], tree['top'])


def test_reader_rejects_what_it_does_not_understand(tmp_path):
    p = tmp_path / 'junk.keras'
    p.write_bytes(b'PK\x03\x04 this is a zip (keras v3 format), not HDF5' + b'\x00' * 600)
    with pytest.raises(h5weights.H5FormatError):
        h5weights.read_datasets(str(p))
    # a chunked layout message (class 2) is refused with a clear message instead of returning garbage
    path = str(tmp_path / 'c.h5')
    h5_writer.write_earliest(path, {'w': np.zeros(4, dtype='<f4')})
    raw = bytearray(open(path, 'rb').read())
    i = raw.index(bytes([3, 1]) + (96).to_bytes(8, 'little'))          # layout v3, contiguous, data at offset 96
    raw[i + 1] = 2
    open(path, 'wb').write(bytes(raw))
    with pytest.raises(h5weights.H5FormatError, match='chunked'):
        h5weights.read_datasets(path)
    # a model whose layer names do not match
    model = CubeSphereUNet2(18, 14, base=8)
    h5_writer.write_earliest(path, {'model_weights': {'dense': {'dense': {'kernel:0': np.zeros((3, 3), dtype='<f4')}}}})
    with pytest.raises(KeyError):
        model.load_keras_weights(path)


def test_keras_layer_names_follow_creation_order():
    # Azure/train_cs.py:209-228: conv_2d_1, conv_2d_1_2, conv_2d_1_3, conv_2d_2, ... are created in this order, Keras
    # names them cube_sphere_conv2d, cube_sphere_conv2d_1, ...; conv_2d_8 is created with name='output'
    assert keras_layer_name('conv_2d_1') == 'cube_sphere_conv2d'
    assert keras_layer_name('conv_2d_1_2') == 'cube_sphere_conv2d_1'
    assert keras_layer_name('conv_2d_5') == 'cube_sphere_conv2d_10'
    assert keras_layer_name('conv_2d_5_2') == 'cube_sphere_conv2d_11'
    assert keras_layer_name('conv_2d_7_3') == 'cube_sphere_conv2d_18'
    assert keras_layer_name('conv_2d_8') == 'output'
