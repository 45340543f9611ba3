"""
Golden vectors for the device-side data feed, produced by the REFERENCE'S OWN
``ArrayDataGenerator.generate`` (DLWP/model/generators.py:636-984) imported on the shim.  Build container only:

    python tests/golden/make_golden_feed.py        # -> tests/golden/feed.npz (committed)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..', 'oracle'))
import tf_shim  # noqa: E402

gen = tf_shim.load_reference_generators()


class ModelStub(object):                  # "instance of a DLWP model, just used for some metadata" (generators.py:651)
    is_convolutional, is_recurrent, impute = True, False, False


rng = np.random.default_rng(7)
T, V, N = 14, 5, 4
array = rng.standard_normal((T, V, 6, N, N)).astype(np.float32)
insol = rng.random((T, 6, N, N)).astype(np.float32)
consts = rng.random((2, 6, N, N)).astype(np.float32)
out = {'array': array, 'insolation_array': insol, 'constants': consts}
# variable selections are slices, as DLWP.model.preprocessing.prepare_data_array returns them (an index list would be
# broadcast against the sample indices by numpy's advanced indexing at generators.py:880)
cases = {'a': dict(input_slice=slice(0, 4), output_slice=slice(0, 3), input_time_steps=2, output_time_steps=2, interval=2),
         'b': dict(input_slice=None, output_slice=None, input_time_steps=2, output_time_steps=2, interval=1),
         'c': dict(input_slice=slice(1, 5, 2), output_slice=slice(3, 4), input_time_steps=1, output_time_steps=3, interval=3)}
samples = {'a': [0, 3, 1, 6], 'b': [9, 0, 10], 'c': [2, 1, 0, 1]}
for k, kw in cases.items():
    g = gen.ArrayDataGenerator(ModelStub(), array, rank=3, batch_size=4, insolation_array=insol, constants=consts,
                               channels_last=True, **kw)
    p, t = g.generate(samples[k])
    assert isinstance(p, list) and len(p) == 2
    out['p_%s' % k], out['const_%s' % k], out['t_%s' % k] = p[0], p[1], t
    out['samples_%s' % k] = np.array(samples[k])
    out['n_sample_%s' % k] = np.array(g._n_sample)
# sequence mode (Azure/train_cs.py:159-163 passes sequence=integration_steps): inputs [p, solar_1.., constants], list targets
seq_cases = {'d': (dict(input_slice=slice(0, 4), output_slice=slice(0, 4), input_time_steps=2, output_time_steps=2, interval=1), 2,
                   [0, 5, 2]),
             'e': (dict(input_slice=None, output_slice=slice(1, 3), input_time_steps=1, output_time_steps=2, interval=2), 3,
                   [0, 0])}
for k, (kw, S, smp) in seq_cases.items():
    g = gen.ArrayDataGenerator(ModelStub(), array, rank=3, batch_size=4, insolation_array=insol, constants=consts,
                               channels_last=True, sequence=S, **kw)
    p, t = g.generate(smp)
    assert isinstance(p, list) and len(p) == S + 1 and isinstance(t, list) and len(t) == S
    out['p_%s' % k], out['const_%s' % k] = p[0], p[-1]
    for s in range(1, S):
        out['solar_%s_%d' % (k, s)] = p[s]
    for s in range(S):
        out['t_%s_%d' % (k, s)] = t[s]
    out['samples_%s' % k] = np.array(smp)
    out['n_sample_%s' % k] = np.array(g._n_sample)
np.savez_compressed(os.path.join(HERE, 'feed.npz'), **out)
print({k: v.shape for k, v in out.items()})
