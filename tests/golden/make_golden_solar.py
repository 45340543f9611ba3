"""
Golden vectors for the insolation of the forced rollout, produced by the REFERENCE'S OWN
`DLWP.util.insolation` (util.py:306-364, pandas Timestamps and all) imported on the TensorFlow shim.  Build container only:

    python tests/golden/make_golden_solar.py        # -> tests/golden/insolation.npz (committed)
"""
import os
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..', 'oracle'))
import tf_shim  # noqa: E402
import cs_solar  # noqa: E402

ref = tf_shim.load_reference_util()
dates = pd.to_datetime(['2013-01-01T00:00', '2013-03-21T06:00', '2016-02-29T18:00', '2016-06-21T12:00', '2017-09-23T03:00',
                        '2018-12-31T21:00'])
out = {'dates': np.array([str(d) for d in dates]),
       'days': np.array([ref.day_of_year(d) for d in dates], dtype=np.float64)}
for n in (4, 12):
    lat, lon = cs_solar.cubed_sphere_latlon(n)
    out['lat_%d' % n], out['lon_%d' % n] = lat, lon
    out['sol_%d' % n] = ref.insolation(dates, lat.reshape(6 * n, n), lon.reshape(6 * n, n)).reshape(len(dates), 6, n, n)
    out['sol_daily_%d' % n] = ref.insolation(dates, lat.reshape(6 * n, n), lon.reshape(6 * n, n), S=2.5,
                                             daily=True).reshape(len(dates), 6, n, n)
lat1, lon1 = np.linspace(-87.5, 87.5, 8), np.arange(0., 360., 30.)
out['lat_1d'], out['lon_1d'] = lat1, lon1
out['sol_1d'] = ref.insolation(dates, lat1, lon1)
np.savez_compressed(os.path.join(HERE, 'insolation.npz'), **out)
print({k: v.shape for k, v in out.items()})
