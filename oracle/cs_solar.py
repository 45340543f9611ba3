"""
CPU oracle (TEST INFRASTRUCTURE ONLY -- never imported by the product path) for the forced rollout of SURVEY.md section 8
row f2: the reference's closed-form insolation and the way TimeSeriesEstimator re-computes it every forecast iteration.

  * ``insolation``            restates DLWP/util.py:306-364 operation by operation, including its float32 day / longitude
                              arithmetic (util.py:340, 348) -- pinned bit-for-bit against the reference's own function
                              executed on the TensorFlow shim (tests/golden/insolation.npz, tests/test_oracle.py).
  * ``day_of_year``           util.py:301-303.
  * ``cubed_sphere_latlon``   an equiangular gnomonic cubed-sphere grid used as test geometry (the reference reads lat / lon
                              from its remapped dataset: extensions.py:282-284; any (6,N,N) lat / lon pair will do).
  * ``forced_forcing``        the insolation channels TimeSeriesEstimator.predict feeds iteration s
                              (extensions.py:272-288): time = t_sample + (s * T_out + n) * dt for input time step n.
"""
import numpy as np


def day_of_year(date):
    """Fractional days since 1 January of the date's year (util.py:301-303); `date` is a numpy datetime64."""
    date = np.datetime64(date, 's')
    year_start = date.astype('datetime64[Y]').astype('datetime64[s]')
    return (date - year_start) / np.timedelta64(1, 's') / 3600. / 24.


def insolation(days, lat, lon, S=1., daily=False):
    """
    util.py:306-364 with the dates already converted to day-of-year (`days`, any float array; the reference casts them to
    float32 at util.py:340).  lat / lon: same-shape arrays in degrees (lon 0-360).  Returns float32 (len(days),) + lat.shape.
    """
    lat = np.asarray(lat)
    lon = np.asarray(lon)
    if lat.ndim == 1:
        lon, lat = np.meshgrid(lon, lat)
    n_dim = lat.ndim
    eps = 23.4441 * np.pi / 180.
    ecc = 0.016715
    om = 282.7 * np.pi / 180.
    beta = np.sqrt(1 - ecc ** 2.)
    days_arr = np.asarray(days, dtype=np.float64).astype(np.float32).copy()
    for _ in range(n_dim):
        days_arr = np.expand_dims(days_arr, -1)
    if daily:
        days_arr = 0.5 + np.round(days_arr)
        new_lon = lon.copy().astype(np.float32)
        new_lon[:] = 0.
    else:
        new_lon = lon.astype(np.float32)
    lambda_m0 = ecc * (1. + beta) * np.sin(om)
    lambda_m = lambda_m0 + 2. * np.pi * (days_arr - 80.5) / 365.
    lambda_ = lambda_m + 2. * ecc * np.sin(lambda_m - om)
    dec = np.arcsin(np.sin(eps) * np.sin(lambda_))
    h = 2 * np.pi * (days_arr + new_lon / 360.)
    rho = (1. - ecc ** 2.) / (1. + ecc * np.cos(lambda_ - om))
    sol = S * (np.sin(np.pi / 180. * lat[None, ...]) * np.sin(dec) -
               np.cos(np.pi / 180. * lat[None, ...]) * np.cos(dec) * np.cos(h)) * rho ** -2.
    sol[sol < 0.] = 0.
    return sol.astype(np.float32)


def cubed_sphere_latlon(n):
    """Cell-centre latitude / longitude (degrees, lon in [0, 360)) of an equiangular cubed sphere, faces 0-3 around the
    equator, 4 = south pole, 5 = north pole (custom.py:759-763).  float64 (6, n, n)."""
    a = (np.arange(n) + 0.5) / n * (np.pi / 2) - np.pi / 4
    x, y = np.meshgrid(np.tan(a), np.tan(a))          # x: along the row (columns), y: down the rows
    lat = np.empty((6, n, n))
    lon = np.empty((6, n, n))
    for f in range(4):
        vx, vy, vz = np.ones_like(x), x, -y            # face centred on lon = f * 90
        lon0 = np.arctan2(vy, vx) + f * np.pi / 2
        lat[f] = np.arctan2(vz, np.hypot(vx, vy))
        lon[f] = lon0
    for f, s in ((4, -1.0), (5, 1.0)):
        vx, vy, vz = -s * y, x, s * np.ones_like(x)
        lat[f] = np.arctan2(vz, np.hypot(vx, vy))
        lon[f] = np.arctan2(vy, vx)
    return np.degrees(lat), np.mod(np.degrees(lon), 360.0)


def forced_forcing(day0, lat, lon, step, t_in, t_out, dt_days, S=1.):
    """Insolation channels of forecast iteration `step` (extensions.py:272-288 with one model step per iteration):
    for input time step n the time is t_sample + (step * t_out + n) * dt.  day0: (B,) day-of-year of every sample's
    first input time.  Returns float32 (B,) + lat.shape + (t_in,)."""
    chans = [insolation(np.asarray(day0, dtype=np.float64) + (step * t_out + n) * dt_days, lat, lon, S=S) for n in range(t_in)]
    return np.stack(chans, axis=-1)
