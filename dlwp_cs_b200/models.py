"""
Wrapper-level mirror of the reference's forecasting API for cubed-sphere models: ``predict`` / ``predict_timeseries`` of
``DLWPTorchNN`` (DLWP/model/models_torch.py:304-379), ``DLWPNeuralNet`` (DLWP/model/models.py:248-302) and
``DLWPFunctional`` (models.py:418-460) -- numpy in, numpy out, the reference's array layout -- over the device-resident
``RolloutEngine`` instead of one ``predict`` call (two host<->device copies plus a framework dispatch per layer) per step.

What the reference returns and this module reproduces exactly (models_torch.py:349-378 == models.py:272-301):

  * ``steps = ceil(time_steps / time_dim)`` model iterations unless ``step_sequence`` (then ``time_steps`` iterations);
  * ``time_series[t] = predict(p)``; ``p <- time_series[t]`` -- or, with ``step_sequence``, ``p <- (p[:, 1:], pr[:, [0]])``
    along the time axis of the input reshaped to ``(sample, time_dim, -1) + feature_shape[1:]`` (models.py:281-291);
  * the result reshaped to ``(steps, sample, time_dim, -1) + feature_shape[1:]`` and, unless ``keep_time_dim``,
    ``[:, :, 0]`` (step_sequence) or transposed ``(0, 2, 1, ...)`` and flattened to
    ``(steps * time_dim, sample, -1) + feature_shape[1:]``.

The reshape arithmetic is the reference's (it assumes the feature axis in front: ``# TODO: implement channels_last``,
models.py:455); it is applied to the series in the predictors' own layout, so for either ``data_format`` the array equals
what the reference's method returns for a model of that layout.  ``forecast_cs_layout`` gives the
``(f_hour, time, varlev, face, height, width)`` array ``verify.add_metadata_to_forecast_cs`` (verify.py:291-325) and
``CubeSphereRemap.convert_from_faces`` (remap/cubesphere.py:437-468) expect.
"""
import numpy as np
import torch

from . import _lib
from .unet import RolloutEngine


def assemble_timeseries(time_series, time_dim, step_sequence=False, keep_time_dim=False):
    """The output arithmetic of predict_timeseries (models_torch.py:372-378 / models.py:295-301) on the stacked model
    outputs ``time_series`` of shape (steps, sample) + feature_shape (non-recurrent model)."""
    time_series = np.asarray(time_series)
    time_steps, sample_dim = time_series.shape[:2]
    feature_shape = time_series.shape[2:]
    time_series = time_series.reshape((time_steps, sample_dim, time_dim, -1) + feature_shape[1:])
    if not keep_time_dim:
        if step_sequence:
            time_series = time_series[:, :, 0]
        else:
            time_series = time_series.transpose((0, 2, 1) + tuple(range(3, 3 + len(feature_shape))))
            time_series = time_series.reshape((time_steps * time_dim, sample_dim, -1) + feature_shape[1:])
    return time_series


def forecast_cs_layout(time_series, data_format='channels_last'):
    """(f_hour, time, ...) forecast with the variable axis in front of the faces -- the channels_first
    ``(f_hour, time, varlev, face, height, width)`` array of verify.py:291-325 -- from a series in either layout."""
    a = np.asarray(time_series)
    if data_format == 'channels_last':
        if a.ndim != 6:
            raise ValueError('expected (f_hour, time, 6, height, width, varlev), got %r' % (a.shape,))
        a = a.transpose(0, 1, 5, 2, 3, 4)
    if a.ndim != 6 or a.shape[3] != 6:
        raise ValueError('expected (f_hour, time, varlev, 6, height, width), got %r' % (a.shape,))
    return np.ascontiguousarray(a)


class CubeSphereForecaster(object):
    """
        fc = CubeSphereForecaster(model, time_dim=2, data_format='channels_first')
        series = fc.predict_timeseries(predictors, 40)            # == DLWPTorchNN.predict_timeseries(predictors, 40)

    model: a ``CubeSphereUNet2`` (channels_last layers) on a CUDA device whose output has the shape of its input
    (the plain rollout of models.py:248-302 feeds the output back unchanged; forced models go through
    ``RolloutEngine.set_solar`` / ``forcing=``).  predictors: float array ``(sample, C, 6, N, N)`` for
    data_format='channels_first' (the layout the reference's reshape assumes) or ``(sample, 6, N, N, C)``.
    dtype: torch.float32 (1e-5 parity: CUDA-core kernels, or with tensor_cores=True the float32-accurate split-bf16
    tensor-core path) or torch.bfloat16 (tensor-core kernels, bf16-stored activations).
    """

    def __init__(self, model, time_dim=1, data_format='channels_last', dtype=torch.float32, use_graph=True,
                 tensor_cores=False):
        if int(time_dim) < 1:
            raise ValueError("'time_dim' must be >= 1")
        if data_format not in ('channels_first', 'channels_last'):
            raise ValueError('The `data_format` argument must be one of "channels_first", "channels_last". '
                             'Received: %s' % data_format)
        self.model = model
        self.time_dim = int(time_dim)
        self.data_format = data_format
        self.dtype = dtype
        self.use_graph = use_graph
        self.tensor_cores = tensor_cores      # float32 tensors on the float32-accurate tensor-core path (RolloutEngine)
        self.is_recurrent = False
        self.is_convolutional = True
        self._engines = {}

    # ---- layout helpers ------------------------------------------------------------------------------------------------
    def _to_cl(self, a):
        return a.permute(0, 2, 3, 4, 1) if self.data_format == 'channels_first' else a

    def _series_from_cl(self, ring):           # (steps, B, 6, N, N, C) -> the predictors' layout
        return ring.permute(0, 1, 5, 2, 3, 4) if self.data_format == 'channels_first' else ring

    def _check(self, predictors, forcing):
        p = np.asarray(predictors)
        if p.ndim != 5:
            raise ValueError('expected 5-dimensional predictors, got shape %r' % (p.shape,))
        face_axis = 2 if self.data_format == 'channels_first' else 1
        if p.shape[face_axis] != 6:
            raise ValueError('expected 6 cube faces on axis %d, got shape %r' % (face_axis, p.shape))
        cf = 0
        if forcing is not None:
            forcing = np.asarray(forcing)
            cf = forcing.shape[1] if self.data_format == 'channels_first' else forcing.shape[-1]
        c = p.shape[1] if self.data_format == 'channels_first' else p.shape[-1]
        if c != self.model.out_channels or c + cf != self.model.in_channels:
            raise ValueError('predictors with %d channels (+%d forcing) do not match a model with %d inputs / %d outputs: '
                             'the rollout feeds the output back as the next input (models.py:450-453)'
                             % (c, cf, self.model.in_channels, self.model.out_channels))
        n = p.shape[3]
        return p, forcing, c, cf, n

    def _engine(self, batch, n, steps, cf):
        key = (batch, n, steps, cf)
        if key not in self._engines:
            self._engines[key] = RolloutEngine(self.model, batch, n, steps, forcing_channels=cf, dtype=self.dtype,
                                               use_graph=self.use_graph, tensor_cores=self.tensor_cores)
        return self._engines[key]

    def repack(self):
        """Call after the model's parameters changed (the engines hold packed copies of the weights)."""
        for e in self._engines.values():
            e.repack()

    # ---- the reference API ---------------------------------------------------------------------------------------------
    def predict(self, predictors, forcing=None):
        """models_torch.py:304-323: one model evaluation, numpy float32 in the predictors' layout."""
        p, forcing, c, cf, n = self._check(predictors, forcing)
        eng = self._engine(p.shape[0], n, 1, cf)
        dev = eng.device
        ring = eng.run(self._to_cl(torch.as_tensor(p, dtype=torch.float32).to(dev)),
                       None if forcing is None else self._to_cl(torch.as_tensor(forcing, dtype=torch.float32).to(dev)))
        out = self._series_from_cl(ring)[0].to(torch.float32)
        return out.cpu().numpy()

    def predict_timeseries(self, predictors, time_steps, step_sequence=False, keep_time_dim=False, verbose=0,
                           forcing=None):
        """models_torch.py:325-379 / models.py:248-302.  ``forcing`` (extension): constant extra input channels appended
        every step (what ``TimeSeriesEstimator`` supplies as insolation / constants, extensions.py:259-308)."""
        time_steps = int(time_steps)
        if time_steps < 1:
            raise ValueError("time_steps must be an int > 0")
        p, forcing, c, cf, n = self._check(predictors, forcing)
        if c % self.time_dim:
            raise ValueError('%d channels cannot be split into time_dim = %d time steps' % (c, self.time_dim))
        if not step_sequence:
            time_steps = int(np.ceil(1. * time_steps / self.time_dim))
        batch = p.shape[0]
        t32 = lambda a: self._to_cl(torch.as_tensor(np.asarray(a), dtype=torch.float32))
        if not step_sequence:
            eng = self._engine(batch, n, time_steps, cf)
            ring = eng.run(t32(p).to(eng.device), None if forcing is None else t32(forcing).to(eng.device))
        else:
            # one predicted time step at a time (models.py:281-291): the next input is the last time_dim - 1 input time
            # steps followed by the first predicted one.  The window slides on the device; no host hop between steps.
            eng = self._engine(batch, n, 1, cf)
            dev = eng.device
            f = None if forcing is None else t32(forcing).to(dev)
            cur = t32(p).to(dev)
            ring = torch.empty((time_steps,) + tuple(cur.shape), dtype=torch.float32, device=dev)
            v = c // self.time_dim
            for t in range(time_steps):
                pr = eng.run(cur, f)[0].to(torch.float32)
                ring[t] = pr
                # channels are time-major (c = t * V + v) in both layouts, so the window shifts by V channels
                cur = torch.cat([cur[..., v:], pr[..., :v]], dim=-1)
        series = self._series_from_cl(ring).to(torch.float32).cpu().numpy()
        return assemble_timeseries(series, self.time_dim, step_sequence, keep_time_dim)

    def forecast(self, predictors, time_steps, forcing=None):
        """``predict_timeseries`` delivered as ``(f_hour, time, varlev, face, height, width)`` (verify.py:291-325)."""
        return forecast_cs_layout(self.predict_timeseries(predictors, time_steps, forcing=forcing), self.data_format)


def require_library():
    """Raise (ImportError) unless libdlwpcs.so is built -- there is no CPU path behind this module."""
    return _lib.load()
