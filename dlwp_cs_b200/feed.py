"""
Device-side data feed: the batch assembly of the reference's ``ArrayDataGenerator`` (DLWP/model/generators.py:636-1011)
with the training array resident in HBM (C48, 7 variables, 40 years of 6-hourly data: 11 GB as float32 -- a small part of
the 180 GB), one transposing-gather launch per batch (``dlwpcs_feed_gather``).  Covers the configuration of the
cubed-sphere training scripts (Azure/train_cs.py:157-163): convolutional model, channels_last, insolation and constants,
``sequence=None``.

    feed = DeviceDataFeed(array, input_slice=slice(0, 7), output_slice=slice(0, 7), input_time_steps=2,
                          output_time_steps=2, interval=2, insolation_array=sol, constants=consts, dtype=torch.bfloat16)
    x, y = feed.generate(samples)          # (B,6,N,N,Cx), (B,6,N,N,Cy) on the device; x already holds [p | constants]
    x, y = feed[i]                         # batch i of the current epoch order (``on_epoch_end`` reshuffles)
"""
import numpy as np
import torch

from . import _lib


def _indices(sel, n):
    """Variable indices of a slice / index list over the varlev axis, negative indices normalised, range-checked (the
    gather kernel reads array[t, idx] unchecked)."""
    if sel is None:
        sel = slice(None)
    if isinstance(sel, slice):
        return list(range(*sel.indices(n)))
    out = []
    for v in sel:
        v = int(v)
        if v < -n or v >= n:
            raise IndexError('variable index %d out of range for %d variables' % (v, n))
        out.append(v % n)
    return out


class DeviceDataFeed(object):
    def __init__(self, array, batch_size=32, input_slice=None, output_slice=None, input_time_steps=1,
                 output_time_steps=1, interval=1, shuffle=False, insolation_array=None, constants=None,
                 drop_remainder=False, dtype=torch.float32, device='cuda', seed=0, sequence=None, rank=0, world=1):
        for v, nm in ((input_time_steps, 'input_time_steps'), (output_time_steps, 'output_time_steps'),
                      (batch_size, 'batch_size'), (interval, 'interval')):
            if int(v) <= 0:
                raise ValueError('%s must be positive' % nm)          # the asserts of generators.py:675-679
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise _lib.DlwpcsError('DeviceDataFeed keeps the data on a CUDA device (no CPU path)')
        array = torch.as_tensor(array)
        if array.dim() != 5 or array.shape[2] != 6 or array.shape[3] != array.shape[4]:
            raise ValueError('array must be (time, varlev, 6, N, N), got %r' % (tuple(array.shape),))
        self.array = array.to(self.device, torch.float32).contiguous()
        self.n = int(array.shape[3])
        self.npix = 6 * self.n * self.n
        self.n_var = int(array.shape[1])
        self.t_in, self.t_out, self.interval = int(input_time_steps), int(output_time_steps), int(interval)
        self.in_idx, self.out_idx = _indices(input_slice, self.n_var), _indices(output_slice, self.n_var)
        self._in_vars = torch.tensor(self.in_idx, dtype=torch.int32, device=self.device)
        self._out_vars = torch.tensor(self.out_idx, dtype=torch.int32, device=self.device)
        self.insolation_array = None
        if insolation_array is not None:
            sol = torch.as_tensor(insolation_array)
            if tuple(sol.shape) != (array.shape[0],) + tuple(array.shape[2:]):
                raise ValueError('spatial dimensions of insolation must be the same as input data; got %r and %r'
                                 % (tuple(sol.shape[1:]), tuple(array.shape[2:])))
            self.insolation_array = sol.to(self.device, torch.float32).contiguous()
        self.constants = None
        if constants is not None:
            c = torch.as_tensor(constants)
            if tuple(c.shape[1:]) != tuple(array.shape[2:]):
                raise ValueError('spatial dimensions of constants must be the same as input data; got %r and %r'
                                 % (tuple(c.shape[1:]), tuple(array.shape[2:])))
            self.constants = c.to(self.device, torch.float32).contiguous()
        self.dtype = dtype
        self.batch_size, self.shuffle, self.drop_remainder = int(batch_size), bool(shuffle), bool(drop_remainder)
        self.sequence = None if sequence is None else int(sequence)
        if self.sequence is not None and self.sequence <= 0:
            raise ValueError('sequence must be positive')
        if self.sequence is not None and self.insolation_array is None:
            raise ValueError('sequence mode hands the insolation of the later steps to the model: insolation_array is required')
        # generators.py:692-695
        self._n_sample = int(array.shape[0]) - self.interval * (self.t_in + self.t_out * (self.sequence or 1)) + 1
        if self._n_sample <= 0:
            raise ValueError('the array is too short for %d + %d time steps at interval %d' % (self.t_in, self.t_out,
                                                                                                self.interval))
        # data parallel: every rank shuffles with the SAME seed and takes every world-th sample of the epoch order, so the
        # ranks see disjoint shards of one global batch (generators.py shuffles once per process; Keras multi_gpu_model
        # splits each batch across the GPUs, models.py:105-110)
        self.rank, self.world = int(rank), int(world)
        if not 0 <= self.rank < self.world:
            raise ValueError('rank %d outside world of %d' % (self.rank, self.world))
        self._rng = np.random.RandomState(seed)
        self.on_epoch_end()

    # ---- shapes (generators.py:739-865, channels_last) ---------------------------------------------------------------
    @property
    def n_input_channels(self):
        return self.t_in * (len(self.in_idx) + (1 if self.insolation_array is not None else 0)) + \
            (0 if self.constants is None else int(self.constants.shape[0]))

    @property
    def n_output_channels(self):
        return self.t_out * len(self.out_idx)

    def on_epoch_end(self):
        self._indices = np.arange(self._n_sample)
        if self.shuffle:
            self._rng.shuffle(self._indices)
        if self.world > 1:           # equal shards: drop the tail that does not divide by the world size
            usable = (self._n_sample // self.world) * self.world
            self._indices = self._indices[:usable][self.rank::self.world]

    def __len__(self):                      # generators.py:986-993 (per rank: batch_size samples of this rank's shard)
        n = len(self._indices)
        if self.drop_remainder:
            return n // self.batch_size
        return int(np.ceil(n / self.batch_size))

    def __getitem__(self, index):           # generators.py:995-1011
        if index < 0 or index >= len(self):
            raise IndexError('batch index %d out of range' % index)
        return self.generate(self._indices[index * self.batch_size:(index + 1) * self.batch_size])

    def _gather(self, s_dev, want_x, want_y):
        b = int(s_dev.numel())
        x = torch.empty((b, 6, self.n, self.n, self.n_input_channels), dtype=self.dtype, device=self.device) if want_x else None
        y = torch.empty((b, 6, self.n, self.n, self.n_output_channels), dtype=self.dtype, device=self.device) if want_y else None
        _lib.check(_lib.load().dlwpcs_feed_gather(
            _lib.ptr(self.array), _lib.ptr(self.insolation_array), _lib.ptr(self.constants), _lib.ptr(s_dev),
            _lib.ptr(self._in_vars), _lib.ptr(self._out_vars), _lib.ptr(x), _lib.ptr(y), b, self.npix, self.n_var,
            len(self.in_idx), len(self.out_idx), self.t_in, self.t_out, self.interval,
            0 if self.constants is None else int(self.constants.shape[0]), _lib.dtype_code(self.dtype), _lib.stream_ptr()))
        return x, y

    def generate(self, samples):
        """-> (x, y); with ``sequence=S``: (x, [solar_1 .. solar_{S-1}], [y_0 .. y_{S-1}]) -- the inputs the reference's
        generator hands to the multi-step Keras model (generators.py:884-893, 964-983) with the constants already appended
        to x: solar_s (B, T_in, 6, N, N, 1) is the insolation at the input times of step s, y_s the targets of step s."""
        if len(samples) == 0:
            samples = np.arange(self._n_sample)
        samples = np.asarray(samples, dtype=np.int64)
        if samples.min() < 0 or samples.max() >= self._n_sample:
            raise IndexError('sample index out of range [0, %d)' % self._n_sample)
        s_dev = torch.from_numpy(samples).to(self.device, non_blocking=True)
        x, y = self._gather(s_dev, True, True)
        if self.sequence is None:
            return x, y
        solars, targets = [], [y]
        for s in range(1, self.sequence):
            sol = torch.stack([self.insolation_array[s_dev + self.interval * (self.t_in * s + k)] for k in range(self.t_in)], dim=1)
            solars.append(sol.unsqueeze(-1).to(self.dtype))
            targets.append(self._gather(s_dev + self.interval * self.t_out * s, False, True)[1])
        return x, solars, targets
