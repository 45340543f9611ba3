"""
The Weyn-2020 cubed-sphere U-Net (``unet2`` of the reference's Azure/train_cs.py:196-228, 277-305) on the engine's
layers, and a device-resident autoregressive rollout (the loop of DLWP/model/models.py:446-454 without the two
host<->device copies per step).

  * ``CubeSphereUNet2``  -- nn.Module; training / generic forward through the differentiable operators.
  * ``RolloutEngine``    -- inference: one kernel launch per CubeSphereConv2D (11 per step); the halo exchange, the
    average pooling, the nearest up-sampling, the channel concatenation, bias and the capped leaky ReLU all live in that
    kernel's load stage / epilogue; state, forcing and the forecast ring stay in HBM; the whole multi-step rollout is
    replayed from one CUDA graph.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .custom import CubeSphereConv2D
from . import functional as F_cs

RELU = ('capped_leaky_relu', 0.1, 10.0)      # keras ReLU(negative_slope=0.1, max_value=10.) train_cs.py:199


def unet2_layer_specs(in_channels, out_channels, base=32):
    """(name, kernel, Cin, Cout) in execution order -- train_cs.py:209-228 + 277-305."""
    b = base
    return [('conv_2d_1', 3, in_channels, b), ('conv_2d_1_2', 3, b, b), ('conv_2d_2', 3, b, 2 * b),
            ('conv_2d_2_2', 3, 2 * b, 2 * b), ('conv_2d_5_2', 3, 2 * b, 4 * b), ('conv_2d_5', 3, 4 * b, 2 * b),
            ('conv_2d_6_2', 3, 4 * b, 2 * b), ('conv_2d_6', 3, 2 * b, b), ('conv_2d_7', 3, 2 * b, b),
            ('conv_2d_7_2', 3, b, b), ('conv_2d_8', 1, b, out_channels)]


def _avg_pool(x):      # AveragePooling3D((1,2,2)), channels_last
    b, f, h, w, c = x.shape
    return x.reshape(b, f, h // 2, 2, w // 2, 2, c).mean(dim=(3, 5))


def _upsample(x):      # UpSampling3D((1,2,2))
    return x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)


class CubeSphereUNet2(nn.Module):
    """``unet2``: 10 halo-padded 3x3 CubeSphereConv2D + capped leaky ReLU, two 2x2 poolings with skip connections, and a
    1x1 output convolution.  channels_last (B,6,N,N,C); N divisible by 4."""

    def __init__(self, in_channels, out_channels, base=32, independent_north_pole=False):
        super().__init__()
        self.in_channels, self.out_channels, self.base = in_channels, out_channels, base
        self.independent_north_pole = independent_north_pole
        for name, k, ci, co in unet2_layer_specs(in_channels, out_channels, base):
            last = name == 'conv_2d_8'
            setattr(self, name, CubeSphereConv2D(
                co, k, padding='valid', data_format='channels_last', dilation_rate=1,
                activation='linear' if last else RELU, independent_north_pole=independent_north_pole,
                flip_north_pole=not independent_north_pole,      # train_cs.py:205-206
                in_channels=ci, fuse_padding=0 if last else 1, name='output' if last else name))

    def layers_in_order(self):
        return [getattr(self, s[0]) for s in unet2_layer_specs(self.in_channels, self.out_channels, self.base)]

    def forward(self, x):
        # pooling / upsampling + concatenation: one 16-byte kernel each way when the channel counts allow it (they do for
        # every base that is a multiple of 8), torch ops otherwise
        def pool(t):
            return F_cs.avg_pool_2x2(t) if _lib.resample_vec_ok(t) else _avg_pool(t)

        def upcat(a, b):
            return F_cs.upsample_concat(a, b) if _lib.resample_vec_ok(a, b) else torch.cat([_upsample(a), b], dim=-1)

        x0 = self.conv_2d_1_2(self.conv_2d_1(x))
        x1 = self.conv_2d_2_2(self.conv_2d_2(pool(x0)))
        x2 = self.conv_2d_5(self.conv_2d_5_2(pool(x1)))
        t = self.conv_2d_6(self.conv_2d_6_2(upcat(x2, x1)))
        t = self.conv_2d_7_2(self.conv_2d_7(upcat(t, x0)))
        return self.conv_2d_8(t)

    # ---- weight interop with the reference's Keras model ----------------------------------------
    def get_weights(self):
        """numpy arrays in the order ``keras.Model.get_weights()`` returns them for the reference's ``unet2`` built by
        Azure/train_cs.py:277-305: layers in creation order (conv_2d_1, conv_2d_1_2, ..., output), each layer in its
        ``add_weight`` order (custom.py:882-914: equatorial_kernel, polar_kernel, [north_pole_kernel,] equatorial_bias,
        polar_bias[, north_pole_bias]); kernels HWIO."""
        out = []
        for layer in self.layers_in_order():
            out += layer.get_weights()
        return out

    def set_weights(self, weights):
        """Inverse of ``get_weights``: takes the list a reference checkpoint yields (``model.get_weights()`` /
        the arrays of its h5 file in the same order)."""
        weights = list(weights)
        per = 6 if self.independent_north_pole else 4
        layers = self.layers_in_order()
        if len(weights) != per * len(layers):
            raise ValueError('expected %d weight arrays (%d layers x %d), got %d' % (per * len(layers), len(layers), per,
                                                                                      len(weights)))
        for i, layer in enumerate(layers):
            chunk = weights[per * i:per * (i + 1)]
            want = (layer.kernel_size[0], layer.kernel_size[1], layer.equatorial_kernel.shape[2], layer.filters)
            if tuple(chunk[0].shape) != want:
                raise ValueError('layer %d: kernel shape %r, expected %r (HWIO)' % (i, tuple(chunk[0].shape), want))
            layer.set_weights(chunk)

    def load_oracle_params(self, params):
        """params: dict 'layer.equatorial_kernel' -> tensor, as produced by oracle.make_unet2_params (tests / bench)."""
        with torch.no_grad():
            for name, p in self.named_parameters():
                p.copy_(params[name].to(p.dtype))


def reference_input_order(t_in, n_var, n_const, n_sol_per_step=1):
    """Engine slots of the reference's input channel packing (generators.py:880-899, train_cs.py:396-407): per input time
    step the variables then the insolation, constants appended -- for ``RolloutEngine(input_order=...)``.  Prognostic slot
    t * n_var + v, forcing slots: insolation of time step t first, then the constants."""
    cp = t_in * n_var
    order = []
    for t in range(t_in):
        order += [t * n_var + v for v in range(n_var)]
        order += [cp + t * n_sol_per_step + k for k in range(n_sol_per_step)]
    order += [cp + t_in * n_sol_per_step + k for k in range(n_const)]
    return order


class RolloutEngine(object):
    """
    Device-resident forecast loop for a ``CubeSphereUNet2``.

        eng = RolloutEngine(model, batch, n, steps, forcing_channels=4, dtype=torch.bfloat16)
        ring = eng.run(state, forcing)            # (steps, B, 6, N, N, Cout) on the device

    state: (B,6,N,N,Cout) prognostic channels; forcing: (B,6,N,N,Cf) or per-step (steps,B,6,N,N,Cf) channels appended to
    the network input each step (insolation / constants, train_cs.py:396-407); the output of step t is the prognostic
    input of step t+1 (models.py:450-453).
    """

    def __init__(self, model, batch, n, steps, forcing_channels=0, dtype=torch.float32, use_graph=True,
                 per_step_forcing=False, device=None, input_order=None):
        if n % 4 != 0:
            raise ValueError('unet2 pools twice: face edge must be divisible by 4')
        self.model, self.batch, self.n, self.steps = model, batch, n, steps
        self.cf = forcing_channels
        self.cp = model.out_channels
        if self.cp + self.cf != model.in_channels:
            raise ValueError('prognostic (%d) + forcing (%d) channels != model input channels (%d)'
                             % (self.cp, self.cf, model.in_channels))
        self.dtype = dtype
        # input_order[i] = engine slot read by the model's i-th input channel: k < Cp is prognostic channel k (fed back from
        # the output), Cp + k is forcing channel k.  Default: prognostic channels first (see reference_input_order).
        self.input_order = list(range(model.in_channels)) if input_order is None else [int(k) for k in input_order]
        if sorted(self.input_order) != list(range(model.in_channels)):
            raise ValueError('input_order must be a permutation of range(%d)' % model.in_channels)
        self.solar = None
        # channel counts padded to multiples of 8 so that every gather is a 16-byte copy: the pad channels carry zero
        # weights on the way in and are written as exact zeros by the output layer
        self.cp_pad = -(-self.cp // 8) * 8
        self.cf_pad = -(-self.cf // 8) * 8 if self.cf else 0
        self.device = device or next(model.parameters()).device
        if self.device.type != 'cuda':
            raise _lib.DlwpcsError('RolloutEngine needs the model on a CUDA device')
        self.use_graph = use_graph
        self.per_step_forcing = per_step_forcing
        dt = _lib.dtype_code(dtype)
        b, base = batch, model.base
        mk = lambda edge, c: torch.empty((b, 6, edge, edge, c), dtype=dtype, device=self.device)
        self.state = torch.zeros((b, 6, n, n, self.cp_pad), dtype=dtype, device=self.device)
        fshape = ((steps,) if per_step_forcing else ()) + (b, 6, n, n, max(self.cf_pad, 8))
        self.forcing = torch.zeros(fshape, dtype=dtype, device=self.device)
        self.ring = torch.empty((steps, b, 6, n, n, self.cp_pad), dtype=dtype, device=self.device)
        self.buf = dict(a=mk(n, base), x0=mk(n, base), b=mk(n // 2, 2 * base), x1=mk(n // 2, 2 * base),
                        c=mk(n // 4, 4 * base), x2=mk(n // 4, 2 * base), d=mk(n // 2, 2 * base), e=mk(n // 2, base),
                        f=mk(n, base), g=mk(n, base))
        S, P, U = _lib.SRC_SAME, _lib.SRC_POOL2, _lib.SRC_UP2
        # (layer, edge, src0, c0, mode0, src1, c1, mode1, dst)
        plan = [('conv_2d_1', n, 'state', self.cp_pad, S, 'forcing' if self.cf else None, self.cf_pad, S, 'a'),
                ('conv_2d_1_2', n, 'a', base, S, None, 0, S, 'x0'),
                ('conv_2d_2', n // 2, 'x0', base, P, None, 0, S, 'b'),
                ('conv_2d_2_2', n // 2, 'b', 2 * base, S, None, 0, S, 'x1'),
                ('conv_2d_5_2', n // 4, 'x1', 2 * base, P, None, 0, S, 'c'),
                ('conv_2d_5', n // 4, 'c', 4 * base, S, None, 0, S, 'x2'),
                ('conv_2d_6_2', n // 2, 'x2', 2 * base, U, 'x1', 2 * base, S, 'd'),
                ('conv_2d_6', n // 2, 'd', 2 * base, S, None, 0, S, 'e'),
                ('conv_2d_7', n, 'e', base, U, 'x0', base, S, 'f'),
                ('conv_2d_7_2', n, 'f', base, S, None, 0, S, 'g'),
                ('conv_2d_8', n, 'g', base, S, None, 0, S, 'out')]
        self.plan = []
        for name, edge, s0, c0, m0, s1, c1, m1, dst in plan:
            layer = getattr(model, name)
            k = layer.kernel_size
            fused = F_cs.resolve_activation(layer.activation)
            cout = self.cp_pad if dst == 'out' else layer.filters
            d = _lib.make_desc(b, edge, c0 + c1, cout, k, (1, 1), (1, 1), layer.fuse_padding, False,
                               layer.flip_north_pole, layer.independent_north_pole, layer.use_bias, fused[0], fused[1],
                               fused[2], dt, dt, c0, m0, c1, m1)
            self.plan.append([name, d, s0, s1, dst, None])
        self.graph = None
        self.graph_host = None
        self.host_ring = None
        self.host_chunk = 0
        self.repack()

    def repack(self):
        """(Re)pack the layer weights into the kernels' layouts -- call after the model's parameters change."""
        for item in self.plan:
            layer = getattr(self.model, item[0])
            ws = [layer.equatorial_kernel, layer.polar_kernel, layer.north_pole_kernel]
            bs = [layer.equatorial_bias, layer.polar_bias, layer.north_pole_bias]
            if item[0] == 'conv_2d_1':          # input channels: [prognostic | pad | forcing | pad]
                ws = [None if w is None else self._pad_in(w.detach()) for w in ws]
            if item[4] == 'out':                # output channels: [prognostic | pad]
                ws = [None if w is None else F.pad(w.detach(), (0, self.cp_pad - self.cp)) for w in ws]
                bs = [None if v is None else F.pad(v.detach(), (0, self.cp_pad - self.cp)) for v in bs]
            item[5] = _lib.pack_weights(item[1], ws[0], ws[1], ws[2], bs[0], bs[1], bs[2])
        self.graph = None
        self.graph_host = None

    def _pad_in(self, w):
        kh, kw, _, co = w.shape
        out = w.new_zeros((kh, kw, self.cp_pad + self.cf_pad, co))
        for i, k in enumerate(self.input_order):     # model input channel i reads engine slot k
            out[:, :, k if k < self.cp else self.cp_pad + (k - self.cp)] = w[:, :, i]
        return out

    @property
    def launches_per_step(self):
        return len(self.plan)

    def _src(self, key, t):
        if key is None:
            return None
        if key == 'state':
            return self.state if t == 0 else self.ring[t - 1]
        if key == 'forcing':
            return self.forcing[t] if self.per_step_forcing else self.forcing
        return self.buf[key]

    def set_solar(self, lat, lon, dt_days, t_in, t_out, day0=None, start_dates=None, S=1.0):
        """
        Forced rollout (TimeSeriesEstimator.predict, extensions.py:259-308): before forecast iteration s the first `t_in`
        forcing channels are overwritten on the device with the insolation (util.py:306-364) at
        t_sample + (s * t_out + n) * dt, n = 0..t_in-1; the remaining forcing channels (constants) stay as loaded.

        lat, lon: (6,N,N) degrees (lon 0-360).  Either `day0` ((B,) day of year of each member's first input time; no
        new-year wrap) or `start_dates` ((B,) numpy datetime64; the day of year restarts on 1 January like util.py:301).
        """
        import numpy as np
        if self.per_step_forcing:
            raise ValueError('set_solar computes the forcing on the device; build the engine without per_step_forcing')
        if t_in > self.cf:
            raise ValueError('%d insolation channels do not fit %d forcing channels' % (t_in, self.cf))
        lat = np.asarray(lat, dtype=np.float64).reshape(-1)
        lon = np.asarray(lon, dtype=np.float64).reshape(-1)
        if lat.size != 6 * self.n * self.n or lon.size != lat.size:
            raise ValueError('lat / lon must have 6 x %d x %d entries' % (self.n, self.n))
        days = np.empty((self.steps, t_in, self.batch), dtype=np.float64)
        for s in range(self.steps):
            for k in range(t_in):
                off = (s * t_out + k) * dt_days
                if start_dates is not None:
                    d = np.asarray(start_dates, dtype='datetime64[s]') + np.timedelta64(int(round(off * 86400)), 's')
                    y0 = d.astype('datetime64[Y]').astype('datetime64[s]')
                    days[s, k] = (d - y0) / np.timedelta64(1, 's') / 3600. / 24.
                else:
                    days[s, k] = np.asarray(day0, dtype=np.float64) + off
        dev = self.device
        self.solar = dict(n=t_in, S=float(S), days=torch.from_numpy(days).to(dev),
                          sinlat=torch.from_numpy(np.sin(np.pi / 180. * lat)).to(dev),
                          coslat=torch.from_numpy(np.cos(np.pi / 180. * lat)).to(dev),
                          lon=torch.from_numpy(lon.astype(np.float32)).to(dev))
        self.graph = None
        self.graph_host = None

    def _step(self, t):
        if self.solar is not None:
            so = self.solar
            _lib.insolation(self.forcing, 0, so['n'], so['sinlat'], so['coslat'], so['lon'], so['days'][t], so['S'])
        for name, d, s0, s1, dst, packed in self.plan:
            out = self.ring[t] if dst == 'out' else self.buf[dst]
            _lib.conv2d_fwd(d, self._src(s0, t), self._src(s1, t), packed, out=out)

    def _enqueue_all(self):
        for t in range(self.steps):
            self._step(t)

    def _ensure_graph(self):
        if self.graph is not None or not self.use_graph:
            return
        # warm the halo-table caches and kernel attributes outside the capture
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            self._step(0)
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._enqueue_all()
        self.graph = g

    def load_inputs(self, state, forcing=None, non_blocking=True):
        """Copy initial conditions (host or device tensors, any float dtype) into the engine's resident buffers."""
        self.state[..., :self.cp].copy_(state, non_blocking=non_blocking)
        if self.cf:
            if forcing is None:
                raise ValueError('this engine was built with %d forcing channels' % self.cf)
            self.forcing[..., :self.cf].copy_(forcing, non_blocking=non_blocking)

    def launch(self):
        """Enqueue the whole rollout on the current stream (graph replay when enabled); no host synchronisation."""
        if self.use_graph:
            self._ensure_graph()
            self.graph.replay()
        else:
            self._enqueue_all()
        return self.forecast

    @property
    def forecast(self):
        """(steps, B, 6, N, N, Cout) view of the resident forecast ring without the pad channels."""
        return self.ring[..., :self.cp]

    def run(self, state, forcing=None):
        self.load_inputs(state, forcing)
        return self.launch()

    # ---- forecast delivered to the host while the rollout is still running ------------------------------------------
    def _enqueue_all_to_host(self):
        """The rollout with the forecast ring streamed to pinned host memory in chunks of `host_chunk` model steps on a
        second stream: the PCIe transfer of steps [t0, t1) overlaps the computation of the following steps (the
        reference copies every step's result to the host before it starts the next one, models.py:446-454)."""
        main = torch.cuda.current_stream(self.device)
        side = self._side
        for t0 in range(0, self.steps, self.host_chunk):
            t1 = min(self.steps, t0 + self.host_chunk)
            for t in range(t0, t1):
                self._step(t)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                stage = self._stage[(t0 // self.host_chunk) % 2]
                stage[:t1 - t0].copy_(self.ring[t0:t1, ..., :self.cp])            # strip the pad channels on the device
                self.host_ring[t0:t1].copy_(stage[:t1 - t0], non_blocking=True)
        main.wait_stream(side)

    def run_to_host(self, state, forcing=None, chunk=5):
        """Initial conditions from (pinned) host tensors in, whole forecast (steps,B,6,N,N,Cout) out in pinned host memory
        owned by the engine.  Everything is enqueued on the current stream; synchronise before reading the result."""
        if self.host_ring is None or self.host_chunk != chunk:
            self.host_chunk = chunk
            shape = (self.steps, self.batch, 6, self.n, self.n, self.cp)
            self.host_ring = torch.empty(shape, dtype=self.dtype).pin_memory()
            self._stage = [torch.empty((chunk,) + shape[1:], dtype=self.dtype, device=self.device) for _ in range(2)]
            self._side = torch.cuda.Stream(device=self.device)
            self.graph_host = None
        self.load_inputs(state, forcing)
        if not self.use_graph:
            self._enqueue_all_to_host()
            return self.host_ring
        if self.graph_host is None:
            self._ensure_graph()                      # warms tables / attributes
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._enqueue_all_to_host()
            self.graph_host = g
        self.graph_host.replay()
        return self.host_ring
