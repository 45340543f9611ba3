#!/usr/bin/env python
"""
bench.py -- forecast steps/sec of the C48 six-face Weyn-2020 U-Net rollout (BASELINE.json configs[1]).

One bench "step" = one full autoregressive rollout (--rollout-steps 6-hourly model steps, default 100) of an ensemble of
--batch initial conditions, i.e. --rollout-steps x 11 launches of the halo-fused cubed-sphere convolution kernel.
`value` counts sample-steps: rollout_steps x batch x K / time, summed over ranks (each rank rolls out its own ensemble
shard: weak scaling, no collective on the data path -- SURVEY.md section 8e "replicas only").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--dtype bf16|fp32]

Timing: CUDA events on the launching stream around every bench step, L2 flushed (512 MiB write) between steps outside
the timed region, barrier + synchronize before and after, max over ranks.
`--impl reference`: the reference's TensorFlow path cannot run here; its CPU restatement (oracle/cs_oracle.py,
torch/oneDNN on all host cores, with the per-step host hop of DLWP/model/models.py:446-454) is timed on a bounded sample.

Besides the headline line's keys the JSON carries: `train_*` (BASELINE configs[2]/[3]: data-parallel training step),
`dp_equivalence` (N > 1: one DP step on N shards == one 1-GPU step on the concatenated batch), and `configs` with the other
BASELINE configurations -- configs[0] (single 3->3 convolution, batch 1), B = 1 rollout latency, configs[4] (C96, 12
variables, 1000 steps).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _oracle():
    """The CPU restatement under oracle/ -- imported by the CPU legs only (cpu_baseline, --impl reference); the GPU
    arm never touches it."""
    p = os.path.join(ROOT, 'oracle')
    if p not in sys.path:
        sys.path.insert(0, p)
    import cs_oracle
    return cs_oracle

import numpy as np  # noqa: E402
import torch  # noqa: E402

C_FORC = 4        # 2 insolation + 2 constants
BASE = 32
METRIC = 'forecast steps/sec C48 6-face U-Net rollout'
UNIT = 'sample-steps/s'


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p['hbm_gbs'], tf_burst=p['bf16_tflops'], tf_sust=p['bf16_tflops_sustained'], src='measured')
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src='fallback')


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.idx), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (('hw_slowdown', 5), ('hw_thermal_slowdown', 6), ('sw_thermal_slowdown', 7),
                              ('sw_power_cap', 8)):
                if len(r) > col and r[col].lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def layer_work(k, cin, cout, edge, batch, in_bytes, out_bytes):
    flop = 2.0 * batch * 6 * edge * edge * k * k * cin * cout
    byts = batch * 6 * edge * edge * (cin * in_bytes + cout * out_bytes)   # conv in + out, halo not counted
    return flop, byts


def synth_inputs(batch, n_face, c_prog, seed=0):
    g = torch.Generator().manual_seed(seed)
    state = torch.randn(batch, 6, n_face, n_face, c_prog, generator=g).clamp_(-5, 5)
    forcing = torch.rand(batch, 6, n_face, n_face, C_FORC, generator=g)
    return state, forcing


def workload_name(n_face, c_prog, rollout_steps, batch):
    return ('unet2 (Weyn-2020) C%d rollout, %d vars x 2 tsteps in/out (+2 solar +2 const), ' % (n_face, c_prog // 2) +
            '%d 6-hr steps, ensemble of %d members per GPU, random-init weights' % (rollout_steps, batch))


def cpu_rollout_rate(batch, steps, threads, n_face, c_prog, repeats=1):
    """The reference path's CPU restatement: oracle rollout, oneDNN conv, numpy hop per step.  -> sample-steps/s"""
    O = _oracle()
    torch.set_num_threads(threads)
    params = O.make_unet2_params(c_prog + C_FORC, c_prog, base=BASE, seed=1)
    state, forcing = synth_inputs(batch, n_face, c_prog)
    with torch.no_grad():
        O.rollout_fast(params, state, forcing, 1)        # warm-up (oneDNN primitive cache)
        best = None
        for _ in range(repeats):
            t0 = time.perf_counter()
            O.rollout_fast(params, state, forcing, steps)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    return batch * steps / best, best


def run_reference(args, rank, world, n_face, c_prog):
    """CPU arm: same members per bench step as the GPU arm, a bounded number of 6-hour steps of the same rollout (the CPU
    cost per sample-step does not depend on how far the rollout has come)."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    b, s = args.ref_batch, args.ref_steps
    O = _oracle()
    torch.set_num_threads(threads)
    params = O.make_unet2_params(c_prog + C_FORC, c_prog, base=BASE, seed=1)
    state, forcing = synth_inputs(b, n_face, c_prog)
    times = []
    with torch.no_grad():
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            O.rollout_fast(params, state, forcing, s)
            if i >= args.warmup:
                times.append(time.perf_counter() - t0)
    total = sum(times)
    value = b * s * len(times) / total
    sample = ('oracle/cs_oracle.py rollout_fast (torch-CPU oneDNN restatement of DLWP/custom.py as LUT gather + batched '
              'faces, models.py:446-454 host hop), fp32, %d members x %d of the %d steps per bench step'
              % (b, s, args.rollout_steps))
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total / len(times),
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': workload_name(n_face, c_prog, args.rollout_steps, args.batch),
                       'rollout_steps': args.rollout_steps, 'batch_per_gpu': args.batch, 'face_edge': n_face},
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': sample},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


def time_layers(eng, flush, reps=10):
    """Per-layer device time inside one model step: the step's launches are enqueued in order (eagerly, not from the
    graph) with a CUDA event between consecutive launches, L2 flushed (512 MiB write) before every repetition of the step
    -- each layer therefore sees the cache state it has in production (whatever part of its predecessor's output the
    126 MB L2 still holds), not a layer-private warm L2.  Median over the repetitions.  -> list of (name, ms)"""
    from dlwp_cs_b200 import _lib
    launches = []
    fused = getattr(eng, 'fused_head', None)
    for i, (name, d, s0, s1, dst, packed) in enumerate(eng.plan):
        if fused is not None and i == fused + 1:
            continue
        o = eng.ring[0] if dst == 'out' else eng.buf[dst]
        if fused is not None and i == fused:          # the 1x1 output layer runs inside this launch
            head = eng.plan[i + 1]
            launches.append((name + '+' + head[0], d, eng._src(s0, 0), eng._src(s1, 0), packed, eng.ring[0], head))
        elif i in getattr(eng, 'pool_out', {}):       # this launch also writes the 2x2 mean of its output
            launches.append((name, d, eng._src(s0, 0), eng._src(s1, 0), packed, o, ('pool', eng.buf[eng.pool_out[i]])))
        else:
            launches.append((name, d, eng._src(s0, 0), eng._src(s1, 0), packed, o, None))

    def go(d, a, b, packed, o, head):
        if head is None:
            _lib.conv2d_fwd(d, a, b, packed, out=o)
        elif head[0] == 'pool':
            _lib.conv2d_fwd_pool(d, a, b, packed, out=o, out_pool=head[1])
        else:
            _lib.conv2d_fwd_head(d, a, b, packed, head[1], head[5], out=o)
    for _ in range(3):
        for name, d, a, b, packed, o, head in launches:
            go(d, a, b, packed, o, head)
    times = {name: [] for name, *_ in launches}
    for _ in range(reps):
        flush.zero_()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(launches) + 1)]
        evs[0].record()
        for i, (name, d, a, b, packed, o, head) in enumerate(launches):
            go(d, a, b, packed, o, head)
            evs[i + 1].record()
        torch.cuda.synchronize()
        for i, (name, *_r) in enumerate(launches):
            times[name].append(evs[i].elapsed_time(evs[i + 1]))
    return [(name, sorted(times[name])[reps // 2]) for name, *_ in launches]


def timed_rollout(eng, flush, steps, warmup):
    """ms per bench step (one whole rollout) of an engine whose inputs are resident, L2 flushed between steps."""
    for _ in range(warmup):
        eng.launch()
    torch.cuda.synchronize()
    ev = []
    for _ in range(steps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        eng.launch()
        e.record()
        ev.append((s, e))
    torch.cuda.synchronize()
    return sum(s.elapsed_time(e) for s, e in ev) / steps


def conv_shape_records(dev, flush, pk):
    """Single CubeSphereConv2D launches (bf16, 3x3, fused halo, bias + ReLU) at batch 64 on shapes that isolate the
    convolution kernel from the network: the 32-wide C48 layer, and 64 -> 64 at C48 (arithmetic intensity 288 > ridge: the
    tensor-bound stress shape).  L2 flushed before every repetition; four back-to-back launches per repetition."""
    from dlwp_cs_b200 import _lib
    out = {}
    for n, cin, cout in ((48, 32, 32), (48, 64, 64)):
        b = 64
        g = torch.Generator().manual_seed(n + cin)
        x = torch.randn(b, 6, n, n, cin, generator=g).bfloat16().to(dev)
        w = [(torch.randn(3, 3, cin, cout, generator=g) * 0.05).to(dev) for _ in range(2)]
        bs = [torch.zeros(cout, device=dev) for _ in range(2)]
        d = _lib.make_desc(b, n, cin, cout, (3, 3), (1, 1), (1, 1), 1, False, True, False, True, _lib.ACT_CAPPED_LEAKY_RELU,
                           0.1, 10.0, _lib.BF16, _lib.BF16)
        packed = _lib.pack_weights(d, w[0], w[1], None, bs[0], bs[1], None)
        y = _lib.conv2d_fwd(d, x, None, packed)
        torch.cuda.synchronize()
        ts = []
        for _ in range(7):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(4):
                _lib.conv2d_fwd(d, x, None, packed, out=y)
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e) / 4)
        ms = sorted(ts)[len(ts) // 2]
        flop, byts = layer_work(3, cin, cout, n, b, 2, 2)
        t_c, t_m = flop / (pk['tf_sust'] * 1e12), byts / (pk['hbm'] * 1e9)
        out['%dto%d_c%d' % (cin, cout, n)] = {
            'us': round(1e3 * ms, 1), 'tflops': round(flop / ms / 1e9, 1), 'gbs': round(byts / ms / 1e6, 1),
            'bound': 'tensor' if t_c > t_m else 'hbm', 'frac_of_roof': round(max(t_c, t_m) * 1e3 / ms, 3),
            'frac_of_bf16_burst_peak': round(flop / ms / 1e9 / pk['tf_burst'], 3)}
    return out


def conv_3to3_record(dev, threads):
    """BASELINE configs[0]: single CubeSphereConv2D forward, C48, 3 -> 3 channels, batch 1, float32."""
    from dlwp_cs_b200 import _lib
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 6, 48, 48, 3, generator=g)
    lim = (6.0 / 54.0) ** 0.5
    w = [(torch.rand(3, 3, 3, 3, generator=g) * 2 - 1) * lim for _ in range(2)]
    b = [(torch.rand(3, generator=g) * 2 - 1) * 0.1 for _ in range(2)]
    d = _lib.make_desc(1, 48, 3, 3, (3, 3), halo=1)
    xd = x.to(dev)
    packed = _lib.pack_weights(d, w[0].to(dev), w[1].to(dev), None, b[0].to(dev), b[1].to(dev))
    for _ in range(5):
        y = _lib.conv2d_fwd(d, xd, None, packed)
    reps = 200
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        _lib.conv2d_fwd(d, xd, None, packed, out=y)
    e1.record()
    e1.synchronize()
    gpu_us = 1e3 * e0.elapsed_time(e1) / reps
    xn, wn, bn = x.numpy(), [t.numpy() for t in w], [t.numpy() for t in b]
    for _ in range(3):
        yh = _lib.conv2d_fwd_host(d, xn, wn[0], wn[1], None, bn[0], bn[1])
    t0 = time.perf_counter()
    for _ in range(50):
        yh = _lib.conv2d_fwd_host(d, xn, wn[0], wn[1], None, bn[0], bn[1])
    host_us = 1e6 * (time.perf_counter() - t0) / 50
    O = _oracle()
    torch.set_num_threads(threads)
    with torch.no_grad():
        f = lambda: O.cube_sphere_conv2d_fast(O.cube_sphere_pad_fast(x, 1), w[0], w[1], b[0], b[1])
        for _ in range(5):
            yc = f()
        t0 = time.perf_counter()
        for _ in range(100):
            yc = f()
        cpu_us = 1e6 * (time.perf_counter() - t0) / 100
    err = float((torch.from_numpy(yh) - yc).abs().max() / yc.abs().max())
    return {'config': 'BASELINE configs[0]: single CubeSphereConv2D fwd (pad 1 fused), C48, 3->3 ch, batch 1, float32',
            'gpu_us_resident': round(gpu_us, 2), 'gpu_us_host_buffers': round(host_us, 1), 'cpu_us': round(cpu_us, 1),
            'cpu_cores': threads, 'cpu_kind': 'port (oracle LUT gather + oneDNN conv)', 'max_rel_diff_vs_cpu': err,
            'note': '2.24 MFLOP / 0.33 MB: launch-latency bound on the GPU; host-buffer call = H2D + pack + kernel + D2H'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=64, help='ensemble members (initial conditions) per GPU')
    ap.add_argument('--rollout-steps', type=int, default=100)
    ap.add_argument('--face-edge', type=int, default=48, help='48 = C48 (headline config), 96 = C96 (BASELINE configs[4])')
    ap.add_argument('--variables', type=int, default=7, help='prognostic variables per time step (x2 time steps)')
    ap.add_argument('--dtype', default='auto', choices=['auto', 'bf16', 'fp32'])
    ap.add_argument('--ref-batch', type=int, default=None, help='CPU arm: members per bench step (default: --batch)')
    ap.add_argument('--ref-steps', type=int, default=25, help='CPU arm: 6-hour steps per bench step (bounded sample)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--host-chunk', type=int, default=int(os.environ.get('DLWPCS_HOST_CHUNK', '2')),
                    help='end-to-end arm: model steps per device->host transfer (RolloutEngine.run_to_host chunk)')
    ap.add_argument('--no-train', action='store_true')
    ap.add_argument('--no-extra', action='store_true', help='skip the configs[0] / B=1 latency / C96 sub-records')
    ap.add_argument('--train-batch', type=int, default=32, help='training samples per GPU per step')
    ap.add_argument('--train-steps', type=int, default=10)
    ap.add_argument('--train-dtype', default='bf16', choices=['bf16', 'f32'])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    if args.ref_batch is None:
        args.ref_batch = args.batch

    n_face, c_prog = args.face_edge, 2 * args.variables
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))

    if args.impl == 'reference':
        run_reference(args, rank, world, n_face, c_prog)
        return

    from dlwp_cs_b200 import _lib
    from dlwp_cs_b200.unet import CubeSphereUNet2, RolloutEngine, unet2_layer_specs
    _lib.load()
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    dist = None
    if world > 1:
        if os.environ.get('NCCL_DEBUG') and not os.environ.get('NCCL_DEBUG_FILE'):
            os.environ['NCCL_DEBUG_FILE'] = '/dev/stderr'     # NCCL logs (version banner, INFO) go to stdout otherwise: keep
                                                              # them, but away from the JSON line
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    torch.manual_seed(1)                      # random-init weights of the architecture (glorot-uniform kernels, zero biases)
    model = CubeSphereUNet2(c_prog + C_FORC, c_prog, base=BASE).to(dev)
    dtype = args.dtype
    eng = None
    if dtype in ('auto', 'bf16'):
        try:
            eng = RolloutEngine(model, args.batch, n_face, args.rollout_steps, forcing_channels=C_FORC,
                                dtype=torch.bfloat16, use_graph=not args.no_graph)
            dtype = 'bf16'
        except _lib.DlwpcsError as e:
            if args.dtype == 'bf16':
                raise
            sys.stderr.write('bf16 tensor-core path unavailable (%s); benchmarking the fp32 path\n' % e)
    if eng is None:
        eng = RolloutEngine(model, args.batch, n_face, args.rollout_steps, forcing_channels=C_FORC,
                            dtype=torch.float32, use_graph=not args.no_graph)
        dtype = 'fp32'
    tdt = torch.bfloat16 if dtype == 'bf16' else torch.float32
    esz = 2 if dtype == 'bf16' else 4

    state, forcing = synth_inputs(args.batch, n_face, c_prog, seed=rank)
    h_state, h_forcing = state.to(tdt).pin_memory(), forcing.to(tdt).pin_memory()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    # ---------------- device-resident arm: inputs already in HBM
    eng.load_inputs(h_state, h_forcing)
    torch.cuda.synchronize()
    for _ in range(args.warmup):
        eng.launch()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = []
    for _ in range(args.steps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        eng.launch()
        e.record()
        ev.append((s, e))
    barrier()
    dev_ms = sum(s.elapsed_time(e) for s, e in ev)

    # ---------------- end-to-end arm: pinned host inputs in, whole forecast out to pinned host memory, every step
    # (RolloutEngine.run_to_host: the device->host transfer of finished model steps overlaps the remaining steps)
    for _ in range(2):
        h_ring = eng.run_to_host(h_state, h_forcing, chunk=args.host_chunk)
    barrier()
    ev2 = []
    for _ in range(args.steps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        h_ring = eng.run_to_host(h_state, h_forcing, chunk=args.host_chunk)
        e.record()
        ev2.append((s, e))
    barrier()
    clocks = sampler.stop()
    e2e_ms = sum(s.elapsed_time(e) for s, e in ev2)

    # ---------------- the ceiling of the end-to-end arm: the same forecast bytes as a plain device -> pinned-host copy, all
    # ranks at once (what the host's PCIe / memory system sinks when nothing else is going on)
    d2h_bytes = h_ring.numel() * esz
    src = eng.ring.reshape(-1)[:h_ring.numel()]               # a contiguous run of exactly the forecast's byte count
    h_flat = h_ring.reshape(-1)
    for _ in range(2):
        h_flat.copy_(src, non_blocking=True)
    barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(3):
        h_flat.copy_(src, non_blocking=True)
    e.record()
    barrier()
    copy_ms = s.elapsed_time(e) / 3
    del src

    dev_ms, e2e_ms, copy_ms = max_over_ranks([dev_ms, e2e_ms, copy_ms])
    units = float(args.rollout_steps) * args.batch * args.steps * world
    value = units / (dev_ms * 1e-3)
    e2e_value = units / (e2e_ms * 1e-3)

    # ---------------- roofline of the dominant kernel (rank 0)
    roof = None
    layers = []
    if rank == 0:
        pk = peaks()
        lt = time_layers(eng, flush)
        specs = {s[0]: s for s in unet2_layer_specs(c_prog + C_FORC, c_prog, BASE)}
        edges = {name: d.n for name, d, *_ in eng.plan}
        tot = sum(ms for _, ms in lt)
        for name, ms in lt:
            # a fused launch ('a+b') carries the algorithmic work of both layers (SURVEY 8(d) counts each layer's input
            # and output; the fused launch moves fewer bytes than that)
            flop = byts = 0.0
            for part in name.split('+'):
                _, k, ci, co = specs[part]
                f_, b_ = layer_work(k, ci, co, edges[part], args.batch, esz, esz)
                flop, byts = flop + f_, byts + b_
            tf, gb = flop / ms / 1e9, byts / ms / 1e6
            t_c = flop / (pk['tf_sust'] * 1e12) if dtype == 'bf16' else 0.0
            t_m = byts / (pk['hbm'] * 1e9)
            layers.append({'layer': name, 'ms': round(ms, 4), 'share': round(ms / tot, 3), 'tflops': round(tf, 1),
                           'gbs': round(gb, 1), 'bound': 'tensor' if t_c > t_m else 'hbm',
                           'frac_of_roof': round(max(t_c, t_m) * 1e3 / ms, 3)})
        top = max(layers, key=lambda r: r['ms'])
        if top['bound'] == 'tensor':
            roof = {'bound': 'tensor', 'achieved': top['tflops'], 'peak': pk['tf_sust'], 'unit': 'TFLOP/s',
                    'frac': round(top['tflops'] / pk['tf_sust'], 4)}
        else:
            roof = {'bound': 'hbm', 'achieved': top['gbs'], 'peak': pk['hbm'], 'unit': 'GB/s',
                    'frac': round(top['gbs'] / pk['hbm'], 4)}
        traffic = None
        for tname in ('r2_traffic.json', 'r1_traffic.json'):     # DRAM bytes of that launch from the ncu --set full capture
            tpath = os.path.join(ROOT, 'profiles', tname)
            if dtype == 'bf16' and os.path.exists(tpath):
                tj = json.load(open(tpath))
                if top['layer'] in tj['dram_bytes_per_launch']:
                    traffic = tj['dram_bytes_per_launch'][top['layer']] * args.batch / tj['batch']
                    break
        step_ms = dev_ms / args.steps / args.rollout_steps          # one 6-hour model step inside the graph-replayed rollout
        t_roof = sum(max(layer_work(s[1], s[2], s[3], edges[s[0]], args.batch, esz, esz)[0] / (pk['tf_sust'] * 1e12)
                         if dtype == 'bf16' else 0.0,
                         layer_work(s[1], s[2], s[3], edges[s[0]], args.batch, esz, esz)[1] / (pk['hbm'] * 1e9))
                     for s in specs.values())
        roof.update({'traffic': traffic, 'kernel': 'cs_conv (%s: %s)' % (dtype, top['layer']), 'peak_source': pk['src'],
                     'share_of_step': top['share'],
                     'whole_step_frac_of_roof': round(t_roof * 1e3 / step_ms, 3),
                     'whole_step_us': round(1e3 * step_ms, 1),
                     'note': 'per-layer times: CUDA events between the launches of one model step enqueued in order, L2 '
                             'flushed before every repetition of the step; whole_step: sum of the per-layer roofline '
                             'times / the measured 6-hour step inside the rollout graph'})

    # ---------------- the other BASELINE configurations (rank 0, N = 1 only: they are single-GPU records)
    extra = None
    threads = os.cpu_count() or 1
    if rank == 0 and world == 1 and not args.no_extra and dtype == 'bf16':
        extra = {}
        extra['conv_3to3_b1'] = conv_3to3_record(dev, threads)
        extra['conv_b64_shapes'] = conv_shape_records(dev, flush, peaks())
        # B = 1 latency of the headline rollout (report both B = 1 latency and batched-ensemble throughput)
        e1 = RolloutEngine(model, 1, n_face, args.rollout_steps, forcing_channels=C_FORC, dtype=tdt)
        e1.load_inputs(h_state[:1], h_forcing[:1])
        ms1 = timed_rollout(e1, flush, 5, 3)
        extra['rollout_b1_latency'] = {'config': 'C%d unet2 rollout, 1 member, %d steps' % (n_face, args.rollout_steps),
                                       'ms_per_rollout': round(ms1, 3),
                                       'us_per_6h_step': round(1e3 * ms1 / args.rollout_steps, 1),
                                       'steps_per_s': round(args.rollout_steps / (ms1 * 1e-3), 1)}
        del e1
        # float32 rollout (the reference's arithmetic): float32-accurate tensor-core path (split bf16 x3, float32 accumulation
        # and storage) next to the CUDA-core float32 parity kernels; same ensemble, same number of steps
        rec = {'config': 'C%d unet2 rollout in float32, %d members, %d steps' % (n_face, args.batch, args.rollout_steps)}
        for key, tcflag, reps in (('tensor_core_split_bf16x3', True, 3), ('cuda_core_fp32', False, 1)):
            ef = RolloutEngine(model, args.batch, n_face, args.rollout_steps, forcing_channels=C_FORC, dtype=torch.float32,
                               tensor_cores=tcflag)
            ef.load_inputs(h_state.float(), h_forcing.float())
            msf = timed_rollout(ef, flush, reps, 1 if not tcflag else 2)
            rec[key] = {'value': round(args.rollout_steps * args.batch / (msf * 1e-3), 1), 'unit': UNIT,
                        'ms_per_rollout': round(msf, 2), 'launches_per_6h_step': ef.launches_per_step}
            del ef
            torch.cuda.empty_cache()
        rec['note'] = ('tensor_core: dlwpcs_split3 + bf16 tcgen05 kernel over [hi|lo|hi] x [w_hi;w_hi;w_lo], ~2^-17 per product '
                       '(tests/test_gpu_tc.py::test_tc32_*); cuda_core: conv_fp32_kernel, the strict 1e-5 parity path')
        extra['rollout_fp32'] = rec
        if n_face == 48:
            # BASELINE configs[4]: C96, 12 variables x 2 time steps, 1000-step rollout, 16 members
            cp96, b96, st96 = 24, 16, 1000
            torch.manual_seed(1)
            m96 = CubeSphereUNet2(cp96 + C_FORC, cp96, base=BASE).to(dev)
            e96 = RolloutEngine(m96, b96, 96, st96, forcing_channels=C_FORC, dtype=tdt)
            s96, f96 = synth_inputs(b96, 96, cp96)
            e96.load_inputs(s96.to(tdt).pin_memory(), f96.to(tdt).pin_memory())
            ms96 = timed_rollout(e96, flush, 3, 3)
            sp96 = unet2_layer_specs(cp96 + C_FORC, cp96, BASE)
            div = {'conv_2d_2': 2, 'conv_2d_2_2': 2, 'conv_2d_5_2': 4, 'conv_2d_5': 4, 'conv_2d_6_2': 2, 'conv_2d_6': 2}
            pk = peaks()
            t_roof96 = sum(max(layer_work(k, ci, co, 96 // div.get(nm, 1), b96, esz, esz)[0] / (pk['tf_sust'] * 1e12),
                               layer_work(k, ci, co, 96 // div.get(nm, 1), b96, esz, esz)[1] / (pk['hbm'] * 1e9))
                           for nm, k, ci, co in sp96)
            rec = {'config': 'BASELINE configs[4]: C96 unet2, 12 vars x 2 tsteps (+2 +2), %d 6-hr steps, %d members' % (st96, b96),
                   'value': round(st96 * b96 / (ms96 * 1e-3), 1), 'unit': UNIT, 'ms_per_rollout': round(ms96, 2),
                   'us_per_6h_step': round(1e3 * ms96 / st96, 1),
                   'roofline': {'whole_step_frac_of_roof': round(t_roof96 * 1e3 / (ms96 / st96), 3),
                                'peak_source': pk['src']}}
            del e96, m96
            torch.cuda.empty_cache()
            if not args.no_cpu_baseline:
                v, secs = cpu_rollout_rate(4, 50, threads, 96, cp96)
                rec['cpu_baseline'] = {'value': round(v, 2), 'unit': UNIT, 'cores': threads, 'kind': 'port',
                                       'sample': '4 members x 50 of the 1000 steps = %.1f s' % secs}
            extra['c96_12var_1000steps'] = rec

    # ---------------- training step (BASELINE.json configs[2]/[3]): batch-sharded DP, one flat NCCL all-reduce per step
    train = None
    dp_check = None
    rollout_launches = args.steps * args.rollout_steps * eng.launches_per_step * 2
    if not args.no_train:
        del eng, flush
        torch.cuda.empty_cache()
        from dlwp_cs_b200.train import DataParallelTrainer
        torch.manual_seed(1)
        tmodel = CubeSphereUNet2(c_prog + C_FORC, c_prog, base=BASE).to(dev)
        trainer = DataParallelTrainer(tmodel, lr=1e-3)
        g = torch.Generator().manual_seed(100 + rank)
        tb = args.train_batch
        ttd = torch.bfloat16 if args.train_dtype == 'bf16' else torch.float32
        xs = torch.randn(tb, 6, n_face, n_face, c_prog + C_FORC, generator=g).to(dev).to(ttd)
        ts = torch.randn(tb, 6, n_face, n_face, c_prog, generator=g).to(dev).to(ttd)
        for _ in range(3):
            trainer.step(xs, ts)
        barrier()
        s_ev, e_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_ev.record()
        for _ in range(args.train_steps):
            loss = trainer.step(xs, ts)
        e_ev.record()
        barrier()
        tms = max_over_ranks([s_ev.elapsed_time(e_ev)])[0] / args.train_steps
        div = {'conv_2d_2': 2, 'conv_2d_2_2': 2, 'conv_2d_5_2': 4, 'conv_2d_5': 4, 'conv_2d_6_2': 2, 'conv_2d_6': 2}
        train = {'metric': 'train samples/sec, unet2 C%d fwd+bwd+Adam' % n_face, 'value': tb * world / (tms * 1e-3),
                 'unit': 'samples/s', 'ms_per_step': tms, 'global_batch': tb * world, 'batch_per_gpu': tb,
                 'dtype': args.train_dtype, 'parallelism': 'dp%d, one flat all-reduce of %d float32 gradients per step'
                                                % (world, trainer.flat.count),
                 'loss': float(loss.item()),
                 # forward + dgrad + wgrad = 3 x the forward FLOP of every layer
                 'algorithmic_gflop_per_sample': round(3e-9 * sum(
                     layer_work(s[1], s[2], s[3], n_face // div.get(s[0], 1), 1, 2, 2)[0]
                     for s in unet2_layer_specs(c_prog + C_FORC, c_prog, BASE)), 3),
                 'note': ('bf16 activations: forward, dgrad and wgrad on tcgen05 kernels (fp32 accumulation in tensor '
                          'memory); float32 master weights, fused Adam')
                 if args.train_dtype == 'bf16' else 'float32 CUDA-core kernels (1e-5 parity path)'}
        train['tflops'] = round(train['value'] * train['algorithmic_gflop_per_sample'] * 1e-3, 1)
        trainer.close()
        del trainer, tmodel

        # ---- data-parallel equivalence (train_tf.py:166): ONE optimizer step on `world` shards ==
        # one single-GPU step on the concatenated batch; parameters after Adam bitwise identical on every rank
        if dist is not None:
            cb = 4                                                     # samples per rank for the check
            gg = torch.Generator().manual_seed(777)                    # the same global batch on every rank
            gx = torch.randn(cb * world, 6, n_face, n_face, c_prog + C_FORC, generator=gg).to(dev).to(ttd)
            gt = torch.randn(cb * world, 6, n_face, n_face, c_prog, generator=gg).to(dev).to(ttd)
            torch.manual_seed(5)
            m_dp = CubeSphereUNet2(c_prog + C_FORC, c_prog, base=BASE).to(dev)
            torch.manual_seed(5)
            m_one = CubeSphereUNet2(c_prog + C_FORC, c_prog, base=BASE).to(dev)
            t_dp = DataParallelTrainer(m_dp, lr=1e-3, use_graph=False)
            t_one = DataParallelTrainer(m_one, lr=1e-3, use_graph=False, distributed=False)
            l_dp = t_dp.step(gx[rank * cb:(rank + 1) * cb], gt[rank * cb:(rank + 1) * cb]).clone()
            l_one = t_one.step(gx, gt).clone()
            torch.cuda.synchronize()
            g_dp, g_one = t_dp.flat.grad / world, t_one.flat.grad      # summed shard-mean gradients / world vs global mean
            gscale = float(g_one.abs().max())
            gdiff = float((g_dp - g_one).abs().max()) / max(gscale, 1e-30)
            lsum = l_dp.clone()
            dist.all_reduce(lsum)
            ldiff = abs(float(lsum) / world - float(l_one)) / max(abs(float(l_one)), 1e-30)
            plist = [torch.empty_like(t_dp.flat.param) for _ in range(world)]
            dist.all_gather(plist, t_dp.flat.param)
            bitwise = all(torch.equal(plist[0], p) for p in plist[1:])
            pdiff = float((t_dp.flat.param - t_one.flat.param).abs().max())
            vals = max_over_ranks([gdiff, ldiff, 0.0 if bitwise else 1.0, pdiff])
            tol = 2e-3 if args.train_dtype == 'bf16' else 1e-4
            dp_check = {'grad_max_rel_diff_vs_1gpu': vals[0], 'loss_rel_diff_vs_1gpu': vals[1],
                        'params_bitwise_across_ranks': vals[2] == 0.0, 'param_max_abs_diff_vs_1gpu_after_adam': vals[3],
                        'tolerance': tol, 'samples_per_rank': cb, 'dtype': args.train_dtype,
                        'ok': bool(vals[0] <= tol and vals[1] <= tol and vals[2] == 0.0 and vals[3] <= 2.5e-3)}
            del t_dp, t_one, m_dp, m_one

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, secs = cpu_rollout_rate(args.ref_batch, args.ref_steps, threads, n_face, c_prog)
        cpu = {'value': v, 'unit': UNIT, 'cores': threads, 'kind': 'port',
               'sample': 'oracle rollout_fast (torch-CPU oneDNN restatement, LUT-gather halo, fp32, host hop per step), '
                         '%d members x %d of the %d steps = %.1f s' % (args.ref_batch, args.ref_steps, args.rollout_steps, secs)}

    if rank == 0:
        h2d = (h_state.numel() + h_forcing.numel()) * esz
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': dev_ms / args.steps, 'higher_is_better': True,
                'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16' if dtype == 'bf16' else 'f32',
                'data': 'synthetic',
                'config': {'workload': workload_name(n_face, c_prog, args.rollout_steps, args.batch),
                           'rollout_steps': args.rollout_steps, 'batch_per_gpu': args.batch, 'face_edge': n_face,
                           'l2': '512 MiB flush write between bench steps; per-step activation set %.0f MiB > L2'
                                 % (args.batch * 6 * n_face * n_face * 32 * esz * 4 / 2 ** 20),
                           'cuda_graph': not args.no_graph, 'parallelism': 'replicas x%d' % world},
                'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h_bytes,
                        'ms_per_step': e2e_ms / args.steps,
                        'd2h_ceiling_gbs_per_gpu': round(d2h_bytes / copy_ms / 1e6, 2),
                        'frac_of_d2h_ceiling': round((copy_ms * args.steps) / e2e_ms, 3),
                        'note': 'ceiling = the same forecast bytes as a plain device->pinned-host copy issued by all '
                                'ranks at once (max over ranks): what the host PCIe / memory system sinks'},
                'gpu_launches': rollout_launches,
                'clocks': clocks, 'roofline': roof, 'cpu_baseline': cpu,
                'train_value': None if train is None else train['value'],
                'train_unit': 'samples/s',
                'train_ms_per_step': None if train is None else train['ms_per_step'],
                'train_global_batch': None if train is None else train['global_batch'],
                'train_tflops': None if train is None else train['tflops'],
                'dp_equivalence': None if dp_check is None else dp_check['ok'],
                'dp_check': dp_check, 'train': train, 'configs': extra, 'layers': layers}
        print(json.dumps(line), flush=True)
    if dist is not None:
        # the captured graphs that held NCCL kernels were released by trainer.close(); tear the communicator down, guarded by
        # a watchdog so that a stuck teardown cannot keep the bench line from being delivered
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        done = threading.Event()

        def _teardown():
            try:
                dist.destroy_process_group()
            finally:
                done.set()
        th = threading.Thread(target=_teardown, daemon=True)
        th.start()
        if not done.wait(30.0):
            sys.stderr.write('destroy_process_group did not return within 30 s; leaving\n')
            sys.stderr.flush()
            os._exit(0)


if __name__ == '__main__':
    main()
