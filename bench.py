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
    """The CPU restatement under oracle/ -- imported by the two CPU legs only (cpu_baseline, --impl reference); the GPU
    arm never touches it."""
    p = os.path.join(ROOT, 'oracle')
    if p not in sys.path:
        sys.path.insert(0, p)
    import cs_oracle
    return cs_oracle

import numpy as np  # noqa: E402
import torch  # noqa: E402

N_FACE = 48
C_PROG = 14       # 7 variables x 2 time steps
C_FORC = 4        # 2 insolation + 2 constants
BASE = 32
# face edge of every unet2 layer relative to the input (two poolings): Azure/train_cs.py:277-305
EDGE_DIV = {'conv_2d_2': 2, 'conv_2d_2_2': 2, 'conv_2d_5_2': 4, 'conv_2d_5': 4, 'conv_2d_6_2': 2, 'conv_2d_6': 2}
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


def layer_work(name, k, cin, cout, edge, batch, in_bytes, out_bytes):
    flop = 2.0 * batch * 6 * edge * edge * k * k * cin * cout
    byts = batch * 6 * edge * edge * (cin * in_bytes + cout * out_bytes)   # conv in + out, halo not counted
    return flop, byts


def synth_inputs(batch, seed=0):
    g = torch.Generator().manual_seed(seed)
    state = torch.randn(batch, 6, N_FACE, N_FACE, C_PROG, generator=g).clamp_(-5, 5)
    forcing = torch.rand(batch, 6, N_FACE, N_FACE, C_FORC, generator=g)
    return state, forcing


def cpu_rollout_rate(batch, steps, threads, repeats=1):
    """The reference path's CPU restatement: oracle rollout, oneDNN conv, numpy hop per step.  -> sample-steps/s"""
    O = _oracle()
    torch.set_num_threads(threads)
    params = O.make_unet2_params(C_PROG + C_FORC, C_PROG, base=BASE, seed=1)
    state, forcing = synth_inputs(batch)
    with torch.no_grad():
        O.rollout_fast(params, state, forcing, 1)        # warm-up (oneDNN primitive cache)
        best = None
        for _ in range(repeats):
            t0 = time.perf_counter()
            O.rollout_fast(params, state, forcing, steps)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    return batch * steps / best, best


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    b, s = args.ref_batch, args.ref_steps
    O = _oracle()
    torch.set_num_threads(threads)
    params = O.make_unet2_params(C_PROG + C_FORC, C_PROG, base=BASE, seed=1)
    state, forcing = synth_inputs(b)
    times = []
    with torch.no_grad():
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            O.rollout_fast(params, state, forcing, s)
            if i >= args.warmup:
                times.append(time.perf_counter() - t0)
    total = sum(times)
    value = b * s * len(times) / total
    sample = ('oracle/cs_oracle.py rollout_fast (torch-CPU oneDNN restatement of DLWP/custom.py as LUT gather + batched faces, models.py:446-454 host hop), '
              'fp32, %d members x %d steps per bench step' % (b, s))
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total / len(times),
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'unet2 C48 rollout, 7 vars x 2 tsteps (+2 solar +2 const in), bounded sample of '
                                   'the %d-step x %d-member job' % (args.rollout_steps, args.batch),
                       'rollout_steps': s, 'batch': b},
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': sample},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


def time_layers(eng, reps=20):
    """Per-layer device time of one model step (CUDA events on the current stream, L2-cold-ish: the ensemble's
    activations are far larger than L2).  -> list of (name, ms)"""
    out = []
    for name, d, s0, s1, dst, packed in eng.plan:
        from dlwp_cs_b200 import _lib
        o = eng.ring[0] if dst == 'out' else eng.buf[dst]
        a, b = eng._src(s0, 0), eng._src(s1, 0)
        for _ in range(3):
            _lib.conv2d_fwd(d, a, b, packed, out=o)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            _lib.conv2d_fwd(d, a, b, packed, out=o)
        e1.record()
        e1.synchronize()
        out.append((name, e0.elapsed_time(e1) / reps))
    return out


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
    ap.add_argument('--ref-batch', type=int, default=16)
    ap.add_argument('--ref-steps', type=int, default=100)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--no-train', action='store_true')
    ap.add_argument('--train-batch', type=int, default=32, help='training samples per GPU per step')
    ap.add_argument('--train-steps', type=int, default=5)
    ap.add_argument('--train-dtype', default='bf16', choices=['bf16', 'f32'])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup

    global N_FACE, C_PROG
    N_FACE, C_PROG = args.face_edge, 2 * args.variables
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))

    if args.impl == 'reference':
        run_reference(args, rank, world)
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
        if os.environ.get('NCCL_DEBUG', '').upper() == 'VERSION':
            os.environ['NCCL_DEBUG'] = 'WARN'       # NCCL would print its version banner on stdout, next to the JSON line
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    torch.manual_seed(1)                      # random-init weights of the architecture (glorot-uniform kernels, zero biases)
    model = CubeSphereUNet2(C_PROG + C_FORC, C_PROG, base=BASE).to(dev)
    dtype = args.dtype
    eng = None
    if dtype in ('auto', 'bf16'):
        try:
            eng = RolloutEngine(model, args.batch, N_FACE, args.rollout_steps, forcing_channels=C_FORC,
                                dtype=torch.bfloat16, use_graph=not args.no_graph)
            dtype = 'bf16'
        except _lib.DlwpcsError as e:
            if args.dtype == 'bf16':
                raise
            sys.stderr.write('bf16 tensor-core path unavailable (%s); benchmarking the fp32 path\n' % e)
    if eng is None:
        eng = RolloutEngine(model, args.batch, N_FACE, args.rollout_steps, forcing_channels=C_FORC,
                            dtype=torch.float32, use_graph=not args.no_graph)
        dtype = 'fp32'
    tdt = torch.bfloat16 if dtype == 'bf16' else torch.float32
    esz = 2 if dtype == 'bf16' else 4

    state, forcing = synth_inputs(args.batch, seed=rank)
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
        h_ring = eng.run_to_host(h_state, h_forcing)
    barrier()
    ev2 = []
    for _ in range(args.steps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        h_ring = eng.run_to_host(h_state, h_forcing)
        e.record()
        ev2.append((s, e))
    barrier()
    clocks = sampler.stop()
    e2e_ms = sum(s.elapsed_time(e) for s, e in ev2)

    t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = t.tolist()
    units = float(args.rollout_steps) * args.batch * args.steps * world
    value = units / (dev_ms * 1e-3)
    e2e_value = units / (e2e_ms * 1e-3)

    # ---------------- roofline of the dominant kernel (rank 0)
    roof = None
    layers = []
    if rank == 0:
        pk = peaks()
        lt = time_layers(eng)
        specs = {s[0]: s for s in unet2_layer_specs(C_PROG + C_FORC, C_PROG, BASE)}
        edges = {name: d.n for name, d, *_ in eng.plan}
        tot = sum(ms for _, ms in lt)
        for name, ms in lt:
            _, k, ci, co = specs[name]
            flop, byts = layer_work(name, k, ci, co, edges[name], args.batch, esz, esz)
            tf, gb = flop / ms / 1e9, byts / ms / 1e6
            tf_peak = pk['tf_sust'] if dtype == 'bf16' else None
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
        tpath = os.path.join(ROOT, 'profiles', 'r1_traffic.json')
        if dtype == 'bf16' and os.path.exists(tpath):            # DRAM bytes of that launch from the ncu --set full capture
            tj = json.load(open(tpath))
            if top['layer'] in tj['dram_bytes_per_launch']:
                traffic = tj['dram_bytes_per_launch'][top['layer']] * args.batch / tj['batch']
        roof.update({'traffic': traffic, 'kernel': 'cs_conv (%s: %s)' % (dtype, top['layer']), 'peak_source': pk['src'],
                     'share_of_step': top['share'],
                     'whole_step_frac_of_roof': round(sum(r['frac_of_roof'] * r['ms'] for r in layers) / tot, 3)})

    # ---------------- training step (BASELINE.json configs[2]/[3]): batch-sharded DP, one flat NCCL all-reduce per step
    train = None
    rollout_launches = args.steps * args.rollout_steps * eng.launches_per_step * 2
    if not args.no_train:
        del eng, flush
        torch.cuda.empty_cache()
        from dlwp_cs_b200.train import DataParallelTrainer
        torch.manual_seed(1)
        tmodel = CubeSphereUNet2(C_PROG + C_FORC, C_PROG, base=BASE).to(dev)
        trainer = DataParallelTrainer(tmodel, lr=1e-3)
        g = torch.Generator().manual_seed(100 + rank)
        tb = args.train_batch
        ttd = torch.bfloat16 if args.train_dtype == 'bf16' else torch.float32
        xs = torch.randn(tb, 6, N_FACE, N_FACE, C_PROG + C_FORC, generator=g).to(dev).to(ttd)
        ts = torch.randn(tb, 6, N_FACE, N_FACE, C_PROG, generator=g).to(dev).to(ttd)
        for _ in range(2):
            trainer.step(xs, ts)
        barrier()
        s_ev, e_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_ev.record()
        for _ in range(args.train_steps):
            loss = trainer.step(xs, ts)
        e_ev.record()
        barrier()
        tt = torch.tensor([s_ev.elapsed_time(e_ev)], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        tms = float(tt.item()) / args.train_steps
        train = {'metric': 'train samples/sec, unet2 C%d fwd+bwd+Adam' % N_FACE, 'value': tb * world / (tms * 1e-3),
                 'unit': 'samples/s', 'ms_per_step': tms, 'global_batch': tb * world, 'batch_per_gpu': tb,
                 'dtype': args.train_dtype, 'parallelism': 'dp%d, one flat all-reduce of %d float32 gradients per step'
                                                % (world, trainer.flat.count),
                 'loss': float(loss.item()),
                 # forward + dgrad + wgrad = 3 x the forward FLOP of every layer
                 'algorithmic_gflop_per_sample': round(3e-9 * sum(
                     layer_work(s[0], s[1], s[2], s[3], N_FACE // EDGE_DIV.get(s[0], 1), 1, 2, 2)[0]
                     for s in unet2_layer_specs(C_PROG + C_FORC, C_PROG, BASE)), 3),
                 'note': ('bf16 activations: forward, dgrad and wgrad on tcgen05 kernels (fp32 accumulation in tensor '
                          'memory); float32 master weights, fused Adam')
                 if args.train_dtype == 'bf16' else 'float32 CUDA-core kernels (1e-5 parity path)'}

    if train is not None:
        train['tflops'] = round(train['value'] * train['algorithmic_gflop_per_sample'] * 1e-3, 1)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, secs = cpu_rollout_rate(args.ref_batch, args.ref_steps, threads)
        cpu = {'value': v, 'unit': UNIT, 'cores': threads, 'kind': 'port',
               'sample': 'oracle rollout_fast (torch-CPU oneDNN restatement, LUT-gather halo, fp32, host hop per step), %d members x %d '
                         'steps = %.1f s' % (args.ref_batch, args.ref_steps, secs)}

    if rank == 0:
        h2d = (h_state.numel() + h_forcing.numel()) * esz
        d2h = h_ring.numel() * esz
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': dev_ms / args.steps, 'higher_is_better': True,
                'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16' if dtype == 'bf16' else 'f32',
                'data': 'synthetic',
                'config': {'workload': 'unet2 (Weyn-2020) C%d rollout, %d vars x 2 tsteps in/out (+2 solar +2 const), ' % (N_FACE, C_PROG // 2) +
                                       '%d 6-hr steps, ensemble of %d members per GPU, random-init weights'
                                       % (args.rollout_steps, args.batch),
                           'rollout_steps': args.rollout_steps, 'batch_per_gpu': args.batch, 'face_edge': N_FACE,
                           'l2': '512 MiB flush write between bench steps; per-step activation set %.0f MiB > L2'
                                 % (args.batch * 6 * N_FACE * N_FACE * 32 * esz * 4 / 2 ** 20),
                           'cuda_graph': not args.no_graph, 'parallelism': 'replicas x%d' % world},
                'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                        'ms_per_step': e2e_ms / args.steps},
                'gpu_launches': rollout_launches,
                'clocks': clocks, 'roofline': roof, 'cpu_baseline': cpu, 'train': train, 'layers': layers}
        print(json.dumps(line), flush=True)
    if dist is not None:
        # the captured training step holds NCCL kernels: tearing the communicator down under a live CUDA graph can block
        # (observed: the process hung in destroy_process_group after the result line).  Everything is measured and printed;
        # synchronise the ranks and leave without the communicator teardown.
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == '__main__':
    main()
