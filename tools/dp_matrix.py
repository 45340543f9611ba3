"""Data-parallel training step time under torchrun for a few collective settings (diagnostics for DESIGN.md section 7).

    python -m torch.distributed.run --nproc-per-node N tools/dp_matrix.py

Env: DP_OVERLAP=0/1 (bucketed all-reduce overlapped with backward), NCCL_MAX_NCHANNELS etc. are read by NCCL itself."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
import torch.distributed as dist
from dlwp_cs_b200 import _lib
from dlwp_cs_b200.unet import CubeSphereUNet2
from dlwp_cs_b200.train import DataParallelTrainer

rank, local, world = int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if os.environ.get('NCCL_DEBUG') and not os.environ.get('NCCL_DEBUG_FILE'):
    os.environ['NCCL_DEBUG_FILE'] = '/dev/stderr'
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
_lib.load()
torch.manual_seed(1)
model = CubeSphereUNet2(18, 14, base=32).to(dev)
tr = DataParallelTrainer(model, lr=1e-3, overlap=os.environ.get('DP_OVERLAP', '1') != '0')
g = torch.Generator().manual_seed(100 + rank)
tb = int(os.environ.get('BATCH', '32'))
xs = torch.randn(tb, 6, 48, 48, 18, generator=g).to(dev).bfloat16()
ts = torch.randn(tb, 6, 48, 48, 14, generator=g).to(dev).bfloat16()
for _ in range(5):
    tr.step(xs, ts)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
steps = 50
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    tr.step(xs, ts)
e1.record()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({'world': world, 'overlap': os.environ.get('DP_OVERLAP', '1'), 'buckets': len(tr._cuts) + 1 if tr._overlap else 1,
                      'nchannels': os.environ.get('NCCL_MAX_NCHANNELS', 'default'), 'algo': os.environ.get('NCCL_ALGO', 'default'),
                      'ms_per_step': round(float(t.item()), 4), 'samples_per_s': round(tb * world / float(t.item()) * 1e3, 1)}), flush=True)
tr.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
