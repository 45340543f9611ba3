"""bench.py on a box without a GPU: the reference arm (CPU restatement, bounded sample) prints the contract's JSON line;
the GPU arm refuses to run instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup',
                          '0', '--ref-batch', '1', '--ref-steps', '1'], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'sample-steps/s' and d['higher_is_better'] is True
    assert d['metric'] == 'forecast steps/sec C48 6-face U-Net rollout' and d['value'] > 0
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert d['gpu_launches'] == 0 and d['vs_baseline'] is None and 'workload' in d['config']


def test_gpu_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--steps', '1'], capture_output=True, text=True,
                         timeout=300)
    assert out.returncode != 0
    assert 'CUDA' in (out.stderr + out.stdout)


def test_reference_arm_under_torchrun_prints_one_line():
    """N > 1 launch of the reference arm (the driver uses torchrun for every N): rank 0 alone runs and prints, the other
    ranks exit 0 without work."""
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
                          '127.0.0.1', '--master-port', '29533', os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
                          '--gpus', '2', '--steps', '1', '--warmup', '0', '--ref-batch', '1', '--ref-steps', '1'],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['n_gpus'] == 2


def test_device_classes_refuse_cpu():
    """No CPU fallback anywhere on the product path: the engine, the trainer and the data feed raise on CPU tensors."""
    import numpy as np
    import pytest
    import torch
    from dlwp_cs_b200 import _lib
    from dlwp_cs_b200.feed import DeviceDataFeed
    from dlwp_cs_b200.train import DataParallelTrainer
    from dlwp_cs_b200.unet import CubeSphereUNet2, RolloutEngine
    model = CubeSphereUNet2(10, 6, base=8)
    with pytest.raises(_lib.DlwpcsError):
        RolloutEngine(model, 1, 8, 1, forcing_channels=4)
    with pytest.raises(_lib.DlwpcsError):
        DataParallelTrainer(model)
    with pytest.raises(_lib.DlwpcsError):
        DeviceDataFeed(np.zeros((8, 3, 6, 4, 4), np.float32), device='cpu')
    with pytest.raises(_lib.DlwpcsError):
        model(torch.zeros(1, 6, 8, 8, 10))
