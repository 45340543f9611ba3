"""Host-side logic of the data-parallel training step on CPU: flat parameter / gradient buffers, batch sharding and the
single all-reduce (gloo, world_size 2).  The CUDA kernels themselves are covered by tests/test_gpu_train.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dlwp_cs_b200.train import FlatBuffers, shard_batch


def _toy():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Tanh(), torch.nn.Linear(7, 3))


def test_flat_buffers_alias_parameters_and_grads():
    net = _toy()
    before = [p.detach().clone() for p in net.parameters()]
    flat = FlatBuffers(net)
    assert flat.count == sum(p.numel() for p in net.parameters())
    for p, b in zip(net.parameters(), before):
        assert torch.equal(p.detach(), b)
        assert p.data.untyped_storage().data_ptr() == flat.param.untyped_storage().data_ptr()
        assert p.grad.untyped_storage().data_ptr() == flat.grad.untyped_storage().data_ptr()
    x = torch.randn(4, 5)
    net(x).pow(2).mean().backward()
    ref = torch.cat([g.reshape(-1) for g in torch.autograd.grad(net(x).pow(2).mean(), list(net.parameters()))])
    assert torch.allclose(flat.grad, ref, atol=1e-7)          # autograd accumulated in place into the flat buffer
    flat.zero_grad()
    assert float(flat.grad.abs().sum()) == 0.0 and all(float(p.grad.abs().sum()) == 0.0 for p in net.parameters())
    with torch.no_grad():
        flat.param.add_(1.0)                                   # an optimizer acting on the flat buffer moves the module
    assert all(torch.allclose(p.detach(), b + 1.0) for p, b in zip(net.parameters(), before))


@pytest.mark.parametrize('gb,world', [(256, 8), (10, 4), (3, 4), (32, 1)])
def test_shard_batch_partitions(gb, world):
    spans = [shard_batch(gb, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == gb
    assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
    sizes = [hi - lo for lo, hi in spans]
    assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, gb, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        net = _toy()
        flat = FlatBuffers(net)
        if rank != 0:
            with torch.no_grad():
                flat.param.mul_(0.0)                           # replicas start different; rank 0's values must win
        flat.broadcast_params(0)
        g = torch.Generator().manual_seed(1)
        x, t = torch.randn(gb, 5, generator=g), torch.randn(gb, 3, generator=g)
        lo, hi = shard_batch(gb, rank, world)
        flat.zero_grad()
        (net(x[lo:hi]) - t[lo:hi]).pow(2).mean().backward()    # mean over the shard, like DataParallelTrainer
        w = flat.all_reduce()
        assert w == world
        if rank == 0:
            torch.save({'grad': flat.grad / w, 'param': flat.param.clone()}, out)
    finally:
        dist.destroy_process_group()


def test_gloo_all_reduce_matches_single_process(tmp_path):
    world, gb = 2, 8
    out = str(tmp_path / 'r0.pt')
    mp.spawn(_worker, args=(world, _free_port(), gb, out), nprocs=world, join=True)
    got = torch.load(out)
    net = _toy()
    flat = FlatBuffers(net)
    g = torch.Generator().manual_seed(1)
    x, t = torch.randn(gb, 5, generator=g), torch.randn(gb, 3, generator=g)
    (net(x) - t).pow(2).mean().backward()                      # equal shards: mean of shard means == global mean
    assert torch.equal(got['param'], flat.param)
    assert torch.allclose(got['grad'], flat.grad, atol=1e-6)


def test_flat_buffers_collect_grads_cpu():
    """FlatBuffers.collect_grads: gradients handed over by autograd (p.grad cleared before backward) end up in the flat
    buffer, p.grad points at the flat views again, parameters without a gradient read as zero."""
    import torch
    from dlwp_cs_b200.train import FlatBuffers
    torch.manual_seed(0)
    m = torch.nn.Sequential(torch.nn.Linear(5, 4), torch.nn.Linear(4, 3))
    unused = torch.nn.Parameter(torch.ones(2))
    m.register_parameter('unused', unused)
    fb = FlatBuffers(m)
    x = torch.randn(7, 5)
    ref = torch.autograd.grad(m(x).pow(2).sum(), [p for p in m.parameters() if p is not unused])
    fb.grad.fill_(123.0)
    for p in fb.params:
        p.grad = None
    m(x).pow(2).sum().backward()
    fb.collect_grads()
    off = 0
    it = iter(ref)
    for p, v in zip(fb.params, fb.grad_views):
        assert p.grad.data_ptr() == v.data_ptr()
        n = p.numel()
        want = torch.zeros_like(p) if p is unused else next(it)
        assert torch.equal(fb.grad[off:off + n].view(p.shape), want)
        off += n


def test_bucket_split_and_partial_collect():
    """Two gradient buckets in reverse layer order (the tail of the flat buffer = the layers whose gradients backward
    produces first); collect_grads(stop=k) copies the head only and re-points the tail."""
    from dlwp_cs_b200.unet import CubeSphereUNet2
    model = CubeSphereUNet2(18, 14, base=32)
    flat = FlatBuffers(model)
    k, off = flat.bucket_split()
    names = [n for n, _ in model.named_parameters()]
    assert names[k] == 'conv_2d_6_2.equatorial_kernel' and off == sum(flat.sizes[:k]) and off % 4 == 0
    assert 0.3 <= (flat.count - off) / flat.count <= 0.5
    cuts = flat.bucket_cuts()
    assert cuts[0] == (k, off) and len(cuts) == 2 and names[cuts[1][0]] == 'conv_2d_2_2.equatorial_kernel'
    assert cuts[1][1] == sum(flat.sizes[:cuts[1][0]]) and cuts[1][1] / flat.count < 0.15       # small exposed head bucket
    for p in flat.params:
        p.grad = torch.ones_like(p)                      # fresh tensors, not views of the flat buffer
    flat.grad.zero_()
    flat.collect_grads(stop=k)
    assert float(flat.grad[:off].min()) == 1.0 and float(flat.grad[off:].abs().max()) == 0.0
    assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(flat.params, flat.grad_views))
