"""GPU parity of the device-side data feed: DeviceDataFeed / dlwpcs_feed_gather against the
golden outputs of the reference's own ArrayDataGenerator.generate and against the oracle restatement at C48 size."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import cs_feed  # noqa: E402
from tests.golden.cases import FEED_CASES  # noqa: E402


def test_feed_vs_reference_golden(golden_dir):
    from dlwp_cs_b200.feed import DeviceDataFeed
    g = np.load(os.path.join(golden_dir, 'feed.npz'))
    for k, kw in FEED_CASES.items():
        feed = DeviceDataFeed(g['array'], batch_size=4, input_slice=kw['input_slice'], output_slice=kw['output_slice'],
                              input_time_steps=kw['t_in'], output_time_steps=kw['t_out'], interval=kw['interval'],
                              insolation_array=g['insolation_array'], constants=g['constants'])
        assert feed._n_sample == int(g['n_sample_' + k])
        x, y = feed.generate(g['samples_' + k])
        assert np.array_equal(x.cpu().numpy(), np.concatenate([g['p_' + k], g['const_' + k]], axis=-1))
        assert np.array_equal(y.cpu().numpy(), g['t_' + k])
        x0, y0 = feed[0]                                   # first batch of the un-shuffled epoch = samples 0..3
        xr, yr = cs_feed.generate(g['array'], np.arange(min(4, feed._n_sample)), insolation_array=g['insolation_array'],
                                  constants=g['constants'], **kw)
        assert np.array_equal(x0.cpu().numpy(), xr) and np.array_equal(y0.cpu().numpy(), yr)


def test_feed_c48_bf16_vs_oracle():
    """Training shapes: C48, 7 variables x 2 time steps + insolation + 2 constants -> 18 input / 14 output channels."""
    from dlwp_cs_b200.feed import DeviceDataFeed
    rng = np.random.default_rng(3)
    T, V, N = 20, 7, 48
    array = rng.standard_normal((T, V, 6, N, N)).astype(np.float32)
    sol = rng.random((T, 6, N, N)).astype(np.float32)
    consts = rng.random((2, 6, N, N)).astype(np.float32)
    feed = DeviceDataFeed(array, batch_size=8, input_time_steps=2, output_time_steps=2, interval=2, shuffle=True,
                          insolation_array=sol, constants=consts, dtype=torch.bfloat16, drop_remainder=True, seed=1)
    assert feed.n_input_channels == 18 and feed.n_output_channels == 14 and len(feed) == feed._n_sample // 8
    samples = feed._indices[:8]
    x, y = feed[0]
    xr, yr = cs_feed.generate(array, samples, slice(None), slice(None), 2, 2, 2, insolation_array=sol, constants=consts)
    assert torch.equal(x.cpu(), torch.from_numpy(xr).bfloat16()) and torch.equal(y.cpu(), torch.from_numpy(yr).bfloat16())
    with pytest.raises(IndexError):
        feed.generate([feed._n_sample])


def test_feed_sequence_vs_reference_golden(golden_dir):
    """sequence=S mode (Azure/train_cs.py:159-163): predictors of step 0, insolation of the later steps, list of targets --
    against the reference generator's own outputs."""
    from dlwp_cs_b200.feed import DeviceDataFeed
    from tests.golden.cases import FEED_SEQ_CASES
    g = np.load(os.path.join(golden_dir, 'feed.npz'))
    for k, (kw, S) in FEED_SEQ_CASES.items():
        feed = DeviceDataFeed(g['array'], batch_size=4, input_slice=kw['input_slice'], output_slice=kw['output_slice'],
                              input_time_steps=kw['t_in'], output_time_steps=kw['t_out'], interval=kw['interval'],
                              insolation_array=g['insolation_array'], constants=g['constants'], sequence=S)
        assert feed._n_sample == int(g['n_sample_' + k])
        x, solars, targets = feed.generate(g['samples_' + k])
        assert len(solars) == S - 1 and len(targets) == S
        assert np.array_equal(x.cpu().numpy(), np.concatenate([g['p_' + k], g['const_' + k]], axis=-1))
        for s in range(1, S):
            assert np.array_equal(solars[s - 1].cpu().numpy(), g['solar_%s_%d' % (k, s)])
        for s in range(S):
            assert np.array_equal(targets[s].cpu().numpy(), g['t_%s_%d' % (k, s)])


def test_feed_rank_shards_are_disjoint_and_indices_checked():
    """Data-parallel feed: same seed on every rank, every world-th sample of the shuffled epoch order."""
    from dlwp_cs_b200.feed import DeviceDataFeed
    rng = np.random.default_rng(5)
    array = rng.standard_normal((23, 3, 6, 4, 4)).astype(np.float32)
    feeds = [DeviceDataFeed(array, batch_size=2, input_time_steps=2, output_time_steps=2, shuffle=True, seed=7,
                            drop_remainder=True, rank=r, world=4) for r in range(4)]
    shards = [set(f._indices.tolist()) for f in feeds]
    assert all(len(s) == feeds[0]._n_sample // 4 for s in shards)
    assert len(set().union(*shards)) == 4 * len(shards[0])          # disjoint
    assert len({len(f) for f in feeds}) == 1                         # same number of batches on every rank
    with pytest.raises(IndexError):
        DeviceDataFeed(array, input_slice=[0, 3])
    f = DeviceDataFeed(array, input_slice=[-1, 0], output_slice=[2])
    assert f.in_idx == [2, 0]
    with pytest.raises(ValueError):
        DeviceDataFeed(array, rank=2, world=2)
