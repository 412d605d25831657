import numpy as np
import pytest
import torch

from shallowspeed_b200.dataset import Dataset, synthetic_mnist, write_reference_files


def test_dp_shard_length_and_dtype():
    # reference tests/test_dataset.py: shard length == (59500 - 59500 % 128) // 4, fp32
    ds = Dataset(None, global_batch_size=128, mubatch_size=8).load(DP_rank=1, DP_size=4)
    assert len(ds) == (59500 - 59500 % 128) // 4
    assert ds.input_X.dtype == torch.float32 and ds.target_y.dtype == torch.float32
    assert ds.get_num_batches() == 464 and ds.get_num_mubatches() == 4
    assert ds.load_micro_batch_input(0, 0).shape == (8, 784)
    assert ds.load_micro_batch_target(463, 3).shape == (8, 10)


def test_dp_layouts_see_the_same_global_batch():
    full = Dataset(None, 128, 32).load(0, 1)
    shards = [Dataset(None, 128, 8).load(r, 4) for r in range(4)]
    for b in (0, 7, 463):
        xb, yb = full.load_batch(b)
        got = torch.empty_like(xb)
        for r, sh in enumerate(shards):
            got[r::4] = sh.load_batch(b)[0]
        assert torch.equal(got, xb)
        mus = torch.cat([full.load_micro_batch_input(b, m) for m in range(4)])
        assert torch.equal(mus, xb)


def test_validation_split_and_determinism():
    val = Dataset(None, 128, 128, validation=True).load(0, 1)
    assert len(val) == 10500 - 10500 % 128 and val.get_num_batches() == 82
    x1, y1 = synthetic_mnist(validation=False, n=1000)
    x2, y2 = synthetic_mnist(validation=False, n=1000)
    assert np.array_equal(x1, x2) and np.array_equal(y1, y2)
    assert y1.sum(axis=1).min() == 1.0 and y1.shape == (1000, 10)


def test_reference_file_format_roundtrip(tmp_path):
    d = write_reference_files(tmp_path / "mnist_784", n=512)
    ds = Dataset(d, 128, 32).load(0, 1)
    assert not ds.synthetic and len(ds) == 512
    x, _ = synthetic_mnist(validation=False, n=512)
    assert np.allclose(ds.input_X.numpy(), x)


def test_bad_configs_assert():
    with pytest.raises(AssertionError):
        Dataset(None, 128, 32).load(0, 3)          # 128 % 3 != 0
    with pytest.raises(AssertionError):
        Dataset(None, 128, 24).load(0, 1)          # mubatch must divide the local batch
