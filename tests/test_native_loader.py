"""Native C++ host prefetcher (csrc/host_loader.h) vs plain torch indexing."""
import pytest
import torch

from eventgrad_b200.data import BatchLoader, ShardSampler, synthetic_source
from eventgrad_b200.data.native_loader import NativeHostBatches, native_available

pytestmark = pytest.mark.skipif(not native_available(), reason="extension not built")


def test_batches_match_torch_indexing_and_keep_partial_batch():
    src = synthetic_source("cifar10", 333)
    nh = NativeHostBatches(src, 64, n_slots=3)
    order = torch.randperm(333, generator=torch.Generator().manual_seed(1))
    seen = 0
    for b, (x, y, slot) in enumerate(nh.batches(order)):
        idx = order[b * 64:(b + 1) * 64]
        assert torch.equal(x, src.images[idx]) and torch.equal(y, src.labels[idx])
        seen += x.shape[0]
        nh.release(slot)
    assert seen == 333 and b == 5                                  # 5 full + 1 partial batch of 13
    # a second epoch on the same object, different order
    order2 = torch.arange(332, -1, -1)
    xs = []
    for x, y, s in nh.batches(order2):
        xs.append(x.clone())          # copy out BEFORE handing the slot back to the worker
        nh.release(s)
    assert torch.equal(torch.cat(xs), src.images[order2])
    nh.close()


def test_backpressure_and_out_of_range():
    src = synthetic_source("mnist", 100)
    nh = NativeHostBatches(src, 10, n_slots=2)
    it = nh.batches(torch.arange(100))
    x0, y0, s0 = next(it)
    x1, y1, s1 = next(it)                                          # both slots in use: the worker must wait
    assert torch.equal(x0, src.images[0:10]) and torch.equal(x1, src.images[10:20])
    nh.release(s0); nh.release(s1)
    rest = []
    for x, y, s in it:
        rest.append((x.clone(), None))
        nh.release(s)
    assert len(rest) == 8 and torch.equal(rest[-1][0], src.images[90:100])
    nh.start_epoch(torch.tensor([0, 1, 500]))                      # index 500 does not exist
    with pytest.raises(IndexError):
        nh.next()
    nh.close()


def test_batchloader_native_equals_python_path():
    src = synthetic_source("mnist", 500)
    a = BatchLoader(src, ShardSampler(500, 2, 1, "random"), 32, "cpu", native="on")
    b = BatchLoader(src, ShardSampler(500, 2, 1, "random"), 32, "cpu", native="off")
    assert a.native is not None and b.native is None
    for (xa, ya), (xb, yb) in zip(a, b):
        assert torch.equal(xa, xb) and torch.equal(ya, yb)
    assert len(list(a)) == len(list(b)) == 8
