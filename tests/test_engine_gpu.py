"""Engine / allocator ordering on the GPU (csrc/engine.cc): a block released to the pool while
device work still touches it must not be handed to another lane (copy lanes, peer GPUs) ahead of
that work."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_pool_reuse_waits_for_pending_compute_work():
    import anand_mxnet_b200 as mx
    n = 16 << 20                                   # 64 MB: the kernels below take ~100 us each
    ctx = mx.gpu(0)
    ones = np.ones(n, np.float32)
    for it in range(6):
        a = mx.nd.array(ones * (it + 1), ctx)
        b = mx.nd.zeros((n,), ctx)
        for _ in range(8):                         # queue compute-lane work that READS a
            b += a
        del a                                      # block returns to the pool with that work pending
        c = mx.nd.array(ones * -1000.0, ctx)       # same size class: re-uses the block via the H2D lane
        got = b.asnumpy()
        assert np.all(got == 8.0 * (it + 1)), (it, got[:4])
        assert np.all(c.asnumpy() == -1000.0)
