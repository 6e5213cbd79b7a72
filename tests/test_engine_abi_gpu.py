"""The engine side of the C ABI (include/b200kv_c_api.h):

  * MXEnginePushSyncND / MXEnginePushAsyncND (include/mxnet/c_api.h:3323-3351): an external
    operation is ordered against KVStore work through the arrays it names;
  * deferred failures: an asynchronous operation that fails surfaces at the next wait on an array
    it was to write, or at WaitAll, exactly once (src/engine/threaded_engine.h:380-387);
  * the reference's tree-reduction switches (MXNET_KVSTORE_USETREE, MXNET_KVSTORE_TREE_ARRAY_BOUND,
    src/kvstore/comm_tree.h:54-56) are accepted and change nothing: tests/python/gpu/test_device.py:
    38-72 run as written.
"""
import ctypes
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SYNC_FN = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_void_p)
ASYNC_FN = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p)


class Ctx(ctypes.Structure):           # mxnet::Context: {dev_type, dev_id}
    _fields_ = [("dev_type", ctypes.c_int32), ("dev_id", ctypes.c_int32)]


class RunCtx(ctypes.Structure):        # mxnet::RunContext
    _fields_ = [("dev_type", ctypes.c_int32), ("dev_id", ctypes.c_int32), ("stream", ctypes.c_void_p),
                ("aux_stream", ctypes.c_void_p), ("is_bulk", ctypes.c_bool)]


@pytest.fixture(scope="module")
def mx():
    import anand_mxnet_b200 as mx
    return mx


def handles(arrs):
    return (ctypes.c_void_p * len(arrs))(*[a._hv for a in arrs])


def host_ptr(mx, a):
    p = ctypes.c_void_p()
    mx.base.check_call(mx.base._LIB.MXNDArrayGetData(a.handle, ctypes.byref(p)))
    return p.value


def push_sync(mx, fn, ctx, const, mutable, name=b"ext_op"):
    lib = mx.base._LIB
    mx.base.check_call(lib.MXEnginePushSyncND(fn, None, None, ctypes.byref(ctx), handles(const), len(const),
                                              handles(mutable), len(mutable), None, 0, name))


def push_async(mx, fn, ctx, const, mutable, name=b"ext_async_op"):
    lib = mx.base._LIB
    mx.base.check_call(lib.MXEnginePushAsyncND(fn, None, None, ctypes.byref(ctx), handles(const), len(const),
                                               handles(mutable), len(mutable), None, 0, name,
                                               ctypes.c_bool(False)))


def test_external_sync_op_orders_against_push_and_pull(mx):
    """pull -> (external host op reads the pulled weights, writes b) -> push(b): every stage sees
    the previous one's result although pull / push only enqueue work."""
    n = 1 << 20
    rng = np.random.default_rng(3)
    kv = mx.kv.create('device')
    w = rng.uniform(-1, 1, n).astype(np.float32)
    kv.init(0, mx.nd.array(w, mx.gpu(0)))
    out = mx.nd.empty((n,), mx.cpu())          # pinned host array: filled by the D2H copy lane
    b = mx.nd.zeros((n,), mx.cpu())
    seen = {}

    def ext(rctx, _param):
        rc = ctypes.cast(rctx, ctypes.POINTER(RunCtx)).contents
        seen['dev_type'] = rc.dev_type
        src = np.ctypeslib.as_array((ctypes.c_float * n).from_address(seen['out_ptr']))
        dst = np.ctypeslib.as_array((ctypes.c_float * n).from_address(seen['b_ptr']))
        dst[:] = src * 2.0 + 1.0
    fn = SYNC_FN(ext)
    for step in range(3):
        kv.pull(0, out=out)                    # asynchronous: returns before the copy has landed
        seen['out_ptr'] = None
        # pointers are taken BEFORE the pull completes on purpose: GetData flushes queued calls
        # but does not wait; the engine push must do the waiting
        seen['out_ptr'] = host_ptr(mx, out)
        seen['b_ptr'] = host_ptr(mx, b)
        push_sync(mx, fn, Ctx(1, 0), [out], [b])
        kv.push(0, b)                          # no updater: stored = b
        w = w * 2.0 + 1.0
        got = mx.nd.empty((n,), mx.gpu(0))
        kv.pull(0, out=got)
        assert np.array_equal(got.asnumpy(), w.astype(np.float32)), step
    assert seen['dev_type'] == 1


def test_external_gpu_op_receives_the_lane_stream(mx):
    """a GPU-context operation gets the compute lane's cudaStream_t: work it enqueues there is
    ordered with the library's kernels by stream order (memset through the CUDA runtime here)."""
    cudart = None
    for name in ("libcudart.so.12", "libcudart.so"):
        try:
            cudart = ctypes.CDLL(name)
            break
        except OSError:
            continue
    if cudart is None:
        import torch
        path = os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so.12")
        cudart = ctypes.CDLL(os.path.abspath(path))
    n = 4096
    a = mx.nd.ones((n,), mx.gpu(0))
    ptr = host_ptr(mx, a)          # device pointer of a
    info = {}

    def ext(rctx, on_complete, _param):
        rc = ctypes.cast(rctx, ctypes.POINTER(RunCtx)).contents
        info['stream'] = rc.stream
        info['dev'] = (rc.dev_type, rc.dev_id)
        rcode = cudart.cudaMemsetAsync(ctypes.c_void_p(ptr), 0, ctypes.c_size_t(n * 4), ctypes.c_void_p(rc.stream))
        mx.base.check_call(mx.base._LIB.B200KVEngineOnComplete(ctypes.c_void_p(on_complete),
                                                               None if rcode == 0 else b"memset failed"))
    fn = ASYNC_FN(ext)
    push_async(mx, fn, Ctx(2, 0), [], [a])
    b = a + a                      # a library kernel after the external op on the same lane
    assert info['dev'] == (2, 0) and info['stream'] == mx.base.get_stream(0)
    assert np.all(b.asnumpy() == 0) and np.all(a.asnumpy() == 0)


def test_failed_async_op_surfaces_at_wait(mx):
    a = mx.nd.ones((16,), mx.cpu())
    b = mx.nd.ones((16,), mx.cpu())
    c = mx.nd.ones((16,), mx.cpu())

    def fail(rctx, on_complete, _param):
        mx.base._LIB.B200KVEngineOnComplete(ctypes.c_void_p(on_complete), b"boom: simulated device failure")
    fn = ASYNC_FN(fail)
    push_async(mx, fn, Ctx(1, 0), [a], [b], name=b"failing_op")       # the push itself succeeds
    # an operation that consumes the failed array is skipped and inherits the failure
    ran = []
    nop = SYNC_FN(lambda rctx, p: ran.append(1))
    push_sync(mx, nop, Ctx(1, 0), [b], [c])
    assert not ran
    a.wait_to_read()                                                   # an unaffected array: fine
    with pytest.raises(mx.base.MXNetError, match="failing_op.*boom"):
        b.wait_to_read()
    b.wait_to_read()                                                   # reported once
    with pytest.raises(mx.base.MXNetError, match="boom"):
        c.wait_to_read()
    mx.nd.waitall()                                                    # nothing left to report
    # WaitAll reports a failure nobody waited for
    push_async(mx, fn, Ctx(1, 0), [], [b], name=b"failing_op2")
    with pytest.raises(mx.base.MXNetError, match="failing_op2"):
        mx.nd.waitall()
    mx.nd.waitall()
    # an operation that WRITES a failed array is skipped as well and the failure stays parked;
    # once it has been reported the array is usable again
    push_async(mx, fn, Ctx(1, 0), [], [b], name=b"failing_op3")
    push_sync(mx, nop, Ctx(1, 0), [], [b])
    assert not ran
    with pytest.raises(mx.base.MXNetError, match="failing_op3"):
        b.wait_to_read()
    push_sync(mx, nop, Ctx(1, 0), [], [b])
    assert ran
    b.wait_to_read()
    mx.nd.waitall()


def test_tree_environment_switches_are_accepted(mx, monkeypatch):
    # tests/python/gpu/test_device.py:38-72 as written (every visible GPU; one on the driver's box)
    import torch
    num_gpus = torch.cuda.device_count()
    shapes = [(1000, 1000), (1000, 1), (1,), (11, 5), (3, 17, 2), (2, 2, 3, 3), (1, 1, 1, 1, 5)]
    keys = range(len(shapes))
    gpus = range(1, 1 + num_gpus)

    def check_dense_pushpull(kv_type):
        for shape, key in zip(shapes, keys):
            for n_gpus in gpus:
                kv_device = mx.kv.create(kv_type)
                a = mx.nd.ones(shape, mx.gpu(0))
                cur_key = str(key * max(gpus) + n_gpus)
                kv_device.init(cur_key, a)
                arr_list = [mx.nd.ones(shape, mx.gpu(x)) for x in range(n_gpus)]
                res = [mx.nd.zeros(shape, mx.gpu(x)) for x in range(n_gpus)]
                kv_device.push(cur_key, arr_list)
                kv_device.pull(cur_key, res)
                for x in range(n_gpus):
                    assert np.sum(np.abs((res[x] - n_gpus).asnumpy())) == 0

    for bound in (None, '1'):
        if bound is not None:
            monkeypatch.setenv('MXNET_KVSTORE_TREE_ARRAY_BOUND', bound)
        for x in ('', '1'):
            monkeypatch.setenv('MXNET_KVSTORE_USETREE', x)
            check_dense_pushpull('local')
            check_dense_pushpull('device')
