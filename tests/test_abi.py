"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/b200kv_c_api.h declares, fails loudly without a GPU (no CPU fallback), and its host-only
pieces (chunk planner, plugin registry, front-end bookkeeping) behave."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="module")
def mx():
    import anand_mxnet_b200 as mx
    return mx


def test_library_exports_every_declared_symbol(mx):
    hdr = open(os.path.join(ROOT, "include", "b200kv_c_api.h")).read()
    names = set(re.findall(r"B200KV_DLL\s+[\w\s\*]+?\b((?:MX|NN|B200KV)\w+)\s*\(", hdr))
    assert len(names) > 70, len(names)
    # the KVStore entry points of the reference's C API that python/mxnet/kvstore/kvstore.py binds
    for required in ("MXKVStoreCreate", "MXKVStoreFree", "MXKVStoreInit", "MXKVStoreInitEx",
                     "MXKVStorePush", "MXKVStorePushEx", "MXKVStorePullWithSparse",
                     "MXKVStorePullWithSparseEx", "MXKVStorePushPull", "MXKVStorePushPullEx",
                     "MXKVStorePullRowSparse", "MXKVStorePullRowSparseEx", "MXKVStoreSetUpdaterEx",
                     "MXKVStoreSetGradientCompression", "MXKVStoreGetType", "MXKVStoreGetRank",
                     "MXKVStoreGetGroupSize", "MXKVStoreIsWorkerNode", "MXKVStoreBarrier",
                     "MXKVStoreSendCommmandToServers", "MXGetLastError"):
        assert required in names
    lib = ctypes.CDLL(os.path.join(ROOT, "anand_mxnet_b200", "libb200kv.so"))
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU behaviour")
def test_fails_loudly_without_gpu(mx):
    with pytest.raises(mx.MXNetError, match="no CPU fallback"):
        mx.kv.create('device')
    with pytest.raises(mx.MXNetError, match="needs a CUDA device"):
        mx.nd.zeros((2, 2))


def test_error_convention(mx):
    lib = mx.base._LIB
    h = ctypes.c_void_p()
    assert lib.NNGetOpHandle(b"Convolution", ctypes.byref(h)) == -1
    assert b"not registered" in lib.MXGetLastError()
    assert lib.NNGetOpHandle(b"sgd_mom_update", ctypes.byref(h)) == 0
    v = ctypes.c_int()
    assert lib.MXGetVersion(ctypes.byref(v)) == 0 and v.value == 10600
    for fn in (lib.MXKVStoreIsWorkerNode, lib.MXKVStoreIsServerNode, lib.MXKVStoreIsSchedulerNode):
        assert fn(ctypes.byref(v)) == 0
    assert lib.MXKVStoreIsWorkerNode(ctypes.byref(v)) == 0 and v.value == 1


def test_chunk_planner(mx):
    """Host-only planner: every element of every key lands in exactly one chunk of <= 4096
    elements; stripes of 32768 elements rotate over the devices; balance within one stripe/key."""
    from bench import resnet50_shapes, bert_base_shapes
    lib = mx.base._LIB
    for shapes in (resnet50_shapes(), bert_base_shapes(), [(3,), (1,), (4097,), (32768 * 3 + 5,)]):
        sizes = [int(np.prod(s)) for s in shapes]
        arr = (ctypes.c_uint64 * len(sizes))(*sizes)
        for ndev in (1, 2, 4, 8):
            counts = (ctypes.c_uint64 * ndev)()
            elems = (ctypes.c_uint64 * ndev)()
            assert lib.B200KVTestPlanChunks(arr, len(sizes), ndev, counts, elems) == 0
            assert sum(elems) == sum(sizes)
            if sum(sizes) > 64 * 32768 * ndev:
                share = sum(sizes) / ndev
                assert max(elems) - min(elems) <= 0.05 * share, (ndev, list(elems))


def test_registry_and_teststore(mx):
    # python/mxnet/kvstore/base.py:221-245,441-455: registry first, native factory otherwise
    @mx.kv.KVStoreBase.register
    class MyStore(mx.kv.KVStoreBase):
        pass
    assert isinstance(mx.kv.create('MyStore'), MyStore)
    assert mx.kv.create('teststore').type == 'teststore'
    assert mx.kv.TestStore.is_capable('optimizer') is False
    assert mx.kv.KVStore.is_capable('optimizer') is True
    with pytest.raises(TypeError):
        mx.kv.create(3)


def test_optimizer_bookkeeping(mx):
    """the hyper-parameter record kv.set_optimizer takes: per-parameter lr / wd multipliers with the
    reference's lookup order (optimizer.py:432-509), no device work involved."""
    opt = mx.optimizer.SGD(learning_rate=0.1, wd=0.01, param_idx2name={0: 'fc_weight', 1: 'fc_bias'})
    opt.set_lr_mult({'fc_bias': 2.0})
    assert opt.multipliers(0) == (1.0, 1.0)
    assert opt.multipliers(1) == (2.0, 0.0)            # biases get wd_mult 0 by name
    opt.set_wd_mult({1: 0.5})                            # an index entry beats the name rule
    assert opt.multipliers(1) == (2.0, 0.5)

    class P(object):
        lr_mult, wd_mult = 3.0, 4.0
    opt.param_dict = {0: P()}                            # a Parameter object beats both tables
    assert opt.multipliers(0) == (3.0, 4.0)
    adam = mx.optimizer.create('adam', learning_rate=3e-4)
    assert adam.lr == 3e-4 and adam.kind == 'adam' and adam.beta2 == 0.999
    assert adam.op_params()['epsilon'] == 1e-8 and mx.optimizer.fused_kind(adam) == 'adam'
    with pytest.raises(ValueError):
        mx.optimizer.create('lars')                      # no fused kernel: callback route only
    with pytest.raises(TypeError):
        mx.optimizer.get_updater(opt)                    # a record carries no update rule


def test_optimizer_record_rules(mx):
    """the rest of the record's contract: defaults, scheduler handling, objects of the reference's
    classes (tests/compat mirror) read as records, pickling without Parameter objects"""
    import pickle
    from compat import mxnet_optimizer as mxopt

    class Sched(object):
        base_lr = 0.5

        def __call__(self, num_update):
            return self.base_lr / (1 + num_update)
    sgd = mx.optimizer.SGD(lr_scheduler=Sched(), momentum=0.9)
    assert sgd.lr is None and sgd.learning_rate == 0.5          # the scheduler supplies the rate
    sgd.num_update = 4
    assert sgd.learning_rate == 0.1
    with pytest.raises(UserWarning):
        sgd.set_learning_rate(0.3)                              # optimizer.py:356-365
    s2 = Sched()
    explicit = mx.optimizer.SGD(learning_rate=0.2, lr_scheduler=s2)
    assert s2.base_lr == 0.2 and explicit.lr == 0.2             # an explicit rate wins (optimizer.py:117-123)
    with pytest.raises(TypeError):
        mx.optimizer.SGD(beta1=0.9)                              # not an SGD parameter
    p = mx.optimizer.Adam(clip_gradient=None, wd=0.01).op_params()
    assert p['clip_gradient'] == 0.0 and p['wd'] == 0.01 and p['lazy_update'] is True
    assert set(mx.optimizer.Test().op_params()) == {'learning_rate', 'wd', 'rescale_grad', 'clip_gradient',
                                                    'multi_precision', 'begin_num_update'}
    # an object of the reference's SGD class is recognised by class name + attributes ...
    ref_sgd = mxopt.SGD(learning_rate=0.05, momentum=0.8, wd=1e-3, param_idx2name={0: 'a_weight', 1: 'a_bias'})
    assert mx.optimizer.fused_kind(ref_sgd) == 'sgd'
    rec = mx.optimizer.record_of(ref_sgd)
    assert rec.kind == 'sgd' and rec.momentum == 0.8 and rec.multipliers(1) == (1.0, 0.0)
    assert rec.op_params()['learning_rate'] == 0.05
    # ... one of another class is not (callback route), and get_updater wraps its own update rule
    lars = mxopt.LARS(learning_rate=0.1)
    assert mx.optimizer.fused_kind(lars) is None
    upd = mx.optimizer.get_updater(lars)
    assert upd.optimizer is lars and upd.states == {}
    # Parameter objects do not travel with a pickled record (optimizer.py:512-519)

    class P(object):
        lr_mult, wd_mult = 2.0, 3.0
    adam = mx.optimizer.Adam(param_dict={7: P()})
    assert adam.multipliers(7) == (2.0, 3.0)
    back = pickle.loads(pickle.dumps(adam))
    assert back.param_dict == {} and back.kind == 'adam' and back.beta1 == 0.9
    # the callback updater's pickle: {index: state}, or (states, optimizer)
    upd.states = {3: None}
    assert pickle.loads(upd.get_states()) == {3: None}
    st, opt = pickle.loads(upd.get_states(dump_optimizer=True))
    assert st == {3: None} and type(opt).__name__ == 'LARS'
