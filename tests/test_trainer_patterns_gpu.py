"""The call patterns of the reference's training loops against the store:
  * update_on_kvstore=False: kv.pushpull(i, grads) all-reduces IN PLACE, then one Updater per device
    applies multi_sgd_mom_update over aggregated tensors (gluon/trainer.py:391-396,456-461);
  * learning-rate changes between steps (Trainer.set_learning_rate) and lr_scheduler;
  * optimizer-state checkpoints (KVStore.save/load_optimizer_states, optimizer.py:2143-2161) in both
    routes, interchangeable between them.
Checker: the CPU oracle model; all comparisons bit-exact."""
import os

import numpy as np
import pytest

import kvoracle as K
from compat import mxnet_optimizer as mxopt   # the reference's optimizer front-end, mirrored (test infrastructure)

pytestmark = pytest.mark.gpu
SHAPES = [(64, 3, 7, 7), (64,), (3,), (128, 64, 3, 3), (1000, 512), (1000,)]


@pytest.fixture(scope="module")
def mx():
    import anand_mxnet_b200 as mx
    return mx


def eq(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.shape == b.shape and np.array_equal(a.view(np.uint8), b.view(np.uint8))


def rnd(rng, s):
    return rng.uniform(-1, 1, s).astype(np.float32)


def test_allreduce_in_place_then_local_updaters(mx, oracle):
    """update_on_kvstore=False: the store only sums (no optimizer on it); every 'device' then runs
    its own Updater -> multi_sgd_mom_update over up to aggregate_num=4 tensors per launch."""
    rng = np.random.default_rng(40)
    ndev = 3                                   # three value slots on one GPU stand for three devices
    kv = mx.kv.create('device')
    keys = list(range(len(SHAPES)))
    w0 = [rnd(rng, s) for s in SHAPES]
    kv.init(keys, [mx.nd.array(w, mx.gpu(0)) for w in w0])
    weights = [[mx.nd.array(w, mx.gpu(0)) for _ in range(ndev)] for w in w0]
    opt = mxopt.SGD(learning_rate=0.05, momentum=0.9, wd=1e-3, rescale_grad=1.0 / 96)
    updaters = [mxopt.get_updater(opt) for _ in range(ndev)]
    ref_w = [w.copy() for w in w0]
    ref_m = [np.zeros_like(w) for w in w0]
    for step in range(3):
        g = [[rnd(rng, s) for _ in range(ndev)] for s in SHAPES]
        grads = [[mx.nd.array(a, mx.gpu(0)) for a in gs] for gs in g]
        for i in keys:                                        # Trainer._allreduce_grads
            kv.pushpull(i, grads[i], priority=-i)             # out=None: result lands in grads
        for d, upd in enumerate(updaters):                    # Trainer._update
            upd(list(keys), [grads[i][d] for i in keys], [weights[i][d] for i in keys])
        for i in keys:
            merged = oracle.reduce(g[i], 'device').reshape(SHAPES[i])
            for d in range(ndev):
                assert eq(grads[i][d].asnumpy(), merged), ("allreduce", step, i, d)
            oracle.multi_sgd_update(ref_w[i].reshape(-1), merged.reshape(-1), ref_m[i].reshape(-1),
                                    K.f32(0.05), K.scalar_param(0.9), K.f32(1e-3),
                                    K.scalar_param(1.0 / 96))
            for d in range(ndev):
                assert eq(weights[i][d].asnumpy(), ref_w[i]), ("update", step, i, d)


@pytest.mark.parametrize("fused", ['1', '0'])
def test_learning_rate_changes_and_scheduler(mx, fused, monkeypatch):
    monkeypatch.setenv('B200KV_FUSED_OPTIMIZER', fused)

    class Sched(object):          # python/mxnet/lr_scheduler.py FactorScheduler-like
        base_lr = 0.2

        def __call__(self, num_update):
            return self.base_lr * (0.5 ** (num_update // 2))
    rng = np.random.default_rng(41)
    for use_sched in (False, True):
        kv = mx.kv.create('device')
        model = K.LocalKVStoreModel('device')
        keys = list(range(len(SHAPES)))
        for k, s in enumerate(SHAPES):
            w = rnd(rng, s)
            kv.init(k, mx.nd.array(w, mx.gpu(0)))
            model.init(k, w)
        if use_sched:
            opt = mxopt.SGD(lr_scheduler=Sched(), momentum=0.9, wd=1e-4)
        else:
            opt = mxopt.SGD(learning_rate=0.1, momentum=0.9, wd=1e-4)
        kv.set_optimizer(opt)
        outs = [mx.nd.empty(s, mx.gpu(0)) for s in SHAPES]
        for step in range(5):
            if use_sched:
                lr = Sched()(step + 1)
            else:
                lr = 0.1 if step < 2 else 0.01
                opt.set_learning_rate(lr)        # Trainer.set_learning_rate between steps
            model.set_optimizer('sgd', lr=lr, momentum=0.9, wd=1e-4)
            grads = [[rnd(rng, s) for _ in range(2)] for s in SHAPES]
            kv.pushpull(keys, [[mx.nd.array(a, mx.gpu(0)) for a in gs] for gs in grads], out=outs)
            for k in keys:
                model.push(k, grads[k])
                assert eq(outs[k].asnumpy(), model.pull(k)), (use_sched, step, k)


def test_optimizer_state_checkpoint_roundtrip(mx, tmp_path, monkeypatch):
    """states saved by the fused route load into the callback route (and back) and training
    continues bit-identically: the pickle layout is the reference's {index: state}."""
    rng = np.random.default_rng(42)
    w0 = [rnd(rng, s) for s in SHAPES]
    grads = [[[rnd(rng, s) for _ in range(2)] for s in SHAPES] for _ in range(4)]
    keys = list(range(len(SHAPES)))

    def make(fused):
        monkeypatch.setenv('B200KV_FUSED_OPTIMIZER', fused)
        kv = mx.kv.create('device')
        kv.init(keys, [mx.nd.array(w, mx.gpu(0)) for w in w0])
        kv.set_optimizer(mxopt.SGD(learning_rate=0.1, momentum=0.9, wd=1e-4))
        return kv

    def run(kv, steps, outs):
        for st in steps:
            kv.pushpull(keys, [[mx.nd.array(a, mx.gpu(0)) for a in gs] for gs in grads[st]], out=outs)
    outs_a = [mx.nd.empty(s, mx.gpu(0)) for s in SHAPES]
    kv_a = make('1')
    run(kv_a, [0, 1, 2, 3], outs_a)                       # uninterrupted reference run (fused)
    want = [o.asnumpy() for o in outs_a]
    for first, second in (('1', '0'), ('0', '1')):
        outs = [mx.nd.empty(s, mx.gpu(0)) for s in SHAPES]
        kv1 = make(first)
        run(kv1, [0, 1], outs)
        f = str(tmp_path / ("states_%s" % first))
        kv1.save_optimizer_states(f)
        mid = [o.asnumpy() for o in outs]
        monkeypatch.setenv('B200KV_FUSED_OPTIMIZER', second)
        kv2 = mx.kv.create('device')
        kv2.init(keys, [mx.nd.array(w, mx.gpu(0)) for w in mid])   # weights saved by the caller
        kv2.set_optimizer(mxopt.SGD(learning_rate=0.1, momentum=0.9, wd=1e-4))
        kv2.load_optimizer_states(f)
        run(kv2, [2, 3], outs)
        for k in keys:
            assert eq(outs[k].asnumpy(), want[k]), (first, second, k)
