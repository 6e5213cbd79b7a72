"""LARS and LAMB optimizer classes (the update_on_kvstore=False half of the path)
driven through the Updater on the GPU, against a numpy model that replays the reference's Python
orchestration (python/mxnet/optimizer/optimizer.py:934-1030, 1271-1360) with the oracle's
operator restatements. Weights agree within 1e-5 relative: the per-layer norms are floating-point
reductions whose association differs between the CPU reference and the GPU (the elementwise
operators themselves are bit-exact, tests/test_ops_gpu.py)."""
import os
import sys

import numpy as np
import pytest
from compat import mxnet_optimizer as mxopt   # the reference's optimizer front-end, mirrored (test infrastructure)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

pytestmark = pytest.mark.gpu

NAMES = ['conv0_weight', 'bn0_gamma', 'bn0_beta', 'fc_weight', 'fc_bias', 'conv1_weight']
SHAPES = [(16, 3, 3, 3), (16,), (16,), (10, 144), (10,), (32, 16, 3, 3)]


def _data(seed, dtype=np.float32):
    rng = np.random.default_rng(seed)
    ws = [rng.uniform(-1, 1, s).astype(dtype) for s in SHAPES]
    gs = [[rng.uniform(-1, 1, s).astype(dtype) for s in SHAPES] for _ in range(3)]
    return ws, gs


def _close(a, b, tol=1e-5):
    a, b = a.astype(np.float64), b.astype(np.float64)
    return np.all(np.abs(a - b) <= tol * np.abs(b) + tol * 1e-2)


def test_lars_updater_matches_model():
    import anand_mxnet_b200 as mx
    import golden_ops as G
    import kvoracle as K
    ws0, gs = _data(3)
    idx2name = dict(enumerate(NAMES))
    lrs_sched = [0.1, 0.1, 0.05]          # the third step changes lr: momentum correction kicks in

    class Sched(object):
        base_lr = 0.1

        def __call__(self, num_update):
            return lrs_sched[min(num_update - 1, 2)]
    opt = mxopt.LARS(learning_rate=0.1, momentum=0.9, wd=1e-4, eta=0.02, eps=1e-6,
                            rescale_grad=1.0 / 32, param_idx2name=idx2name, lr_scheduler=Sched())
    opt.set_wd_mult({})
    upd = mxopt.get_updater(opt)
    ctx = mx.gpu(0)
    w_nd = [mx.nd.array(w, ctx) for w in ws0]
    for step in range(3):
        g_nd = [mx.nd.array(g, ctx) for g in gs[step]]
        upd(list(range(len(NAMES))), g_nd, w_nd)
    got = [w.asnumpy() for w in w_nd]

    # ---- numpy model of the same three steps
    b = G.OracleOps()
    n = len(NAMES)
    skip = [nm.endswith(('gamma', 'beta', 'bias')) for nm in NAMES]
    wd_mult = [1.0 if nm.endswith('_weight') else 0.0 for nm in NAMES]
    ws = [w.copy() for w in ws0]
    moms = [np.zeros_like(w) for w in ws0]
    last_lr = cur_lr = None
    agg = 4
    for step in range(3):
        # Updater: one dtype group, chunks of aggregate_num tensors, each chunk = one _update_impl
        for c0 in range(0, n, agg):
            idx = list(range(c0, min(c0 + agg, n)))
            if cur_lr is not None:
                last_lr = cur_lr
            lr = lrs_sched[min(step, 2)]       # num_update == step + 1 inside the call
            if cur_lr is None:
                last_lr = lr
            cur_lr = lr
            order = [k for k in idx if not skip[k]] + [k for k in idx if skip[k]]
            nb = len([k for k in idx if not skip[k]])
            new_lrs = np.array([lr for _ in order], np.float32)
            new_wds = np.array([1e-4 * wd_mult[k] for k in order], np.float32)
            if nb:
                wsq, gsq = np.zeros(nb, np.float32), np.zeros(nb, np.float32)
                b.invoke('multi_sum_sq', [ws[k] for k in order[:nb]], [wsq], num_arrays=nb)
                b.invoke('multi_sum_sq', [gs[step][k] for k in order[:nb]], [gsq], num_arrays=nb)
                out = np.zeros(nb, np.float32)
                b.invoke('multi_lars', [new_lrs[:nb].copy(), wsq, gsq, new_wds[:nb].copy()], [out],
                         eta=0.02, eps=1e-06, rescale_grad=1.0 / 32)
                new_lrs[:nb] = out
            momentum = 0.9 * (cur_lr / last_lr)
            ins = []
            for k in order:
                ins += [ws[k], gs[step][k], moms[k]]
            ins += [new_lrs, new_wds]
            b.invoke('preloaded_multi_sgd_mom_update', ins, [ws[k] for k in order],
                     num_weights=len(order), rescale_grad=1.0 / 32, momentum=momentum)
    for k in range(n):
        assert _close(got[k], ws[k]), NAMES[k]
    # the gamma / beta / bias tensors never see a norm: bit-exact
    for k in range(n):
        if skip[k]:
            assert np.array_equal(got[k], ws[k]), NAMES[k]


@pytest.mark.parametrize("aggregate", [True, False])
@pytest.mark.parametrize("mp", [False, True])
def test_lamb_updater_matches_model(aggregate, mp, monkeypatch):
    import anand_mxnet_b200 as mx
    import golden_ops as G
    monkeypatch.setenv('MXNET_OPTIMIZER_AGGREGATION_SIZE', '45' if aggregate else '1')
    dtype = np.float16 if mp else np.float32
    ws0, gs = _data(5, dtype)
    n = len(NAMES)
    kw = dict(learning_rate=0.01, beta1=0.9, beta2=0.98, epsilon=1e-6, wd=0.01, rescale_grad=0.5,
              lower_bound=0.1, upper_bound=5.0, clip_gradient=0.4)
    opt = mxopt.LAMB(multi_precision=mp, **kw)
    upd = mxopt.get_updater(opt)
    ctx = mx.gpu(0)
    w_nd = [mx.nd.array(w, ctx, dtype) for w in ws0]
    for step in range(3):
        g_nd = [mx.nd.array(g, ctx, dtype) for g in gs[step]]
        upd(list(range(n)), g_nd, w_nd)
    got = [w.asnumpy() for w in w_nd]

    b = G.OracleOps()
    ws = [w.copy() for w in ws0]
    w32 = [w.astype(np.float32) for w in ws0]
    means = [np.zeros(s, np.float32) for s in SHAPES]
    vars_ = [np.zeros(s, np.float32) for s in SHAPES]
    for step in range(3):
        t = step + 1
        if aggregate:
            ins = []
            for k in range(n):
                ins += [ws[k], gs[step][k], means[k], vars_[k]] + ([w32[k]] if mp else [])
            b.invoke('_multi_mp_lamb_update' if mp else '_multi_lamb_update', ins, ws,
                     learning_rates=[0.01] * n, wds=[0.01] * n, beta1=0.9, beta2=0.98, epsilon=1e-06,
                     rescale_grad=0.5, bias_correction=True, num_tensors=n, step_count=[t] * n,
                     clip_gradient=0.4, lower_bound=0.1, upper_bound=5.0)
        else:
            for k in range(n):
                g = np.zeros(SHAPES[k], np.float32)
                p1 = dict(beta1=0.9, beta2=0.98, epsilon=1e-06, t=t, bias_correction=True, wd=0.01,
                          rescale_grad=0.5, clip_gradient=0.4)
                master = w32[k] if mp else ws[k]
                if mp:
                    b.invoke('mp_lamb_update_phase1', [ws[k], gs[step][k], means[k], vars_[k], w32[k]],
                             [g], **p1)
                else:
                    b.invoke('lamb_update_phase1', [ws[k], gs[step][k], means[k], vars_[k]], [g], **p1)
                r1 = np.array([np.sqrt((master.astype(np.float64) ** 2).sum())], np.float32)
                r2 = np.array([np.sqrt((g.astype(np.float64) ** 2).sum())], np.float32)
                p2 = dict(lr=0.01, lower_bound=0.1, upper_bound=5.0)
                if mp:
                    b.invoke('mp_lamb_update_phase2', [ws[k], g, r1, r2, w32[k]], [ws[k]], **p2)
                else:
                    b.invoke('lamb_update_phase2', [ws[k], g, r1, r2], [ws[k]], **p2)
    tol = 2e-3 if mp else 1e-5
    for k in range(n):
        assert _close(got[k], ws[k], tol), (NAMES[k], aggregate, mp)


def test_update_on_kvstore_false_pattern_with_lars():
    """Trainer._allreduce_grads + _update with update_on_kvstore=False (gluon/trainer.py:391-396,
    456-461): pushpull(i, grads, out=grads) then the local updater on every replica."""
    import anand_mxnet_b200 as mx
    import torch
    ndev = min(torch.cuda.device_count(), 2)
    ctxs = [mx.gpu(i) for i in range(ndev)]
    ws0, gs = _data(11)
    kv = mx.kv.create('device')
    n = len(NAMES)
    weights = [[mx.nd.array(w, c) for c in ctxs] for w in ws0]
    for i in range(n):
        kv.init(i, weights[i][0])
    opt = mxopt.LARS(learning_rate=0.1, momentum=0.9, wd=1e-4, eta=0.02,
                            rescale_grad=1.0 / (16 * ndev), param_idx2name=dict(enumerate(NAMES)))
    upds = [mxopt.get_updater(opt) for _ in ctxs]
    for step in range(2):
        grads = [[mx.nd.array(gs[step][i] * (d + 1), c) for d, c in enumerate(ctxs)] for i in range(n)]
        for i in range(n):
            kv.pushpull(i, grads[i], out=grads[i], priority=-i)
        for d, upd in enumerate(upds):
            upd(list(range(n)), [grads[i][d] for i in range(n)], [weights[i][d] for i in range(n)])
    # replicas stay identical
    for i in range(n):
        a = weights[i][0].asnumpy()
        for d in range(1, ndev):
            assert np.array_equal(a, weights[i][d].asnumpy()), NAMES[i]
        assert np.all(np.isfinite(a))
        assert not np.array_equal(a, ws0[i])
