"""CPU-only checks of host-side policy code: launch-support helpers of anand_mxnet_b200.dist, the
bench's byte accounting, the optimizer front-end's bookkeeping for the LARS / LAMB classes (no
arithmetic runs here -- the operators need a GPU)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_nvls_rule(monkeypatch):
    import anand_mxnet_b200 as mx
    monkeypatch.delenv('B200KV_NVLS', raising=False)
    assert [mx.dist.nvls_wanted(w) for w in (2, 4, 8, 16)] == [False, False, True, True]
    monkeypatch.setenv('B200KV_NVLS', 'auto')
    assert mx.dist.nvls_wanted(8) and not mx.dist.nvls_wanted(4)
    monkeypatch.setenv('B200KV_NVLS', '1')
    assert mx.dist.nvls_wanted(2)
    monkeypatch.setenv('B200KV_NVLS', '0')
    assert not mx.dist.nvls_wanted(8)


def test_numa_binding_is_best_effort(monkeypatch):
    """no GPU here: the helper must leave the affinity alone and return None, never raise"""
    import anand_mxnet_b200 as mx
    before = os.sched_getaffinity(0)
    assert mx.dist.bind_to_gpu_numa_node(0) is None
    assert os.sched_getaffinity(0) == before
    monkeypatch.setenv('B200KV_BIND_NUMA', '0')
    assert mx.dist.bind_to_gpu_numa_node(0) is None


def test_bench_byte_accounting():
    import bench
    n = sum(int(np.prod(s)) for s in bench.resnet50_shapes())
    assert n == 25549486 and len(bench.resnet50_shapes()) == 157
    nb = sum(int(np.prod(s)) for s in bench.bert_base_shapes())
    assert nb == 109482240 and len(bench.bert_base_shapes()) == 199
    # value: push + pull payload, same definition at every N
    for w in (1, 2, 4, 8):
        assert bench.payload_bytes('resnet50_sgd', w) == w * n * 4 * 2
    # roofline: 24 B / element at N=1 (SGD momentum), bus bandwidth bytes per GPU at N>=2
    assert bench.algorithmic_bytes('resnet50_sgd', 1) == n * 24
    assert bench.algorithmic_bytes('bert_adam', 1) == nb * 32
    assert bench.algorithmic_bytes('resnet50_sgd', 8) == int(n * 4 * 2 * 7 / 8)


def test_lars_and_lamb_bookkeeping():
    import anand_mxnet_b200 as mx
    names = {0: 'conv_weight', 1: 'bn_gamma', 2: 'fc_bias'}
    opt = mx.optimizer.create('lars', learning_rate=0.1, momentum=0.9, wd=1e-4, param_idx2name=names)
    opt.set_wd_mult({})
    assert opt._get_wds([0, 1, 2]) == [1e-4, 0.0, 0.0]          # only *_weight parameters decay
    opt._update_count([0, 1, 2])
    assert opt._get_lrs([0, 1, 2]) == [0.1, 0.1, 0.1]
    assert opt.last_lr == 0.1 and opt.cur_lr == 0.1
    opt.lr = 0.05
    opt._get_lrs([0])
    assert (opt.last_lr, opt.cur_lr) == (0.1, 0.05)            # what the momentum correction uses
    lamb = mx.optimizer.create('lamb', learning_rate=0.01)
    assert lamb.aggregate_num == 45 and lamb.epsilon == 1e-6 and lamb.bias_correction
    assert isinstance(mx.optimizer.get_updater(lamb), mx.optimizer.Updater)
