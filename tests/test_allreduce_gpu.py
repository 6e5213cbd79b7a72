"""NCCL fallback store (kv.create('allreducestore')): torch.distributed all-reduce of a packed bucket
+ the local fused update. NCCL chooses the summation order, so the bar is the reference's own
nightly bound: sum|delta| / sum|ref| < 1e-6 (tests/nightly/test_kvstore.py:95-98)."""
import os
import socket

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]
SHAPES = [(4, 4), (100, 100), (3,), (70001, 3)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _grad(rank, step, k, shape):
    return np.random.default_rng(500 * rank + 13 * step + k).uniform(-1, 1, shape).astype(np.float32)


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "oracle")]
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    errors = []
    try:
        import kvoracle as K
        import anand_mxnet_b200 as mx
        ctx = mx.gpu(rank)
        kv = mx.kv.create('allreducestore')
        assert kv.rank == rank and kv.num_workers == world
        model = K.LocalKVStoreModel('device')
        keys = list(range(len(SHAPES)))
        for k, s in enumerate(SHAPES):
            w = np.random.default_rng(7 + k).uniform(-1, 1, s).astype(np.float32)
            kv.init(k, mx.nd.array(w, ctx))
            model.init(k, w)
        kv.set_optimizer(mx.optimizer.SGD(learning_rate=0.1, momentum=0.9, wd=1e-4,
                                          rescale_grad=1.0 / (32 * world)))
        model.set_optimizer('sgd', lr=0.1, momentum=0.9, wd=1e-4, rescale_grad=1.0 / (32 * world))
        outs = [mx.nd.empty(s, ctx) for s in SHAPES]
        for step in range(3):
            vals = [mx.nd.array(_grad(rank, step, k, s), ctx) for k, s in enumerate(SHAPES)]
            kv.pushpull(keys, vals, out=outs)
            for k, s in enumerate(SHAPES):
                model.push(k, [_grad(r, step, k, s) for r in range(world)])
                got, want = outs[k].asnumpy(), model.pull(k)
                err = np.abs(got - want).sum() / np.abs(want).sum()
                if not err < 1e-6:
                    errors.append("step %d key %d rel L1 %g" % (step, k, err))
        mx.nd.waitall()
    except Exception:  # noqa
        import traceback
        errors.append(traceback.format_exc())
    finally:
        q.put((rank, errors))
        dist.destroy_process_group()


def test_allreduce_fallback_store():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = []
    try:
        results = [q.get(timeout=300) for _ in procs]
    finally:
        for p in procs:
            p.join(timeout=30)
            if p.is_alive():
                p.kill()
    for rank, errors in results:
        assert not errors, (rank, errors[:3])
