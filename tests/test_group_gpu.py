"""One-rank-per-GPU store (CUDA-IPC peer group, in-kernel barriers): parity against the oracle.
Spawns one process per GPU (2, and 4 when available) with a gloo bootstrap group on 127.0.0.1."""
import os
import socket

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]

SHAPES = [(4, 4), (100, 100), (3,), (1027,), (70001, 3), (1500, 1500)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _grad(rank, step, k, shape):
    return np.random.default_rng(1000 * rank + 17 * step + k).uniform(-1, 1, shape).astype(np.float32)


def _worker(rank, world, port, q, nvls=False):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "oracle")]
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["LOCAL_RANK"] = str(rank)
    os.environ["B200KV_IPC_ARENA_MB"] = "512"
    os.environ["B200KV_NVLS"] = "1" if nvls else "0"
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    errors = []
    try:
        import kvoracle as K
        import anand_mxnet_b200 as mx
        mx.dist.init_peer_group(rank)
        ctx = mx.gpu(rank)
        keys = list(range(len(SHAPES)))

        in_switch = nvls and mx.dist.has_multicast()

        def eq(a, b):
            a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
            if in_switch:
                # the NVSwitch chooses the summation order: the reference's own bound applies
                # (tests/nightly/test_kvstore.py:95-98: sum|delta| / sum|ref| < 1e-6)
                return a.shape == b.shape and \
                    np.abs(a.astype(np.float64) - b).sum() <= 1e-6 * max(np.abs(b).sum(), 1e-30)
            return np.array_equal(a.view(np.uint8), b.view(np.uint8))
        for kvtype, order in (("device", "device"), ("local", "local")):
            # ---- plain reduce + broadcast (no optimizer)
            kv = mx.kv.create(kvtype)
            if not (kv.rank == rank and kv.num_workers == world):
                errors.append("rank/num_workers wrong")
            kv.init(keys, [mx.nd.zeros(s, ctx) for s in SHAPES])
            outs = [mx.nd.empty(s, ctx) for s in SHAPES]
            for step in range(2):
                vals = [mx.nd.array(_grad(rank, step, k, s), ctx) for k, s in enumerate(SHAPES)]
                kv.pushpull(keys, vals, out=outs)
                for k, s in enumerate(SHAPES):
                    want = K.get_oracle().reduce([_grad(r, step, k, s) for r in range(world)], order)
                    if not eq(outs[k].asnumpy().ravel(), want):
                        errors.append("reduce %s step %d key %d" % (kvtype, step, k))
            # ---- fused SGD-momentum, state striped over the ranks
            kv = mx.kv.create(kvtype)
            model = K.LocalKVStoreModel(order)
            for k, s in enumerate(SHAPES):
                w = np.random.default_rng(99 + k).uniform(-1, 1, s).astype(np.float32)
                kv.init(k, mx.nd.array(w, ctx))
                model.init(k, w)
            kv.set_optimizer(mx.optimizer.SGD(learning_rate=0.1, momentum=0.9, wd=1e-4,
                                              rescale_grad=1.0 / (64 * world), clip_gradient=0.01))
            model.set_optimizer('sgd', lr=0.1, momentum=0.9, wd=1e-4,
                                rescale_grad=1.0 / (64 * world), clip_gradient=0.01)
            # one gradient comes from torch memory (outside the IPC arena): staged automatically
            for step in range(3):
                vals = []
                for k, s in enumerate(SHAPES):
                    g = _grad(rank, 10 + step, k, s)
                    if k == 1:
                        vals.append(mx.nd.from_torch(torch.from_numpy(g).cuda()))
                    else:
                        vals.append(mx.nd.array(g, ctx))
                torch.cuda.synchronize()
                kv.pushpull(keys, vals, out=outs)
                for k, s in enumerate(SHAPES):
                    model.push(k, [_grad(r, 10 + step, k, s) for r in range(world)])
                    if not eq(outs[k].asnumpy(), model.pull(k)):
                        errors.append("sgd %s step %d key %d" % (kvtype, step, k))
            fresh = [mx.nd.empty(s, ctx) for s in SHAPES]
            kv.pull(keys, out=fresh)
            for k in keys:
                if not eq(fresh[k].asnumpy(), model.pull(k)):
                    errors.append("pull %s key %d" % (kvtype, k))
        # ---- row_sparse key: every rank owns a row range of the table (weight + state), reads the
        # peers' gradients through IPC, pulls gather rows from the owning ranks (bit-exact)
        for optname in ('sgd_mom', 'adam'):
            shape = (997, 16)
            w = np.random.default_rng(71).uniform(-1, 1, shape).astype(np.float32)
            kv = mx.kv.create('device')
            kv.init('emb', mx.nd.array(w, ctx).tostype('row_sparse'))
            if optname == 'sgd_mom':
                kv.set_optimizer(mx.optimizer.SGD(learning_rate=0.1, momentum=0.9, wd=1e-3, rescale_grad=0.5))
            else:
                kv.set_optimizer(mx.optimizer.Adam(learning_rate=1e-3, wd=0.01, rescale_grad=0.5))
            sp = K.scalar_param
            m, v = np.zeros(shape, np.float32), np.zeros(shape, np.float32)
            o = K.get_oracle()

            def rsp_of(r, t):
                rg = np.random.default_rng(500 + 31 * r + t)
                n = 0 if (r == 1 and t == 2) else 60          # one rank pushes an all-zero gradient once
                idx = np.sort(rg.choice(shape[0], n, replace=False)).astype(np.int64)
                return idx, rg.uniform(-1, 1, (n, shape[1])).astype(np.float32)
            for t in range(1, 4):
                idx, val = rsp_of(rank, t)
                if len(idx):
                    g = mx.nd.sparse.row_sparse_array((val, idx), shape=shape, ctx=ctx)
                else:
                    g = mx.nd.sparse.zeros('row_sparse', shape, ctx)
                kv.push('emb', g)
                parts = [rsp_of(r, t) for r in range(world)]
                parts = [p for p in parts if len(p[0])]
                gi, gv = o.rsp_reduce([p[0] for p in parts], [p[1] for p in parts])
                if optname == 'sgd_mom':
                    o.sgd_mom_rsp_update(w, m, gi, gv, sp(0.1), sp(0.9), sp(1e-3), sp(0.5), None)
                else:
                    o.adam_rsp_update(w, m, v, gi, gv, sp(K.adam_lr(1e-3, 0.9, 0.999, t)), sp(0.9),
                                      sp(0.999), sp(1e-8), sp(0.01), sp(0.5), None)
                ids = np.random.default_rng(900 + rank + 7 * t).integers(0, shape[0], 250)
                out = mx.nd.sparse.zeros('row_sparse', shape, ctx)
                kv.row_sparse_pull('emb', out=out, row_ids=mx.nd.array(ids, ctx, np.int64))
                u = np.unique(ids)
                if not np.array_equal(out.indices.asnumpy(), u):
                    errors.append("rsp %s step %d ids" % (optname, t))
                elif not np.array_equal(out.data.asnumpy().view(np.uint32), w[u].view(np.uint32)):
                    errors.append("rsp %s step %d rows" % (optname, t))
        # kv.init: "only the value supplied by worker with rank 0 is used" (kvstore.py:136-141)
        kv = mx.kv.create('device')
        vals0 = [np.random.default_rng(5 + k).uniform(-1, 1, s).astype(np.float32)
                 for k, s in enumerate(SHAPES)]
        kv.init(keys, [mx.nd.array(v + np.float32(rank), ctx) for v in vals0])
        got = [mx.nd.empty(s, ctx) for s in SHAPES]
        kv.pull(keys, out=got)
        for k in keys:
            if not eq(got[k].asnumpy(), vals0[k]):
                errors.append("init broadcast key %d" % k)
        mx.nd.waitall()
        dist.barrier()
        mx.dist.destroy_peer_group()
    except Exception as e:  # noqa
        import traceback
        errors.append(traceback.format_exc())
    finally:
        q.put((rank, errors))
        if nvls and rank == 0:
            print('NVLS multicast in use:', 'in_switch' in dir() and in_switch)
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nvls", [(2, False), (4, False), (2, True), (4, True)])
def test_peer_group_parity(world, nvls):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, nvls)) for r in range(world)]
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
