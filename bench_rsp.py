#!/usr/bin/env python
"""bench_rsp.py -- row_sparse push + row_sparse_pull (BASELINE.json configs[4]): embedding table
(1 000 000, 512) fp32, 10 000 distinct hot rows per GPU (1 %), SGD on the store, then every GPU pulls
its own id list (unsorted, 5 % duplicates). Single process, all visible GPUs (the reference's
process model for kvstore('device')); with one GPU the N values all live on it.

    python bench_rsp.py [--gpus N] [--steps K]

Prints one JSON line: ms per push / pull, algorithmic GB/s (rows*(row_len*4+8) in and
union rows out for the push; unique rows * row_len*4*2 + ids*8 for a pull), and the CPU oracle
(reference semantics, src/ndarray/ndarray_function.cc:59-175) timed on the same inputs.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=0, help="0 = all visible")
    ap.add_argument("--values", type=int, default=8, help="values per push (GPUs in the reference)")
    ap.add_argument("--rows", type=int, default=1000000)
    ap.add_argument("--row-len", type=int, default=512)
    ap.add_argument("--hot", type=int, default=10000)
    ap.add_argument("--steps", type=int, default=50)
    args = ap.parse_args()
    import torch
    import anand_mxnet_b200 as mx
    import kvoracle as K
    ngpu = args.gpus or torch.cuda.device_count()
    nval = args.values
    shape = (args.rows, args.row_len)
    rng = np.random.default_rng(0xB200)
    ctxs = [mx.gpu(i % ngpu) for i in range(nval)]
    kv = mx.kv.create('device')
    # the stored weight holds every row (a dense weight in row_sparse form, as gluon creates it)
    kv.init('emb', mx.nd.sparse.row_sparse_array(
        (rng.uniform(-1, 1, shape).astype(np.float32), np.arange(args.rows, dtype=np.int64)),
        shape=shape, ctx=mx.gpu(0)))
    kv.set_optimizer(mx.optimizer.SGD(learning_rate=0.1, momentum=0.0, wd=0.0, rescale_grad=1.0 / nval))
    idx = [np.sort(rng.choice(args.rows, args.hot, replace=False)).astype(np.int64) for _ in range(nval)]
    val = [rng.uniform(-1, 1, (args.hot, args.row_len)).astype(np.float32) for _ in range(nval)]
    grads = [mx.nd.sparse.row_sparse_array((val[i], idx[i]), shape=shape, ctx=ctxs[i]) for i in range(nval)]
    pull_ids = []
    for i in range(nval):
        ids = np.concatenate([idx[i], rng.choice(idx[i], args.hot // 20)])
        rng.shuffle(ids)
        pull_ids.append(mx.nd.array(ids, ctxs[i], np.int64))
    outs = [mx.nd.sparse.zeros('row_sparse', shape, ctxs[i]) for i in range(nval)]
    union = np.unique(np.concatenate(idx))

    def push():
        kv.push('emb', grads)

    def pull():
        kv.row_sparse_pull('emb', out=outs, row_ids=pull_ids)

    for _ in range(2):
        push()
        pull()
    mx.nd.waitall()
    mx.base.reset_kernel_launch_count()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        push()
    mx.nd.waitall()
    t_push = (time.perf_counter() - t0) / args.steps
    launches_push = mx.base.kernel_launch_count() / args.steps
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pull()
    mx.nd.waitall()
    t_pull = (time.perf_counter() - t0) / args.steps
    row_bytes = args.row_len * 4
    # every source row + id read once; the union's weight rows read and written once (the summed
    # gradient is consumed in registers by the fused lazy SGD step, it is never stored)
    push_bytes = nval * args.hot * (row_bytes + 8) + len(union) * row_bytes * 2
    pull_bytes = sum(len(np.unique(p.asnumpy())) for p in pull_ids) * row_bytes * 2 + \
        sum(p.shape[0] for p in pull_ids) * 8
    # CPU oracle on the same inputs (reduce only; the reference's OMP variant is not header-callable)
    o = K.get_oracle()
    t0 = time.perf_counter()
    o.rsp_reduce(idx, val)
    t_cpu = time.perf_counter() - t0
    print(json.dumps({
        "metric": "row_sparse_push_pull", "n_gpus": ngpu, "values_per_push": nval,
        "table": list(shape), "hot_rows_per_value": args.hot, "union_rows": int(len(union)),
        "push_ms": t_push * 1e3, "pull_ms": t_pull * 1e3,
        "push_GBps": push_bytes / t_push / 1e9, "pull_GBps": pull_bytes / t_pull / 1e9,
        "kernel_launches_per_push": launches_push,
        "cpu_oracle_reduce_ms": t_cpu * 1e3, "cpu_kind": "port (single thread)",
        "note": "host-timed over %d back-to-back calls + waitall; push never blocks the host, pull "
                "blocks once per call for the unique counts (the reference blocks per output)" % args.steps}))


if __name__ == "__main__":
    main()
