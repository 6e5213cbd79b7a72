"""bench_multi.py -- the N>1 arm of bench.py: one rank per GPU (torchrun), kvstore('device') inside a
libb200kv peer group. Every rank pushes its own gradient set and pulls the updated weights; per step
each rank launches ONE fused kernel that reduces the stripes it owns over IPC-mapped NVLink peer
memory, applies the optimizer, and writes the new weights into every rank's weight arrays
(reduce-scatter + update + all-gather in one kernel, no NCCL on the data path).

value  = whole-job push+pull payload rate: N x 2 x gradient-set bytes / time (same definition as
         the N=1 arm, so the per-N values are comparable)
roofline.achieved = all-reduce bus bandwidth PER GPU (size * 2(N-1)/N / time, tools/bandwidth/
         measure.py:137-138) against 900 GB/s/dir NVLink 5 (measured peer copy 770 GB/s/dir)
parity = the weights every rank pulled after the first two steps against the CPU oracle's
         kvstore('local') model fed ALL ranks' seeded gradients (rank 0 regenerates them from the
         seeds): bit-exact in the peer-load mode, <= 1e-6 relative L1 when the NVSwitch sums
configs = short legs over the other BASELINE configs at this N (BERT-base + Adam; row_sparse table
         sharded by row range over the ranks)
"""
import json
import os
import sys

import numpy as np


def _max_over_ranks(torch, dist, dev, ms):
    t = torch.tensor([ms], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def _all_ok(torch, dist, dev, ok):
    t = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(t.item())


def run_dense(mx, torch, dist, stream, local, rank, world, workload, steps, warmup, full):
    from bench import (WORKLOADS, UNIT, ClockSampler, make_optimizer, algorithmic_bytes, payload_bytes,
                       flat_set, weight_seed, grad_seed, oracle_expected, compare_sets, c_step_fn,
                       per_key_step_fns, time_region, kernel_label)
    import gc
    dev = torch.device("cuda", local)
    # every rank starts a workload from the same allocator state (the NVLS mode needs the operands at
    # the same arena offsets on all ranks): release what earlier legs left behind, everywhere
    gc.collect()
    mx.nd.waitall()
    dist.barrier()
    shapes = WORKLOADS[workload]["shapes"]()
    sizes = [int(np.prod(s)) for s in shapes]
    keys = list(range(len(shapes)))
    n_elem = sum(sizes)
    ctx = mx.gpu(local)
    # arrays are created THROUGH the library, i.e. inside this rank's IPC arena: peers address them
    # directly (no staging copies)
    w0 = flat_set(weight_seed(), sizes)              # identical initial weights on every rank
    g0 = flat_set(grad_seed(rank), sizes)            # per-rank gradients
    weights0 = [mx.nd.array(w0[k].reshape(shapes[k]), ctx) for k in keys]
    grads = [mx.nd.array(g0[k].reshape(shapes[k]), ctx) for k in keys]
    outs = [mx.nd.empty(s, ctx) for s in shapes]
    torch.cuda.synchronize()
    kv = mx.kv.create("device")
    assert kv.rank == rank and kv.num_workers == world
    kv.init(keys, weights0)
    kv.set_optimizer(make_optimizer(mx, workload, world))
    # ---- parity: the first two steps (the first one also plans the launch collectively)
    for _ in range(2):
        kv.pushpull(keys, grads, out=outs)
    mx.nd.waitall()
    # did the launch really sum in the switch? (the library falls back to peer loads when the ranks'
    # operands do not sit at the same arena offsets) -- ask it, do not assume
    nvls = "nvls" in mx.base.last_kernel_info()[0]
    nvls = _all_ok(torch, dist, dev, nvls)
    got = [o.asnumpy() for o in outs]
    want = oracle_expected(workload, world, 2) if rank == 0 else None
    # every rank holds the same pulled weights: rank 0 checks its own against the oracle, the
    # others check theirs against rank 0's bits
    parity = None
    flat_got = torch.from_numpy(np.concatenate([g.reshape(-1) for g in got])).to(dev)
    ref0 = flat_got.clone()
    dist.broadcast(ref0, 0)
    same_as_rank0 = bool(torch.equal(ref0.view(torch.int32), flat_got.view(torch.int32)))
    all_same = _all_ok(torch, dist, dev, same_as_rank0)
    if rank == 0:
        parity = compare_sets(got, want, exact=not nvls)
        parity["after_steps"] = 2
        parity["all_ranks_hold_identical_weights"] = all_same
        parity["ok"] = bool(parity["ok"] and all_same)
    del flat_got, ref0

    step = c_step_fn(mx, kv, keys, grads, outs)
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    mx.base.reset_kernel_launch_count()
    torch.cuda.synchronize()
    dist.barrier()
    ms = time_region(torch, stream, step, steps)
    dist.barrier()
    launches = mx.base.kernel_launch_count()
    ms_step = _max_over_ranks(torch, dist, dev, ms)           # max over ranks, device-timed
    clocks = sampler.stop() if rank == 0 else None
    bus_per_gpu = algorithmic_bytes(workload, world)           # bytes per GPU per step
    busbw = bus_per_gpu / (ms_step * 1e-3) / 1e9
    pay = payload_bytes(workload, world)
    res = {"workload": WORKLOADS[workload]["desc"], "ms_per_step": ms_step, "steps": steps,
           "value": pay / (ms_step * 1e-3) / 1e9, "unit": UNIT, "gpu_launches": int(launches),
           "nvls_in_switch_reduce": nvls,
           "roofline": {"bound": "nvlink", "achieved": busbw, "peak": 900.0, "unit": "GB/s",
                        "frac": busbw / 900.0, "frac_of_measured_770": busbw / 770.0, "traffic": None,
                        "kernel": kernel_label(workload, world, nvls),
                        "note": "bus bandwidth per GPU = size*2(n-1)/n/time vs 900 GB/s/dir nominal"},
           "parity": parity, "clocks": clocks}
    if not full:
        del kv
        return res

    # ---- the per-key call patterns of the reference's callers (Trainer / measure.py)
    fe = {}
    trainer_pattern, measure_pattern, _ = per_key_step_fns(mx, kv, keys, grads, outs)

    def flushed(fn):
        def g():
            fn()
            mx.base.flush_all()
        return g
    for name, fn in (("per_key_pushpull_priority_minus_i", trainer_pattern),
                     ("per_key_push_all_then_pull_all", measure_pattern)):
        try:
            for _ in range(3):
                flushed(fn)()
            mx.nd.waitall()
            torch.cuda.synchronize()
            dist.barrier()
            t = time_region(torch, stream, flushed(fn), steps)
            t = _max_over_ranks(torch, dist, dev, t)
            fe[name] = {"ms_per_step": t, "vs_grouped_call": t / ms_step}
        except Exception as e:
            fe[name] = {"error": str(e)[:200]}
            break                       # ranks may have diverged: do not issue more collectives
    res["frontends"] = fe

    # ---- end to end: host (pinned) gradients in, weights back to host, same call
    hgrads = [mx.nd.array(g0[k].reshape(shapes[k]), mx.cpu()) for k in keys]
    houts = [mx.nd.empty(s, mx.cpu()) for s in shapes]
    kv2 = mx.kv.create("device")
    kv2.init(keys, weights0)
    kv2.set_optimizer(make_optimizer(mx, workload, world))
    e2e_steps = max(3, min(steps, 10))
    for _ in range(2):
        kv2.pushpull(keys, hgrads, out=houts)
    mx.nd.waitall()
    e2e_parity = None
    if rank == 0:
        # (the staged arm may or may not sum in the switch: the tolerant bound covers both)
        e2e_parity = compare_sets([h.asnumpy() for h in houts], want,
                                  exact=not (mx.dist.nvls_wanted(world) and mx.dist.has_multicast()))
    hstep = c_step_fn(mx, kv2, keys, hgrads, houts)
    hstep()
    mx.nd.waitall()
    torch.cuda.synchronize()
    dist.barrier()
    ems = time_region(torch, stream, hstep, e2e_steps, after=mx.nd.waitall)
    e2e_ms = _max_over_ranks(torch, dist, dev, ems)
    per_dir = n_elem * 4 / (e2e_ms * 1e-3) / 1e9
    res["e2e"] = {"value": pay / (e2e_ms * 1e-3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": n_elem * 4,
                  "d2h_bytes_per_step": n_elem * 4, "ms_per_step": e2e_ms, "steps": e2e_steps,
                  "note": "bytes per rank", "parity": e2e_parity,
                  "roofline": {"bound": "pcie", "achieved": per_dir, "unit": "GB/s per direction per GPU",
                               "peak": None, "frac": None,
                               "note": "the N=1 line measures the pinned-copy peak of the box (both "
                                       "directions busy); all N ranks share the host's memory system"}}
    if os.environ.get("B200KV_E2E_DIAG"):
        # where does the host-buffer step spend its time? one direction at a time
        diag = {}
        dgr = [mx.nd.array(g0[k].reshape(shapes[k]), ctx) for k in keys]
        dou = [mx.nd.empty(s, ctx) for s in shapes]
        for name, vv, oo in (("h2d_only(host grads -> device weights)", hgrads, dou),
                             ("d2h_only(device grads -> host weights)", dgr, houts)):
            kvd = mx.kv.create("device")
            kvd.init(keys, weights0)
            kvd.set_optimizer(make_optimizer(mx, workload, world))
            for _ in range(2):
                kvd.pushpull(keys, vv, out=oo)
            st = c_step_fn(mx, kvd, keys, vv, oo)
            st()
            mx.nd.waitall()
            torch.cuda.synchronize()
            dist.barrier()
            t = time_region(torch, stream, st, e2e_steps, after=mx.nd.waitall)
            diag[name] = _max_over_ranks(torch, dist, dev, t)
        res["e2e"]["diagnostics_ms"] = diag
    return res


def run_rsp_group(mx, torch, dist, stream, local, rank, world, steps):
    """BASELINE configs[4] in the rank-per-GPU store: every rank pushes ITS 10 000 rows of the
    (1M, 512) table; the table and the optimizer state are sharded by row range over the ranks; then
    every rank pulls its own id list. All ranks check their pulled rows against the oracle."""
    import bench_rsp as R
    import kvoracle as K
    dev = torch.device("cuda", local)
    ctx = mx.gpu(local)
    shape = (R.ROWS, R.ROW_LEN)
    w = R.table_init()
    idx, val, pull = R.make_inputs(world)
    kv = mx.kv.create('device')
    kv.init('emb', mx.nd.sparse.row_sparse_array((w, np.arange(R.ROWS, dtype=np.int64)), shape=shape, ctx=ctx))
    kv.set_optimizer(mx.optimizer.SGD(learning_rate=R.LR, momentum=0.0, wd=0.0, rescale_grad=1.0 / world))
    grad = mx.nd.sparse.row_sparse_array((val[rank], idx[rank]), shape=shape, ctx=ctx)
    ids = mx.nd.array(pull[rank], ctx, np.int64)
    out = mx.nd.sparse.zeros('row_sparse', shape, ctx)

    def push():
        kv.push('emb', grad)

    def pull_fn():
        kv.row_sparse_pull('emb', out=out, row_ids=ids)
    for _ in range(2):
        push()
    pull_fn()
    mx.nd.waitall()
    o = K.get_oracle()
    gi, sub = R.oracle_rows(w, idx, val, 2, world)
    ok = R.check_pull(o, gi, sub, w, pull[rank], out.indices.asnumpy(), out.data.asnumpy())
    ok = _all_ok(torch, dist, dev, ok)
    torch.cuda.synchronize()
    dist.barrier()
    push_ms = _max_over_ranks(torch, dist, dev, R._time(torch, stream, push, steps, mx.nd.waitall))
    dist.barrier()
    pull_ms = _max_over_ranks(torch, dist, dev, R._time(torch, stream, pull_fn, steps, mx.nd.waitall))
    pb, lb, union = R.alg_bytes(idx, pull)
    del kv
    return {"workload": "row_sparse push + row_sparse_pull, table (%d, %d) fp32 sharded by row range over "
                        "%d ranks, %d rows per rank, lazy SGD on the store" % (R.ROWS, R.ROW_LEN, world, R.HOT),
            "union_rows": int(union), "push_ms": push_ms, "pull_ms": pull_ms, "steps": steps,
            "value": (pb + lb) / ((push_ms + pull_ms) * 1e-3) / 1e9, "unit": "GB/s (algorithmic, whole job)",
            "roofline": {"bound": "hbm+nvlink", "unit": "GB/s", "peak": None, "traffic": None,
                         "push": {"algorithmic_bytes": pb, "achieved": pb / (push_ms * 1e-3) / 1e9},
                         "pull": {"algorithmic_bytes": lb, "achieved": lb / (pull_ms * 1e-3) / 1e9},
                         "note": "whole-job algorithmic bytes; each rank moves 1/N of them, (N-1)/N of "
                                 "which cross NVLink"},
            "parity": {"mode": "bit-exact (ids and rows, every rank checks its own pull)", "ok": bool(ok),
                       "after_pushes": 2}}


def run_multi_gpu(args):
    os.environ.setdefault("B200KV_IPC_ARENA_MB", "16384")
    import torch
    import torch.distributed as dist
    import anand_mxnet_b200 as mx
    from bench import METRIC, UNIT, config_block, opt_kwargs

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    assert world == args.gpus, "launch with torchrun --nproc-per-node %d" % args.gpus
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    mx.dist.init_peer_group(local)
    stream = torch.cuda.Stream(device=dev)      # shared by torch events and the library's kernels
    torch.cuda.set_stream(stream)
    mx.base.set_stream(local, stream.cuda_stream)
    assert mx.base.get_stream(local) == stream.cuda_stream
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle"))

    if args.workload == "rsp":
        leg = run_rsp_group(mx, torch, dist, stream, local, rank, world, args.steps)
        if rank == 0:
            print(json.dumps({"metric": "kvstore_row_sparse_push_pull_GBps", "value": leg["value"],
                              "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": 2,
                              "ms_per_step": leg["push_ms"] + leg["pull_ms"], "higher_is_better": True,
                              "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                              "config": config_block("rsp", world), "impl_detail": leg,
                              "roofline": leg["roofline"], "parity": leg["parity"], "cpu_baseline": None,
                              "gpu_launches": None}))
        ok = leg["parity"]["ok"]
    else:
        main = run_dense(mx, torch, dist, stream, local, rank, world, args.workload, args.steps,
                         args.warmup, full=True)
        configs = {}
        if not args.no_config_legs:
            other = "bert_adam" if args.workload == "resnet50_sgd" else "resnet50_sgd"
            configs[other] = run_dense(mx, torch, dist, stream, local, rank, world, other,
                                       max(3, min(args.steps, 10)), 3, full=False)
            configs["rsp"] = run_rsp_group(mx, torch, dist, stream, local, rank, world,
                                           max(5, min(args.steps, 20)))
        ok = True
        if rank == 0:
            line = {
                "metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": main["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config_block(args.workload, world),
                "impl_detail": {
                    "store": "kvstore('device')",
                    "parallelism": "one rank per GPU, stripes of the key space owned round-robin, fused "
                                   "reduce-scatter+update+all-gather kernel over IPC peer memory",
                    "bus_bandwidth_GBps_per_gpu": main["roofline"]["achieved"],
                    "bus_bandwidth_formula": "size * 2(n-1)/n / time (tools/bandwidth/measure.py:137)",
                    "numa_bound": mx.dist._state.get('numa_node') is not None,
                    "nvls_in_switch_reduce": main["nvls_in_switch_reduce"],
                    "optimizer": opt_kwargs(args.workload, world)[1]},
                "roofline": main["roofline"], "parity": main["parity"], "cpu_baseline": None,
                "e2e": main["e2e"], "gpu_launches": main["gpu_launches"], "clocks": main["clocks"],
                "frontends": main["frontends"], "configs": configs}
            print(json.dumps(line))
            ok = bool(main["parity"]["ok"] and (main["e2e"]["parity"] or {}).get("ok", True) and all(
                (c.get("parity") or {}).get("ok", True) for c in configs.values()))
            if not ok:
                sys.stderr.write("bench.py: PARITY FAILURE against the oracle\n")
    okt = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(okt, 0)
    dist.barrier()
    mx.nd.waitall()
    mx.dist.destroy_peer_group()
    dist.destroy_process_group()
    if not bool(okt.item()):
        sys.exit(3)
