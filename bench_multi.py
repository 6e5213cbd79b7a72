"""bench_multi.py -- the N>1 arm of bench.py: one rank per GPU (torchrun), kvstore('device') inside a
libb200kv peer group. Every rank pushes its own ResNet-50 gradient set and pulls the updated
weights; per step each rank launches ONE fused kernel that reduces the stripes it owns over
IPC-mapped NVLink peer memory, applies SGD-momentum, and writes the new weights into every rank's
weight arrays (reduce-scatter + update + all-gather in one kernel, no NCCL on the data path).

value  = whole-job push+pull payload rate: N x 2 x gradient-set bytes / time (same definition as
         the N=1 arm, so the per-N values are comparable)
roofline.achieved = that bus bandwidth PER GPU against 900 GB/s/dir NVLink 5 (measured peer copy
         770 GB/s/dir, B200_PROFILING.md)
"""
import ctypes
import json
import os

import numpy as np


def run_multi_gpu(args):
    import torch
    import torch.distributed as dist
    import anand_mxnet_b200 as mx
    from bench import (WORKLOADS, METRIC, UNIT, SGD_KW, ADAM_KW, ClockSampler, make_optimizer,
                       algorithmic_bytes, payload_bytes, VALUE_FORMULA)
    from anand_mxnet_b200.kvstore.base import _ctype_key_value

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

    shapes = WORKLOADS[args.workload]["shapes"]()
    keys = list(range(len(shapes)))
    n_elem = sum(int(np.prod(s)) for s in shapes)
    ctx = mx.gpu(local)
    # arrays are created THROUGH the library, i.e. inside this rank's IPC arena: peers address them
    # directly (no staging copies); torch only fills them with random numbers through DLPack views
    gen = torch.Generator(device=dev)

    def lib_arrays(seed):
        arrs = [mx.nd.empty(s, ctx) for s in shapes]
        if seed is not None:
            gen.manual_seed(seed)
            for a in arrs:
                mx.nd.to_torch(a).uniform_(-1, 1, generator=gen)
        return arrs

    weights0 = lib_arrays(0xB200 + 777)          # identical initial weights on every rank
    grads = lib_arrays(0xB200 + 1000 * rank)     # per-rank gradients
    outs = lib_arrays(None)
    torch.cuda.synchronize()
    kv = mx.kv.create("device")
    assert kv.rank == rank and kv.num_workers == world
    kv.init(keys, weights0)
    kv.set_optimizer(make_optimizer(mx, args.workload, world))
    kv.pushpull(keys, grads, out=outs)           # plan once (collective), hands lr/multipliers over
    ckeys, cvals, _ = _ctype_key_value(keys, grads)
    _, couts, _ = _ctype_key_value(keys, outs)
    lib, handle, nkeys, zero = mx.base._LIB, kv.handle, ctypes.c_uint(len(keys)), ctypes.c_int(0)

    def step():
        rc = lib.MXKVStorePushPull(handle, nkeys, ckeys, nkeys, ckeys, cvals, couts, zero)
        if rc != 0:
            raise RuntimeError(lib.MXGetLastError().decode())

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    mx.base.reset_kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    dist.barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    dist.barrier()
    launches = mx.base.kernel_launch_count()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)        # max over ranks, device-timed
    ms_step = ms.item() / args.steps

    # ---- end to end: host (pinned) gradients in, weights back to host, same call
    rng = np.random.default_rng(0xB200 + rank)
    hgrads = [mx.nd.array(rng.uniform(-1, 1, s).astype(np.float32), mx.cpu()) for s in shapes]
    houts = [mx.nd.empty(s, mx.cpu()) for s in shapes]
    kv2 = mx.kv.create("device")
    kv2.init(keys, weights0)
    kv2.set_optimizer(make_optimizer(mx, args.workload, world))
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        kv2.pushpull(keys, hgrads, out=houts)
    mx.nd.waitall()
    torch.cuda.synchronize()
    dist.barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(stream)
    for _ in range(e2e_steps):
        kv2.pushpull(keys, hgrads, out=houts)
    mx.nd.waitall()
    f1.record(stream)
    torch.cuda.synchronize()
    ems = torch.tensor([f0.elapsed_time(f1)], device=dev)
    dist.all_reduce(ems, op=dist.ReduceOp.MAX)
    e2e_ms = ems.item() / e2e_steps

    clocks = sampler.stop() if rank == 0 else None
    bus_per_gpu = algorithmic_bytes(args.workload, world)      # bytes per GPU per step
    if rank == 0:
        busbw = bus_per_gpu / (ms_step * 1e-3) / 1e9
        pay = payload_bytes(args.workload, world)
        line = {
            "metric": METRIC, "value": pay / (ms_step * 1e-3) / 1e9, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload]["desc"], "store": "kvstore('device')",
                       "parallelism": "one rank per GPU, stripes of the key space owned round-robin, "
                                      "fused reduce-scatter+update+all-gather kernel over IPC peer memory",
                       "value_formula": VALUE_FORMULA,
                       "bus_bandwidth_GBps_per_gpu": busbw,
                       "bus_bandwidth_formula": "size * 2(n-1)/n / time (tools/bandwidth/measure.py:137)",
                       "l2": "per-GPU working set > 126 MB L2, no flush needed",
                       "numa_bound": mx.dist._state.get('numa_node') is not None,
                       "nvls_in_switch_reduce": bool(mx.dist.nvls_wanted(world) and mx.dist.has_multicast()),
                       "parity_mode": ("1e-6 relative (in-switch summation order)"
                                       if mx.dist.nvls_wanted(world) and mx.dist.has_multicast()
                                       else "bit-exact vs the reference CPU store"),
                       "optimizer": SGD_KW if WORKLOADS[args.workload]["opt"] == "sgd" else ADAM_KW},
            "roofline": {"bound": "nvlink", "achieved": busbw, "peak": 900.0, "unit": "GB/s",
                         "frac": busbw / 900.0, "frac_of_measured_770": busbw / 770.0,
                         "traffic": None, "kernel": "dense_fused_kernel<float,%d,SGD>" % world,
                         "note": "bus bandwidth per GPU = size*2(n-1)/n/time vs 900 GB/s/dir nominal"},
            "cpu_baseline": None,
            "e2e": {"value": pay / (e2e_ms * 1e-3) / 1e9, "unit": UNIT,
                    "h2d_bytes_per_step": n_elem * 4, "d2h_bytes_per_step": n_elem * 4,
                    "ms_per_step": e2e_ms, "steps": e2e_steps, "note": "bytes per rank"},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line))
    dist.barrier()
    mx.nd.waitall()
    mx.dist.destroy_peer_group()
    dist.destroy_process_group()
