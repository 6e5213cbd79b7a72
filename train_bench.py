#!/usr/bin/env python
"""train_bench.py -- ResNet-50 samples/sec with the B200 KVStore as the gradient synchroniser
(BASELINE.json configs[2]: "ResNet-50 Gluon Trainer bf16, batch 256/GPU, kvstore('device')").

    python train_bench.py                                   # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 train_bench.py --gpus 8

The forward/backward is plain torch (torchvision ResNet-50, bf16, channels_last, synthetic images):
plumbing, not the product. The measured path is what the reference's Trainer does after backward
(gluon/trainer.py:371-396): kv.pushpull(i, grads, out=weights) for every parameter with the
optimizer ON THE STORE -- here one grouped call = one fused kernel per GPU that reduces the bf16
gradients of all ranks over NVLink peer memory, applies SGD-momentum to the fp32 master weights
(the reference's multi_precision semantics, optimizer.py:584-594; bf16 instead of fp16 is an
extension) and writes the new bf16 weights straight into every rank's model parameters.

Parameters and gradients live in arrays created through the library (the IPC arena when N>1), and
torch sees them through zero-copy DLPack views, so autograd writes gradients where the peers read.

Reported: samples/s with the device store, the fwd+bwd-only rate, the sync time per step, and the
rate the same job would reach with the reference's CPU kvstore('local') arithmetic for the sync
(grad D2H + CommCPU reduce + SGD + H2D, timed on this box's host).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-cpu-arm", action="store_true")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import torchvision
    import anand_mxnet_b200 as mx

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    assert world == args.gpus
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        mx.dist.init_peer_group(local)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    mx.base.set_stream(local, stream.cuda_stream)
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0)

    model = torchvision.models.resnet50().to(dev).to(torch.bfloat16).to(memory_format=torch.channels_last)
    model.train()
    params = [p for p in model.parameters()]
    ctx = mx.gpu(local)
    # re-home parameters and gradients in library arrays; torch keeps zero-copy views of them
    w_nd, g_nd = [], []
    for p in params:
        if p.dim() == 4:
            # conv weights stay channels_last: the library array is the dense (O,H,W,I) buffer and
            # torch sees its (O,I,H,W) permutation; the store's kernels are layout-agnostic
            o, i, h, w = p.shape
            shape, perm = (o, h, w, i), (0, 3, 1, 2)
        else:
            shape, perm = tuple(p.shape), None
        wn = mx.nd.empty(shape, ctx, 'bfloat16')
        gn = mx.nd.empty(shape, ctx, 'bfloat16')
        wt, gt = mx.nd.to_torch(wn), mx.nd.to_torch(gn)
        if perm:
            wt, gt = wt.permute(*perm), gt.permute(*perm)
        wt.copy_(p.data)
        gt.zero_()
        p.data = wt
        p.grad = gt
        w_nd.append(wn)
        g_nd.append(gn)
    n_param = sum(p.numel() for p in params)
    keys = list(range(len(params)))
    kv = mx.kv.create("device")
    kv.init(keys, w_nd)
    kv.set_optimizer(mx.optimizer.SGD(learning_rate=0.1, momentum=0.9, wd=1e-4,
                                      rescale_grad=1.0 / (args.batch * world), multi_precision=True))
    x = torch.randn(args.batch, 3, 224, 224, device=dev, dtype=torch.bfloat16).to(memory_format=torch.channels_last)
    y = torch.randint(0, 1000, (args.batch,), device=dev)
    lossf = torch.nn.CrossEntropyLoss()

    def fwd_bwd():
        for p in params:
            p.grad.zero_()
        loss = lossf(model(x).float(), y)
        loss.backward()
        return loss

    def sync():
        kv.pushpull(keys, g_nd, out=w_nd)

    def timed(fn, n):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(n):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def full_step():
        fwd_bwd()
        sync()

    for _ in range(args.warmup):
        full_step()
    ms_step = timed(full_step, args.steps)
    ms_fb = timed(fwd_bwd, args.steps)
    ms_sync = timed(sync, max(args.steps, 20))
    loss = fwd_bwd().item()

    # ---- the same job with the reference's CPU kvstore('local') doing the sync
    cpu = None
    if rank == 0 and not args.no_cpu_arm:
        from bench import cpu_kvstore_step_fn, best_cpu_threads
        thr = best_cpu_threads("resnet50_sgd", world)
        cstep, kind, cores = cpu_kvstore_step_fn("resnet50_sgd", world, thr)
        cstep()
        t0 = time.perf_counter()
        n = 0
        while n < 2 or (time.perf_counter() - t0 < 8.0 and n < 50):
            cstep()
            n += 1
        cpu_ms = (time.perf_counter() - t0) / n * 1e3
        # PCIe legs of that route: every GPU's gradients to the host, new weights back (fp32 as the
        # reference would hold them), measured with pinned torch tensors on this GPU
        h = torch.empty(n_param, dtype=torch.float32).pin_memory()
        d = torch.empty(n_param, dtype=torch.float32, device=dev)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(5):
            h.copy_(d, non_blocking=True)
            d.copy_(h, non_blocking=True)
        e1.record(stream)
        torch.cuda.synchronize()
        pcie_ms = e0.elapsed_time(e1) / 5
        cpu_sync_ms = cpu_ms + pcie_ms
        cpu = {"kind": kind, "cores": cores, "reduce_update_ms": cpu_ms, "pcie_d2h_h2d_ms": pcie_ms,
               "sync_ms": cpu_sync_ms,
               "samples_per_sec": args.batch * world / ((ms_fb + cpu_sync_ms) * 1e-3)}
    if rank == 0:
        line = {"metric": "resnet50_train_samples_per_sec", "value": args.batch * world / (ms_step * 1e-3),
                "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "dtype": "bf16",
                "data": "synthetic",
                "config": {"model": "torchvision resnet50 (161 parameter tensors, %d elements)" % n_param,
                           "batch_per_gpu": args.batch, "grad_sync": "kvstore('device') pushpull, "
                           "SGD momentum on the store, fp32 master weights, one fused kernel per GPU"},
                "fwd_bwd_only_samples_per_sec": args.batch * world / (ms_fb * 1e-3),
                "fwd_bwd_ms": ms_fb, "sync_ms": ms_sync, "loss": loss,
                "sync_bus_GBps_per_gpu": (n_param * 2 * 2 * (world - 1) / world / (ms_sync * 1e-3) / 1e9)
                if world > 1 else None,
                "cpu_kvstore_local": cpu,
                "speedup_vs_cpu_kvstore_sync": (cpu["sync_ms"] / ms_sync) if cpu else None}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        mx.nd.waitall()
        mx.dist.destroy_peer_group()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
