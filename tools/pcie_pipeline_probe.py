#!/usr/bin/env python
"""Measurement harness (not product code): what does an IDEAL copy-engine pipeline over PCIe reach on
this box for the host-buffer (e2e) step -- S bytes of gradients in, a device kernel over them, S
bytes of weights out -- as a function of the bucket size? Pure torch: pinned buffers, three streams
(H2D, compute, D2H) chained by events, back-to-back steps, CUDA-event timed.

    python tools/pcie_pipeline_probe.py [MB=102.2]
"""
import sys
import torch

MB = float(sys.argv[1]) if len(sys.argv) > 1 else 102.2
n = int(MB * 1e6 / 4)
dev = torch.device("cuda", 0)
h_in = torch.empty(n, dtype=torch.float32).pin_memory()
h_out = torch.empty(n, dtype=torch.float32).pin_memory()
d_in = torch.empty(n, device=dev)
d_w = torch.zeros(n, device=dev)
d_out = torch.empty(n, device=dev)
s_in, s_k, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()


def step(bucket_elems, n_copies_per_bucket=1):
    for b0 in range(0, n, bucket_elems):
        b1 = min(n, b0 + bucket_elems)
        with torch.cuda.stream(s_in):
            per = (b1 - b0 + n_copies_per_bucket - 1) // n_copies_per_bucket
            for c0 in range(b0, b1, per):
                d_in[c0:min(b1, c0 + per)].copy_(h_in[c0:min(b1, c0 + per)], non_blocking=True)
            e_in = torch.cuda.Event()
            e_in.record(s_in)
        with torch.cuda.stream(s_k):
            s_k.wait_event(e_in)
            torch.add(d_w[b0:b1], d_in[b0:b1], alpha=-0.1, out=d_out[b0:b1])
            e_k = torch.cuda.Event()
            e_k.record(s_k)
        with torch.cuda.stream(s_out):
            s_out.wait_event(e_k)
            per = (b1 - b0 + n_copies_per_bucket - 1) // n_copies_per_bucket
            for c0 in range(b0, b1, per):
                h_out[c0:min(b1, c0 + per)].copy_(d_out[c0:min(b1, c0 + per)], non_blocking=True)


def timed(bucket_mb, ncopies, steps=10):
    be = max(1, int(bucket_mb * 1e6 / 4))
    for _ in range(3):
        step(be, ncopies)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s_in)
    for _ in range(steps):
        step(be, ncopies)
    torch.cuda.synchronize()
    e1.record(s_out)
    torch.cuda.synchronize()
    # e0 on s_in at start, e1 after everything: elapsed across streams is valid after full sync
    t0 = torch.cuda.Event(enable_timing=True)
    return e0.elapsed_time(e1) / steps


print("S = %.1f MB per direction per step" % (n * 4 / 1e6))
for bucket in (102.2, 51.1, 25.6, 12.8, 8.0, 4.0, 2.0):
    for ncopies in (1, 8):
        ms = timed(bucket, ncopies)
        print("bucket %6.1f MB  copies/bucket %d  %.3f ms/step  %.1f GB/s per direction" % (
            bucket, ncopies, ms, n * 4 / (ms * 1e-3) / 1e9))
