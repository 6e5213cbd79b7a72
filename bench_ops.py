#!/usr/bin/env python
"""bench_ops.py -- the update_on_kvstore=False optimizer operators on one GPU:

  lars        ResNet-50 set (157 tensors / 25.5 M fp32): multi_sum_sq(w) + multi_sum_sq(g) over the
              LARS layers, multi_lars, then preloaded_multi_sgd_mom_update over everything -- the
              exact operator sequence of mx.optimizer.LARS (optimizer.py:934-1030) with ONE
              aggregated call instead of 4-tensor chunks
  lamb        BERT-base set (199 tensors / 109.5 M fp32): _multi_lamb_update in 45-tensor calls
  adamw       BERT-base set: _multi_adamw_update in 50-tensor calls

Prints one JSON line per workload: ms per optimizer step (CUDA events on the launching stream,
inputs resident in HBM, working sets > L2), algorithmic GB/s against the measured HBM peak, kernel
launches per step, and the oracle (CPU port of the reference kernels, single thread) on a bounded
sample of the same tensors.

Algorithmic bytes per element: lars 4+4 (norm passes over the LARS layers) + 20 (w,g,m read; w,m
written); lamb 16+12 (step 1: w,g,m,v read; m,v,temp written) + 8+4 (step 2); adamw 16+12.
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


def flat_views(mx, torch, shapes, seed, lo=-1.0, hi=1.0):
    n = sum(int(np.prod(s)) for s in shapes)
    pad = sum((-int(np.prod(s))) % 128 for s in shapes)
    buf = torch.empty(n + pad, device='cuda', dtype=torch.float32)
    g = torch.Generator(device='cuda').manual_seed(seed)
    buf.uniform_(lo, hi, generator=g)
    out, off = [], 0
    for s in shapes:
        k = int(np.prod(s))
        out.append(mx.nd.from_torch(buf[off:off + k].view(*s)))
        off += k + ((-k) % 128)
    return out, buf


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    import torch
    import anand_mxnet_b200 as mx
    import kvoracle as K
    from bench import resnet50_shapes, bert_base_shapes as bert_shapes, measured_peaks
    peaks, peak_src = measured_peaks()
    hbm = float(peaks.get("hbm_gbs", 6572.2))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    mx.base.set_stream(0, stream.cuda_stream)
    o = K.get_oracle()

    def timed(step):
        for _ in range(args.warmup):
            step()
        torch.cuda.synchronize()
        mx.base.reset_kernel_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        host_ms = (time.perf_counter() - t0) * 1e3 / args.steps   # time to ISSUE a step (no sync)
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.steps, mx.base.kernel_launch_count() / args.steps, host_ms

    def emit(name, ms, launches, nbytes, n_elem, cpu_ms, cpu_elems, extra, host_ms=None):
        gbs = nbytes / (ms * 1e-3) / 1e9
        print(json.dumps({
            "metric": "optimizer_step_GBps", "workload": name, "value": gbs, "unit": "GB/s",
            "ms_per_step": ms, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "elements": n_elem, "algorithmic_bytes": nbytes, "gpu_launches_per_step": launches, "host_issue_ms_per_step": host_ms,
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s",
                         "frac": gbs / hbm, "peak_source": peak_src},
            "cpu_baseline": {"kind": "port", "cores": 1, "ms_per_step_extrapolated": cpu_ms * n_elem / max(cpu_elems, 1),
                             "sample": "%d of %d elements, single thread" % (cpu_elems, n_elem)},
            "data": "synthetic", **extra}))

    # ------------------------------------------------------------------------------------ LARS
    if args.only in ("", "lars"):
        shapes = resnet50_shapes()
        n = len(shapes)
        n_elem = sum(int(np.prod(s)) for s in shapes)
        ws, _b1 = flat_views(mx, torch, shapes, 1)
        gs, _b2 = flat_views(mx, torch, shapes, 2)
        ms_, _b3 = flat_views(mx, torch, shapes, 3, 0.0, 0.0)
        lars_idx = [i for i, s in enumerate(shapes) if len(s) > 1]       # conv / fc weights
        rest = [i for i in range(n) if len(shapes[i]) == 1]              # gamma / beta / bias
        order = lars_idx + rest
        nb = len(lars_idx)
        lars_elem = sum(int(np.prod(shapes[i])) for i in lars_idx)
        lrs0 = mx.nd.array(np.full(n, 0.1, np.float32), mx.gpu(0))
        lrs = mx.nd.array(np.full(n, 0.1, np.float32), mx.gpu(0))
        wds = mx.nd.array(np.array([1e-4] * nb + [0.0] * (n - nb), np.float32), mx.gpu(0))
        w_o, g_o, m_o = [ws[i] for i in order], [gs[i] for i in order], [ms_[i] for i in order]
        flat = [x for t in zip(w_o, g_o, m_o) for x in t]
        wsq = mx.nd.zeros((nb,), mx.gpu(0))
        gsq = mx.nd.zeros((nb,), mx.gpu(0))

        def step():
            mx.nd.multi_sum_sq(*w_o[:nb], num_arrays=nb, out=wsq)
            mx.nd.multi_sum_sq(*g_o[:nb], num_arrays=nb, out=gsq)
            mx.nd.multi_lars(lrs0[:nb], wsq, gsq, wds[:nb], eta=0.001, eps=0, rescale_grad=1.0 / 256,
                             out=lrs[:nb])
            mx.nd.preloaded_multi_sgd_mom_update(*(flat + [lrs, wds]), out=w_o, num_weights=n,
                                                 rescale_grad=1.0 / 256, momentum=0.9)
        ms, launches, host_ms = timed(step)
        # CPU port on the largest tensor
        big = max(range(n), key=lambda i: int(np.prod(shapes[i])))
        w, g, m = (x.asnumpy().ravel() for x in (ws[big], gs[big], ms_[big]))
        t0 = time.perf_counter()
        o.multi_sum_sq([w])
        o.multi_sum_sq([g])
        o.multi_sgd_update(w, g, m, 0.1, momentum=0.9, wd=1e-4, rescale=1.0 / 256)
        cpu_ms = (time.perf_counter() - t0) * 1e3
        emit("lars_resnet50", ms, launches, lars_elem * 8 + n_elem * 20, n_elem, cpu_ms, w.size,
             {"config": {"workload": "ResNet-50 set, LARS (multi_sum_sq x2, multi_lars, "
                                     "preloaded_multi_sgd_mom_update), %d LARS layers" % nb}}, host_ms)
        del ws, gs, ms_, flat, w_o, g_o, m_o, _b1, _b2, _b3

    # ----------------------------------------------------------------------------- LAMB / AdamW
    for name in ("lamb", "adamw"):
        if args.only not in ("", name):
            continue
        shapes = bert_shapes()
        n = len(shapes)
        n_elem = sum(int(np.prod(s)) for s in shapes)
        ws, _b1 = flat_views(mx, torch, shapes, 11)
        gs, _b2 = flat_views(mx, torch, shapes, 12)
        means, _b3 = flat_views(mx, torch, shapes, 13, 0.0, 0.0)
        vars_, _b4 = flat_views(mx, torch, shapes, 14, 0.0, 0.0)
        per_call = 45 if name == "lamb" else 50
        rs = mx.nd.array(np.array([1.0 / 256], np.float32), mx.gpu(0))
        state = {"t": 0}

        def step():
            state["t"] += 1
            for s0 in range(0, n, per_call):
                sl = slice(s0, min(s0 + per_call, n))
                k = sl.stop - sl.start
                if name == "lamb":
                    mx.nd.contrib.multi_lamb_update(ws[sl], gs[sl], means[sl], vars_[sl], out=ws[sl],
                                                    step_count=[state["t"]] * k, lrs=[1e-3] * k,
                                                    wds=[0.01] * k, beta1=0.9, beta2=0.999,
                                                    epsilon=1e-6, rescale_grad=1.0 / 256)
                else:
                    mx.nd.contrib.multi_adamw_update(ws[sl], gs[sl], means[sl], vars_[sl], rs,
                                                     out=ws[sl], lrs=[1e-3] * k, wds=[0.01] * k,
                                                     etas=[1.0] * k, beta1=0.9, beta2=0.999,
                                                     epsilon=1e-6)
        ms, launches, host_ms = timed(step)
        big = 1 if int(np.prod(shapes[1])) < 4_000_000 else 1
        big = min(range(n), key=lambda i: abs(int(np.prod(shapes[i])) - 2359296))
        w, g, m, v = (x.asnumpy().ravel() for x in (ws[big], gs[big], means[big], vars_[big]))
        t0 = time.perf_counter()
        if name == "lamb":
            o.multi_lamb_update([w], [g], [m], [v], [1], [1e-3], [0.01], rescale=1.0 / 256)
        else:
            o.multi_adamw_update([w], [g], [m], [v], 1.0 / 256, [1e-3], [0.01], [1.0])
        cpu_ms = (time.perf_counter() - t0) * 1e3
        per_elem = 40 if name == "lamb" else 28
        emit("%s_bert_base" % name, ms, launches, n_elem * per_elem, n_elem, cpu_ms, w.size,
             {"config": {"workload": "BERT-base set, %s in %d-tensor calls" %
                                     ("_multi_lamb_update" if name == "lamb" else "_multi_adamw_update",
                                      per_call)}}, host_ms)
        del ws, gs, means, vars_, _b1, _b2, _b3, _b4


if __name__ == "__main__":
    main()
