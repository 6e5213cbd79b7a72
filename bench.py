#!/usr/bin/env python
"""bench.py -- KVStore push+pull throughput on the ResNet-50 gradient set (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload ...]

One "step" = one pass of the hot path over the synthetic gradient set: push the gradients of all
157 ResNet-50 tensors (25 549 486 fp32 elements, uniform [-1,1)), reduce them, apply SGD with
momentum ON THE STORE, and pull the new weights back -- one grouped MXKVStorePushPull call, which
the library turns into ONE fused kernel launch per GPU.

Printed JSON (one line, rank 0):
  value      push+pull payload GB/s of the whole job: n_gpus x (bytes pushed + bytes pulled per GPU)
             / step time -- the same definition at every N, so the per-N lines are comparable --
             inputs and outputs resident in HBM, timed with CUDA events on the launching stream
  e2e        same metric through the same C-ABI call with HOST buffers (pinned CPU-context
             NDArrays): the H2D copy of the gradients and the D2H copy of the weights are inside
             the timed region
  roofline   dominant kernel: ALGORITHMIC bytes per launch (24 B/element for
             SGD-momentum at N=1; all-reduce bus bandwidth per GPU, tools/bandwidth/measure.py:137-138,
             at N>=2) / mean launch duration (CUDA events) against the measured HBM copy bandwidth
             (MEASURED_PEAKS.json) or 900 GB/s/dir NVLink
  cpu_baseline   the reference's CPU kvstore('local') arithmetic (oracle/_ref, else the oracle
             port) on this box's host cores, bounded sample
`--impl reference` times that CPU path as the whole job (the driver's reference arm).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "kvstore_push_pull_GBps"
UNIT = "GB/s"
VALUE_FORMULA = ("n_gpus * (bytes pushed + bytes pulled per GPU) / time = n_gpus * 2 * gradient-set "
                 "bytes / time, identical at every N; roofline.achieved uses the kernel's algorithmic "
                 "bytes (HBM, N=1) or the all-reduce bus bandwidth per GPU (NVLink, N>=2)")


# ------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------
def resnet50_shapes():
    """Gradient tensors of the reference's ResNet-50 symbol (example/image-classification/symbols/
    resnet.py, num_layers=50: units [3,4,6,3], filters [64,256,512,1024,2048], bottleneck, 1000
    classes): 157 arrays, 25 549 486 elements."""
    shapes = [(3,), (3,), (64, 3, 7, 7), (64,), (64,)]  # bn_data gamma/beta, conv0, bn0
    filters = [64, 256, 512, 1024, 2048]
    units = [3, 4, 6, 3]
    for stage in range(4):
        nf = filters[stage + 1]
        for u in range(units[stage]):
            cin = filters[stage] if u == 0 else nf
            q = nf // 4
            shapes += [(cin,), (cin,), (q, cin, 1, 1), (q,), (q,), (q, q, 3, 3), (q,), (q,),
                       (nf, q, 1, 1)]
            if u == 0:
                shapes.append((nf, cin, 1, 1))  # projection shortcut
    shapes += [(2048,), (2048,), (1000, 2048), (1000,)]
    return shapes


def bert_base_shapes():
    """BERT-base parameter set: 199 arrays, 109 482 240 elements."""
    H, L, FF, V = 768, 12, 3072, 30522
    shapes = [(V, H), (512, H), (2, H), (H,), (H,)]
    for _ in range(L):
        shapes += [(H, H), (H,)] * 4 + [(H,), (H,), (FF, H), (FF,), (H, FF), (H,), (H,), (H,)]
    shapes += [(H, H), (H,)]
    return shapes


WORKLOADS = {
    "resnet50_sgd": dict(shapes=resnet50_shapes, opt="sgd", bytes_per_elem=24,
                         desc="ResNet-50 gradient set, 157 fp32 tensors / 25549486 elements, "
                              "SGD momentum fused on the store"),
    "bert_adam": dict(shapes=bert_base_shapes, opt="adam", bytes_per_elem=32,
                      desc="BERT-base gradient set, 199 fp32 tensors / 109482240 elements, Adam "
                           "fused on the store"),
}

SGD_KW = dict(learning_rate=0.1, momentum=0.9, wd=1e-4)
ADAM_KW = dict(learning_rate=1e-3, beta1=0.9, beta2=0.999, epsilon=1e-8, wd=0.01)


def payload_bytes(workload, n_gpus):
    """`value`: bytes PUSHED plus bytes PULLED per step over the whole job -- the same definition at
    every N (each GPU pushes its gradient set and pulls the weights), so the per-N values are
    comparable (weak scaling). The roofline uses the kernel's ALGORITHMIC bytes instead."""
    n_elem = sum(int(np.prod(s)) for s in WORKLOADS[workload]["shapes"]())
    return n_gpus * n_elem * 4 * 2


def algorithmic_bytes(workload, n_gpus):
    n_elem = sum(int(np.prod(s)) for s in WORKLOADS[workload]["shapes"]())
    if n_gpus == 1:
        return n_elem * WORKLOADS[workload]["bytes_per_elem"]
    # all-reduce bus bandwidth, per GPU (tools/bandwidth/measure.py:137-138)
    return int(n_elem * 4 * 2 * (n_gpus - 1) / n_gpus)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler(object):
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
                for name, val in zip(names, r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(smax) if smax else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline
# ------------------------------------------------------------------------------------------------
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def best_cpu_threads(workload, n_src):
    """The reference sizes its CPU kernels' OpenMP team from the visible cores
    (engine::OpenMP::GetRecommendedOMPThreadCount); on a big shared host that is far from the
    fastest choice for 157 mostly-small tensors, so the baseline gets the BEST team size from a
    short probe (one step each, ascending, stop once clearly past the optimum)."""
    best, best_t = 1, float("inf")
    cands = sorted({c for c in (1, 2, 4, 8, 16, 32, 64, host_cores()) if c <= host_cores()})
    for c in cands:
        step, _, _ = cpu_kvstore_step_fn(workload, n_src, c)
        step()
        t0 = time.perf_counter()
        step()
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
        elif dt > 2.0 * best_t:
            break
    return best


def cpu_kvstore_step_fn(workload, n_src, threads=None):
    """One kvstore('local') step on host cores: CommCPU reduce of n_src gradient buffers per key
    (comm.h:357-410), then the optimizer kernel per key (one updater call per key, OMP inside, as
    the reference's callback path does), then the copy to n_src outputs (comm.h:209-224)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import kvoracle as K
    ref = K.ref()
    o = K.get_oracle()
    cores = threads or host_cores()
    shapes = WORKLOADS[workload]["shapes"]()
    rng = np.random.default_rng(0xB200)
    sizes = [int(np.prod(s)) for s in shapes]
    W = [rng.uniform(-1, 1, n).astype(np.float32) for n in sizes]
    G = [[rng.uniform(-1, 1, n).astype(np.float32) for _ in range(n_src)] for n in sizes]
    M = [np.zeros(n, np.float32) for n in sizes]
    V = [np.zeros(n, np.float32) for n in sizes]
    OUT = [[np.empty(n, np.float32) for _ in range(n_src)] for n in sizes]
    kind = "reference" if ref is not None else "port"
    sp, f32 = K.scalar_param, K.f32
    opt = WORKLOADS[workload]["opt"]
    state = {"t": 0}

    def step():
        state["t"] += 1
        t = state["t"]
        for k in range(len(sizes)):
            if n_src > 1:
                bufs = [g for g in G[k]]
                if ref is not None:
                    merged_bufs = [bufs[0].copy()] + bufs[1:]      # CopyFromTo(src[0] -> merged)
                    ref.reduce_inplace(merged_bufs, 4)             # MXNET_KVSTORE_REDUCTION_NTHREADS=4
                    merged = merged_bufs[0]
                else:
                    merged = o.reduce(bufs, "local", 4)
            else:
                merged = G[k][0]
            if opt == "sgd":
                if ref is not None:
                    ref.multi_sgd_update([W[k]], [merged], [M[k]], [f32(0.1)], [f32(1e-4)],
                                         sp(0.9), sp(1.0 / 256), None, cores)
                else:
                    o.multi_sgd_update(W[k], merged, M[k], f32(0.1), sp(0.9), f32(1e-4),
                                       sp(1.0 / 256), None, cores)
            else:
                lr_t = sp(K.adam_lr(1e-3, 0.9, 0.999, t))
                fn = ref.adam_update if ref is not None else o.adam_update
                fn(W[k], merged, M[k], V[k], lr_t, sp(0.9), sp(0.999), sp(1e-8), sp(0.01), 1.0,
                   None, cores)
            for out in OUT[k]:
                np.copyto(out, W[k])
    return step, kind, cores


def run_reference(args):
    """The reference arm: the reference's own CPU implementation of the path on this box's cores."""
    n_src = max(1, args.gpus)
    step, kind, cores = cpu_kvstore_step_fn(args.workload, n_src, best_cpu_threads(args.workload, n_src))
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    value = payload_bytes(args.workload, n_src) / dt / 1e9
    sample = "%d full steps of %s with %d host gradient buffers per key" % (
        args.steps, args.workload, n_src)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload]["desc"], "store": "kvstore('local') on CPU",
                       "values_per_key": n_src, "value_formula": VALUE_FORMULA},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": sample, "host_cores_visible": host_cores()},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# the B200 arm
# ------------------------------------------------------------------------------------------------
def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def make_optimizer(mx, workload, n_gpus):
    if WORKLOADS[workload]["opt"] == "sgd":
        return mx.optimizer.SGD(rescale_grad=1.0 / (256 * n_gpus), **SGD_KW)
    return mx.optimizer.Adam(rescale_grad=1.0, **ADAM_KW)


def run_single_gpu(args):
    import torch
    import anand_mxnet_b200 as mx
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    # a dedicated non-default stream shared by torch (events, data generation) and the library
    # (B200KVEngineSetStream): the CUDA events below are recorded on the stream the kernels run on
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    mx.base.set_stream(0, stream.cuda_stream)
    assert mx.base.get_stream(0) == stream.cuda_stream
    shapes = WORKLOADS[args.workload]["shapes"]()
    keys = list(range(len(shapes)))
    n_elem = sum(int(np.prod(s)) for s in shapes)
    gen = torch.Generator(device=dev)
    gen.manual_seed(0xB200)
    sizes = [int(np.prod(s)) for s in shapes]
    offs = np.concatenate([[0], np.cumsum([(n + 127) // 128 * 128 for n in sizes])]).astype(np.int64)

    def flat_views(fill):
        """one flat buffer per role (three torch kernels in total), one 512-byte aligned view per
        tensor -- the arrays stay separate NDArrays, exactly what a framework hands the store"""
        flat = torch.empty(int(offs[-1]), device=dev)
        if fill:
            flat.uniform_(-1, 1, generator=gen)
        return flat, [flat[int(offs[i]):int(offs[i]) + sizes[i]].view(shapes[i]) for i in range(len(shapes))]

    # ---- device-resident arm
    kv = mx.kv.create("device")
    _w_flat, w_views = flat_views(True)
    kv.init(keys, [mx.nd.from_torch(t) for t in w_views])
    kv.set_optimizer(make_optimizer(mx, args.workload, 1))
    _g_flat, grads_t = flat_views(True)
    _o_flat, outs_t = flat_views(False)
    grads = [mx.nd.from_torch(t) for t in grads_t]
    outs = [mx.nd.from_torch(t) for t in outs_t]
    torch.cuda.synchronize()
    kv.pushpull(keys, grads, out=outs)      # python front-end once: hands lr / multipliers over
    # the timed call is the C-ABI entry point itself with prebuilt argument arrays, as a compiled
    # host would issue it; the python front-end's per-call marshalling is reported separately
    from anand_mxnet_b200.kvstore.base import _ctype_key_value
    ckeys, cvals, _ = _ctype_key_value(keys, grads)
    _, couts, _ = _ctype_key_value(keys, outs)
    lib, handle, nkeys = mx.base._LIB, kv.handle, ctypes.c_uint(len(keys))
    zero = ctypes.c_int(0)

    def step():
        rc = lib.MXKVStorePushPull(handle, nkeys, ckeys, nkeys, ckeys, cvals, couts, zero)
        if rc != 0:
            raise RuntimeError(lib.MXGetLastError().decode())

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    mx.base.reset_kernel_launch_count()
    t_all0, t_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t_all0.record(stream)
    for i in range(args.steps):
        step()
    t_all1.record(stream)
    torch.cuda.synchronize()
    launches = mx.base.kernel_launch_count()
    clocks = sampler.stop()
    ms_total = t_all0.elapsed_time(t_all1)
    ms_step = ms_total / args.steps
    # the timed region is exactly K launches of the dominant kernel back to back on this stream:
    # its average launch duration (incl. launch gaps) is the region's duration / K
    ms_kernel = ms_step
    alg = algorithmic_bytes(args.workload, 1)
    pay = payload_bytes(args.workload, 1)
    value = pay / (ms_step * 1e-3) / 1e9
    peaks, peak_src = measured_peaks()
    achieved = alg / (ms_kernel * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": _ncu_traffic(args.workload), "kernel":
                "dense_fused_kernel<float,1,SGD>" if WORKLOADS[args.workload]["opt"] == "sgd"
                else "dense_fused_kernel<float,1,Adam>", "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg, "ms_per_launch": ms_kernel}

    # ---- the same loop through the python front-end (kv.pushpull): includes ctypes marshalling
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    th = time.perf_counter()
    p0.record(stream)
    for _ in range(args.steps):
        kv.pushpull(keys, grads, out=outs)
    p1.record(stream)
    host_py_us = (time.perf_counter() - th) / args.steps * 1e6
    torch.cuda.synchronize()
    py_ms = p0.elapsed_time(p1) / args.steps
    th = time.perf_counter()
    for _ in range(args.steps):
        step()
    host_c_us = (time.perf_counter() - th) / args.steps * 1e6
    torch.cuda.synchronize()
    frontends = {"python_api_GBps": pay / (py_ms * 1e-3) / 1e9, "python_api_ms_per_step": py_ms,
                 "python_api_host_us_per_call": host_py_us, "c_abi_host_us_per_call": host_c_us}

    # ---- end-to-end arm: same C-ABI call, HOST (pinned) gradient and weight buffers
    kv2 = mx.kv.create("device")
    kv2.init(keys, [mx.nd.array(np.zeros(s, np.float32), mx.cpu()) for s in shapes])
    kv2.set_optimizer(make_optimizer(mx, args.workload, 1))
    rng = np.random.default_rng(0xB200)
    hgrads = [mx.nd.array(rng.uniform(-1, 1, s).astype(np.float32), mx.cpu()) for s in shapes]
    houts = [mx.nd.empty(s, mx.cpu()) for s in shapes]
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        kv2.pushpull(keys, hgrads, out=houts)
    mx.nd.waitall()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(e2e_steps):
        kv2.pushpull(keys, hgrads, out=houts)
    mx.nd.waitall()        # the last D2H copies run on the copy-out lane: wait before stamping
    e1.record(stream)
    torch.cuda.synchronize()
    e2e_ms = e0.elapsed_time(e1) / e2e_steps
    e2e = {"value": pay / (e2e_ms * 1e-3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": n_elem * 4,
           "d2h_bytes_per_step": n_elem * 4, "ms_per_step": e2e_ms, "steps": e2e_steps}

    # ---- cpu baseline (bounded sample, rank 0, N=1)
    cpu = None
    if not args.no_cpu_baseline:
        cstep, kind, cores = cpu_kvstore_step_fn(args.workload, 1, best_cpu_threads(args.workload, 1))
        cstep()
        t0 = time.perf_counter()
        n = 0
        while n < 3 or (time.perf_counter() - t0 < 10.0 and n < 200):
            cstep()
            n += 1
        cdt = (time.perf_counter() - t0) / n
        cpu = {"value": pay / cdt / 1e9, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": "%d full steps of %s (reduce of 1 value + optimizer + copy-out per key), "
                         "%.1f ms/step" % (n, args.workload, cdt * 1e3)}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload]["desc"], "store": "kvstore('device')",
                       "call": "one grouped MXKVStorePushPull (C ABI, prebuilt argument arrays) over "
                               "all keys per step; python front-end timing under 'frontends'",
                       "l2": "working set %.0f MB per step > 126 MB L2, no flush needed" % (alg / 1e6),
                       "value_formula": VALUE_FORMULA,
                       "optimizer": SGD_KW if WORKLOADS[args.workload]["opt"] == "sgd" else ADAM_KW},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks, "frontends": frontends}
    print(json.dumps(line))


def _ncu_traffic(workload):
    """dram bytes per launch of the dominant kernel from the committed ncu capture of THIS workload
    (profiles/r01_dense_fused_traffic.json holds the ResNet-50 SGD-momentum launch), else None."""
    if workload != "resnet50_sgd":
        return None
    p = os.path.join(ROOT, "profiles", "r01_dense_fused_traffic.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                return json.load(f).get("dram_bytes_per_launch")
        except Exception:
            return None
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="resnet50_sgd", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        if rank == 0:
            run_reference(args)
        return
    if args.gpus == 1 and int(os.environ.get("WORLD_SIZE", "1")) == 1:
        run_single_gpu(args)
        return
    from bench_multi import run_multi_gpu
    run_multi_gpu(args)


if __name__ == "__main__":
    main()
