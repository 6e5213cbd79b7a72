"""CPU-only test of the N>1 host logic: world_size-2 `gloo` job on 127.0.0.1 exercising the
launcher-supplied all-gather callback and the rank-major operand-offset exchange that plan building
uses in one-rank-per-GPU mode (csrc/group.cc GatherI64), including its mismatch detection. The
CUDA-IPC mapping itself needs GPUs and is covered by tests/test_group_gpu.py."""
import ctypes
import os
import socket

import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import anand_mxnet_b200 as mx
        cb = mx.dist.make_allgather_callback(None)
        lib = mx.base._LIB
        n = 6
        mine = (ctypes.c_int64 * n)(*[1000 * rank + i for i in range(n)])
        out = (ctypes.c_int64 * (n * world))()
        rc = lib.B200KVTestGatherI64(world, cb, None, mine, n, out)
        ok = rc == 0 and list(out) == [1000 * r + i for r in range(world) for i in range(n)]
        # ranks that disagree on the call (different lengths) must get an error, not garbage
        n_bad = n if rank == 0 else n - 1
        mine2 = (ctypes.c_int64 * n_bad)(*range(n_bad))
        rc2 = lib.B200KVTestGatherI64(world, cb, None, mine2, n_bad, out)
        msg = lib.MXGetLastError().decode()
        ok = ok and rc2 == -1 and "different KVStore calls" in msg
        q.put((rank, ok, msg if not ok else ""))
    finally:
        dist.destroy_process_group()


def test_offset_exchange_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, msg in results:
        assert ok, (rank, msg)


def _mailbox_worker(rank, world, port, q):
    import random
    import time
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import anand_mxnet_b200 as mx
        cb = mx.dist.make_allgather_callback(None)
        lib = mx.base._LIB
        ok, msg = True, ""
        # n = 6: one slot; n = 1000: 8000 bytes = two 4 KB pieces per gather; many rounds so that a
        # fast rank runs ahead of a slow one (the two-slot protocol must hold it back)
        for n, rounds in ((6, 300), (1000, 120)):
            mine = (ctypes.c_int64 * n)(*[100000 * rank + i for i in range(n)])
            out = (ctypes.c_int64 * (n * world))()
            used = ctypes.c_int(0)
            if rank == world - 1:
                time.sleep(0.05 * random.random())
            rc = lib.B200KVTestMailbox(rank, world, cb, None, mine, n, rounds, out, ctypes.byref(used))
            want = [100000 * r + i + rounds for r in range(world) for i in range(n)]
            if rc != 0 or list(out) != want:
                ok, msg = False, "n=%d rc=%d %s" % (n, rc, lib.MXGetLastError().decode())
            if used.value != 1:
                ok, msg = False, "shared memory mailbox not used (n=%d)" % n
        q.put((rank, ok, msg))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_shared_memory_mailbox_allgather(world):
    """the host-side metadata exchange of the rank-per-GPU store (group.cc: shared-memory mailbox
    built through ONE exchange over the launcher's callback, then spin-wait all-gathers)"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_mailbox_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, msg in results:
        assert ok, (rank, msg)
