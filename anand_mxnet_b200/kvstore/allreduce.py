"""NCCL fallback store for one-rank-per-GPU jobs whose GPUs cannot map each other's memory (no CUDA
IPC / no peer access): the cross-rank sum is a torch.distributed all-reduce (NCCL over whatever
links exist) on ONE packed fusion buffer per dtype, then the local library store applies the fused
optimizer step to the reduced gradient and writes the weights -- the role KVStoreNCCL
(src/kvstore/kvstore_nccl.h:62-551) plays in the reference, with a flat bucket instead of grouped
per-key ncclReduce/ncclBcast calls.

This is the FALLBACK: the summation order is NCCL's (not the reference's), so results match the
oracle to ~1e-6 relative rather than bit-for-bit, and the transfer is not fused with the update.
The primary N>1 path is the peer-memory kernel (mx.dist.init_peer_group + kv.create('device')).

    kv = mx.kv.create('allreducestore')      # after torch.distributed.init_process_group
"""
import numpy as np

from .base import KVStoreBase
from . import base as _base
from ..ndarray import NDArray, from_torch, to_torch


@KVStoreBase.register
class AllReduceStore(KVStoreBase):
    """kv.create('allreducestore')"""

    def __init__(self):
        import torch.distributed as dist
        assert dist.is_initialized(), "AllReduceStore needs an initialised torch.distributed job"
        self._dist = dist
        self._local = _base.create('device')   # the native single-GPU store does the update
        self._flat = {}                         # (dtype, total) -> flat torch buffer
        self._views = {}

    # ---- KVStoreBase interface
    def broadcast(self, key, value, out, priority=0):
        self.init(key, value)
        self.pull(key, out=out, priority=priority)

    def init(self, key, value):
        self._local.init(key, value)

    def set_optimizer(self, optimizer):
        self._local.set_optimizer(optimizer)

    @staticmethod
    def is_capable(capability):
        if capability.lower() == KVStoreBase.OPTIMIZER:
            return True
        raise ValueError('Unknown capability: {}'.format(capability))

    @property
    def type(self):
        return 'dist_device_allreduce'

    @property
    def rank(self):
        return self._dist.get_rank()

    @property
    def num_workers(self):
        return self._dist.get_world_size()

    def save_optimizer_states(self, fname, dump_optimizer=False):
        self._local.save_optimizer_states(fname, dump_optimizer)

    def load_optimizer_states(self, fname):
        self._local.load_optimizer_states(fname)

    def pull(self, key, out=None, priority=0, ignore_sparse=True):
        self._local.pull(key, out=out, priority=priority, ignore_sparse=ignore_sparse)

    def push(self, key, value, priority=0):
        self._local.push(key, self._allreduce(key, value), priority=priority)

    def pushpull(self, key, value, out=None, priority=0):
        reduced = self._allreduce(key, value)
        self._local.pushpull(key, reduced, out=out if out is not None else value, priority=priority)

    # ---- the collective
    def _allreduce(self, key, value):
        import torch
        keys = key if isinstance(key, (list, tuple)) else [key]
        vals = value if isinstance(key, (list, tuple)) else [value]
        for v in vals:
            assert isinstance(v, NDArray), "AllReduceStore: one value per key per rank"
        ts = [to_torch(v) for v in vals]
        sig = (tuple(keys), tuple(t.data_ptr() for t in ts))
        if sig not in self._views:
            total = sum(t.numel() for t in ts)
            flat = torch.empty(total, dtype=ts[0].dtype, device=ts[0].device)
            views, off = [], 0
            for t in ts:
                views.append(flat[off:off + t.numel()].view(t.shape))
                off += t.numel()
            self._views[sig] = (flat, views, [from_torch(v) for v in views])
        flat, views, nds = self._views[sig]
        torch._foreach_copy_(views, ts)                     # pack
        self._dist.all_reduce(flat, op=self._dist.ReduceOp.SUM)
        torch.cuda.current_stream().synchronize()           # the library runs on its own stream
        return nds if isinstance(key, (list, tuple)) else nds[0]
