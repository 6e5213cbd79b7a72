"""mx.kv / mx.kvstore namespace (python/mxnet/kvstore/__init__.py)."""
from .base import KVStoreBase, TestStore, create
from .kvstore import KVStore
from .allreduce import AllReduceStore

__all__ = ['KVStoreBase', 'KVStore', 'TestStore', 'AllReduceStore', 'create']
