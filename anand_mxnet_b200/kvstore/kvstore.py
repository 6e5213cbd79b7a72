"""KVStore front-end over the C ABI -- python/mxnet/kvstore/kvstore.py:54-727 mirrored method by
method (each method = one ``_LIB.MXKVStore*`` call wrapped in ``check_call``).

The one behavioural addition: ``set_optimizer`` hands SGD / Adam / Test to the library's fused
kernels (``B200KVStoreSetOptimizer``) instead of installing the per-key Python updater callback;
any other optimizer object -- or ``B200KV_FUSED_OPTIMIZER=0`` -- keeps the reference behaviour
(``_set_updater(get_updater(optimizer))``: the object's own ``update`` runs per key)."""
import ctypes
import os
import pickle

from ..base import _LIB, check_call, c_str, c_str_array, mx_uint, py_str, NDArrayHandle, KVStoreHandle
from ..ndarray import NDArray, _ndarray_cls, array as _nd_array
from .. import optimizer as opt
from .base import _ctype_key_value, _ctype_dict, KVStoreBase

__all__ = ['KVStore']


def _updater_wrapper(updater):
    """A wrapper for the user-defined handle (kvstore.py:34-41)."""
    def updater_handle(key, lhs_handle, rhs_handle, _):
        lhs = _ndarray_cls(NDArrayHandle(lhs_handle))
        rhs = _ndarray_cls(NDArrayHandle(rhs_handle))
        updater(key, lhs, rhs)
    return updater_handle




class KVStore(KVStoreBase):
    """A key-value store for synchronization of values, over multiple devices."""

    def __init__(self, handle):
        assert isinstance(handle, KVStoreHandle)
        self.handle = handle
        self._updater = None
        self._updater_func = None
        self._str_updater_func = None
        self._fused = None       # the Optimizer object whose step runs in the fused kernels
        self._fused_sent = {}    # last scalars handed to the library
        self._mult_sent = {}
        self._keys = []          # every key this front-end initialised, in order

    def __del__(self):
        try:
            check_call(_LIB.MXKVStoreFree(self.handle))
        except Exception:
            pass

    def broadcast(self, key, value, out, priority=0):
        self.init(key, value)
        self.pull(key, out=out, priority=priority)

    @staticmethod
    def is_capable(capability):
        if capability.lower() == KVStoreBase.OPTIMIZER:
            return True
        raise ValueError('Unknown capability: {}'.format(capability))

    def init(self, key, value):
        ckeys, cvals, use_str_keys = _ctype_key_value(key, value)
        for k in (key if isinstance(key, (list, tuple)) else [key]):
            if k not in self._keys:
                self._keys.append(k)
        if use_str_keys:
            check_call(_LIB.MXKVStoreInitEx(self.handle, mx_uint(len(ckeys)), ckeys, cvals))
        else:
            check_call(_LIB.MXKVStoreInit(self.handle, mx_uint(len(ckeys)), ckeys, cvals))

    def push(self, key, value, priority=0):
        ckeys, cvals, use_str_keys = _ctype_key_value(key, value)
        self._sync_fused(key)
        if use_str_keys:
            check_call(_LIB.MXKVStorePushEx(self.handle, mx_uint(len(ckeys)), ckeys, cvals,
                                            ctypes.c_int(priority)))
        else:
            check_call(_LIB.MXKVStorePush(self.handle, mx_uint(len(ckeys)), ckeys, cvals,
                                          ctypes.c_int(priority)))

    def pull(self, key, out=None, priority=0, ignore_sparse=True):
        assert out is not None
        ckeys, cvals, use_str_keys = _ctype_key_value(key, out)
        if use_str_keys:
            check_call(_LIB.MXKVStorePullWithSparseEx(self.handle, mx_uint(len(ckeys)), ckeys, cvals,
                                                      ctypes.c_int(priority),
                                                      ctypes.c_bool(ignore_sparse)))
        else:
            check_call(_LIB.MXKVStorePullWithSparse(self.handle, mx_uint(len(ckeys)), ckeys, cvals,
                                                    ctypes.c_int(priority),
                                                    ctypes.c_bool(ignore_sparse)))

    def pushpull(self, key, value, out=None, priority=0):
        cvkeys, cvals, use_str_keys = _ctype_key_value(key, value)
        if out is not None:
            cokeys, couts, _ = _ctype_key_value(key, out)
        else:
            cokeys, couts = cvkeys, cvals
        self._sync_fused(key)
        if use_str_keys:
            check_call(_LIB.MXKVStorePushPullEx(self.handle, mx_uint(len(cvkeys)), cvkeys,
                                                mx_uint(len(cokeys)), cokeys, cvals, couts,
                                                ctypes.c_int(priority)))
        else:
            check_call(_LIB.MXKVStorePushPull(self.handle, mx_uint(len(cvkeys)), cvkeys,
                                              mx_uint(len(cokeys)), cokeys, cvals, couts,
                                              ctypes.c_int(priority)))

    def row_sparse_pull(self, key, out=None, priority=0, row_ids=None):
        assert out is not None
        assert row_ids is not None
        if isinstance(row_ids, NDArray):
            row_ids = [row_ids]
        assert isinstance(row_ids, list), "row_ids should be NDArray or list of NDArray"
        first_out = out
        single_rowid = False
        if len(row_ids) == 1 and isinstance(out, list):
            single_rowid = True
            first_out = [out[0]]
        ckeys, cvals, use_str_keys = _ctype_key_value(key, first_out)
        _, crow_ids, _ = _ctype_key_value(key, row_ids)
        assert len(crow_ids) == len(cvals), "the number of row_ids doesn't match the number of values"
        if use_str_keys:
            check_call(_LIB.MXKVStorePullRowSparseEx(self.handle, mx_uint(len(ckeys)), ckeys, cvals,
                                                     crow_ids, ctypes.c_int(priority)))
        else:
            check_call(_LIB.MXKVStorePullRowSparse(self.handle, mx_uint(len(ckeys)), ckeys, cvals,
                                                   crow_ids, ctypes.c_int(priority)))
        # the result can be copied to other devices without invoking row_sparse_pull
        # if the indices are the same (kvstore.py:480-485)
        if single_rowid:
            for out_i in out[1:]:
                out[0].copyto(out_i)

    def set_gradient_compression(self, compression_params):
        if ('device' in self.type) or ('dist' in self.type):
            ckeys, cvals = _ctype_dict(compression_params)
            check_call(_LIB.MXKVStoreSetGradientCompression(self.handle,
                                                            mx_uint(len(compression_params)),
                                                            ckeys, cvals))
        else:
            raise Exception('Gradient compression is not supported for this type of kvstore')

    # ------------------------------------------------------------------ optimizer
    def set_optimizer(self, optimizer):
        """kvstore.py:543-590. Single-node stores: install the optimizer as the store's updater --
        natively fused for SGD / Adam / Test (a hyper-parameter record, or an object of the
        reference's classes of those names), through the updater callback otherwise."""
        kind = opt.fused_kind(optimizer)
        if kind is not None and os.environ.get('B200KV_FUSED_OPTIMIZER', '1') != '0':
            self._set_fused(opt.record_of(optimizer))
        else:
            self._fused = None
            self._set_updater(opt.get_updater(optimizer))

    def _set_fused(self, record):
        params = record.op_params()
        keys = list(params.keys())
        vals = [repr(float(v)) if isinstance(v, float) else str(v) for v in params.values()]
        check_call(_LIB.B200KVStoreSetOptimizer(self.handle, c_str(record.kind), mx_uint(len(keys)),
                                                c_str_array(keys), c_str_array(vals)))
        self._fused = record
        self._updater = None
        self._fused_sent = {'lr': record.learning_rate, 'rescale': record.rescale_grad,
                            'rest': self._rest_of(params)}
        self._mult_sent = {}

    @staticmethod
    def _rest_of(params):
        return tuple(sorted((k, v) for k, v in params.items()
                            if k not in ('learning_rate', 'rescale_grad', 'begin_num_update')))

    def _native_key(self, key):
        if isinstance(key, int):
            return key
        k = ctypes.c_int()
        check_call(_LIB.B200KVStoreLookupKey(self.handle, c_str(key), ctypes.byref(k)))
        return k.value

    def _sync_fused(self, key):
        """Hand what may have changed on the Python side since the last call to the library: the
        per-key (lr, wd) multipliers of the keys about to be pushed, the learning rate (scheduler
        included), rescale_grad."""
        o = self._fused
        if o is None:
            return
        src = getattr(o, '_source', None)
        if src is not None:      # an optimizer object of the reference's classes: re-read its fields
            for name in ('lr', 'wd', 'rescale_grad', 'lr_mult', 'wd_mult', 'idx2name', 'param_dict'):
                setattr(o, name, getattr(src, name))
            for name in ('momentum', 'clip_gradient', 'beta1', 'beta2', 'epsilon'):
                if hasattr(src, name):
                    setattr(o, name, getattr(src, name))
        rest = self._rest_of(o.op_params())
        if rest != self._fused_sent['rest']:
            # wd / momentum / clip_gradient / betas changed on the live optimizer: re-configure (the
            # library keeps update counts, multipliers and states of an optimizer of the same kind)
            mult = self._mult_sent
            self._set_fused(o)
            self._mult_sent = mult
        keys = key if isinstance(key, (list, tuple)) else [key]
        changed = []
        for k in keys:
            m = o.multipliers(k)
            if self._mult_sent.get(k) != m:
                self._mult_sent[k] = m
                changed.append((self._native_key(k), m))
        if changed:
            n = len(changed)
            check_call(_LIB.B200KVStoreSetKeyMultipliers(
                self.handle, mx_uint(n), (ctypes.c_int * n)(*[c[0] for c in changed]),
                (ctypes.c_double * n)(*[c[1][0] for c in changed]),
                (ctypes.c_double * n)(*[c[1][1] for c in changed])))
        if o.lr_scheduler is not None:
            # Optimizer._update_count then _get_lr (optimizer.py:412-441): the scheduler sees
            # num_update = max over every index's count, this push included
            lr = o.lr_scheduler(max(o.num_update, self._native_num_update(keys)))
        else:
            lr = o.lr
        if lr != self._fused_sent['lr']:
            check_call(_LIB.B200KVStoreSetLearningRate(self.handle, ctypes.c_double(lr)))
            self._fused_sent['lr'] = lr
        if o.rescale_grad != self._fused_sent['rescale']:
            check_call(_LIB.B200KVStoreSetRescaleGrad(self.handle, ctypes.c_double(o.rescale_grad)))
            self._fused_sent['rescale'] = o.rescale_grad

    def _native_num_update(self, keys):
        """Optimizer.num_update as it stands after the push of `keys` that is about to be issued"""
        c = ctypes.c_int()
        check_call(_LIB.B200KVStoreGetNumUpdate(self.handle, ctypes.byref(c)))
        m = c.value
        for k in keys:
            check_call(_LIB.B200KVStoreGetUpdateCount(self.handle, ctypes.c_int(self._native_key(k)),
                                                      ctypes.byref(c)))
            m = max(m, c.value + 1)
        return m

    @property
    def type(self):
        kv_type = ctypes.c_char_p()
        check_call(_LIB.MXKVStoreGetType(self.handle, ctypes.byref(kv_type)))
        return py_str(kv_type.value)

    @property
    def rank(self):
        rank = ctypes.c_int()
        check_call(_LIB.MXKVStoreGetRank(self.handle, ctypes.byref(rank)))
        return rank.value

    @property
    def num_workers(self):
        size = ctypes.c_int()
        check_call(_LIB.MXKVStoreGetGroupSize(self.handle, ctypes.byref(size)))
        return size.value

    def save_optimizer_states(self, fname, dump_optimizer=False):
        """Same pickle as Updater.get_states (optimizer.py:2155-2161): {index: state}, or
        (states, optimizer) with dump_optimizer. Fused route: the states are fetched from the
        library in the layout the reference's optimizer would have created -- SGD: momentum, or
        (momentum, fp32 master) under multi_precision (optimizer.py:584-594); Adam: (mean, var), or
        (fp32 master, (mean, var)) (optimizer.py:256-268) -- and the optimizer record carries the
        library's update counts, so a resumed Adam keeps its bias-correction step."""
        if self._fused is not None:
            states = self._fused_states()
            if dump_optimizer:
                target = getattr(self._fused, '_source', None) or self._fused
                counts = {k: self._native_count(k) for k in self._keys}
                target._index_update_count = counts
                target.num_update = max([target.num_update] + list(counts.values()))
                payload = pickle.dumps((states, target))
            else:
                payload = pickle.dumps(states)
        else:
            assert self._updater is not None, "Cannot save states for distributed training"
            payload = self._updater.get_states(dump_optimizer)
        with open(fname, 'wb') as fout:
            fout.write(payload)

    def load_optimizer_states(self, fname):
        with open(fname, 'rb') as f:
            data = f.read()
        if self._fused is not None:
            states = pickle.loads(data)
            if isinstance(states, tuple) and len(states) == 2:
                states, saved = states
                if opt.fused_kind(saved) is not None:       # Updater.set_states swaps the optimizer in
                    self._set_fused(opt.record_of(saved))
                counts = getattr(saved, '_index_update_count', None) or {}
                for k, c in counts.items():
                    if k in self._keys:
                        check_call(_LIB.B200KVStoreSetUpdateCount(self.handle,
                                                                  ctypes.c_int(self._native_key(k)),
                                                                  ctypes.c_int(int(c))))
            self._load_fused_states(states)
        else:
            assert self._updater is not None, "Cannot load states for distributed training"
            self._updater.set_states(data)

    def _native_count(self, key):
        c = ctypes.c_int()
        check_call(_LIB.B200KVStoreGetUpdateCount(self.handle, ctypes.c_int(self._native_key(key)),
                                                  ctypes.byref(c)))
        return c.value

    def _fused_states(self):
        kind = self._fused.kind
        out = {}
        for k in self._keys:
            nk = self._native_key(k)

            def get(sid):
                h = NDArrayHandle()
                if _LIB.B200KVStoreGetOptimizerState(self.handle, ctypes.c_int(nk), ctypes.c_int(sid),
                                                     ctypes.byref(h)) != 0:
                    return None
                return NDArray(h)
            w32 = get(2)
            if kind == 'sgd':
                mom = get(0)
                out[k] = (mom, w32) if w32 is not None else mom
            elif kind == 'adam':
                mv = (get(0), get(1))
                if mv[0] is None and w32 is None:
                    continue                                # never pushed: no state yet
                out[k] = (w32, mv) if w32 is not None else mv
        return out

    def _load_fused_states(self, states):
        kind = self._fused.kind
        for k, st in states.items():
            nk = self._native_key(k)
            slots = {}                                      # state_id -> array
            if kind == 'sgd':
                if isinstance(st, (tuple, list)):
                    slots[0], slots[2] = st[0], st[1]
                else:
                    slots[0] = st
            elif kind == 'adam':
                if isinstance(st, (tuple, list)) and len(st) == 2 and isinstance(st[1], (tuple, list)):
                    slots[2], (slots[0], slots[1]) = st[0], st[1]
                elif isinstance(st, (tuple, list)):
                    slots[0], slots[1] = st[0], st[1]
            for sid, arr in slots.items():
                if arr is None:
                    continue
                check_call(_LIB.B200KVStoreSetOptimizerState(self.handle, ctypes.c_int(nk),
                                                             ctypes.c_int(sid), arr.handle))

    def _set_updater(self, updater):
        """kvstore.py:658-696"""
        self._updater = updater
        _updater_proto = ctypes.CFUNCTYPE(None, ctypes.c_int, NDArrayHandle, NDArrayHandle,
                                          ctypes.c_void_p)
        self._updater_func = _updater_proto(_updater_wrapper(updater))
        _str_updater_proto = ctypes.CFUNCTYPE(None, ctypes.c_char_p, NDArrayHandle, NDArrayHandle,
                                              ctypes.c_void_p)
        self._str_updater_func = _str_updater_proto(_updater_wrapper(updater))
        check_call(_LIB.MXKVStoreSetUpdaterEx(self.handle, self._updater_func,
                                              self._str_updater_func, None))

    def _barrier(self):
        check_call(_LIB.MXKVStoreBarrier(self.handle))

    def set_bucket_bytes(self, nbytes):
        """B200 extension: queue per-key push / pull / pushpull calls and fuse them into one launch
        per device every `nbytes` of operands (0 = run each call immediately, the default). Queued
        calls are flushed whenever an array is waited on or read, so results are unchanged."""
        check_call(_LIB.B200KVStoreSetBucketBytes(self.handle, ctypes.c_size_t(int(nbytes))))

    def flush(self):
        check_call(_LIB.B200KVStoreFlush(self.handle))

    def _send_command_to_servers(self, head, body):
        check_call(_LIB.MXKVStoreSendCommmandToServers(self.handle, mx_uint(head), c_str(body)))
