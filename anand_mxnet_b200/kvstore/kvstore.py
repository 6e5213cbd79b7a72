"""KVStore front-end over the C ABI -- python/mxnet/kvstore/kvstore.py:54-727 mirrored method by
method (each method = one ``_LIB.MXKVStore*`` call wrapped in ``check_call``).

The one behavioural addition: ``set_optimizer`` hands SGD / Adam / Test to the library's fused
kernels (``B200KVStoreSetOptimizer``) instead of installing the per-key Python updater callback;
any other optimizer -- or ``B200KV_FUSED_OPTIMIZER=0`` -- keeps the reference behaviour
(``_set_updater(opt.get_updater(optimizer))``)."""
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


_FUSED_KINDS = {'SGD': 'sgd', 'Adam': 'adam', 'Test': 'test'}


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
        self._mult_sent = set()

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
        natively fused for SGD / Adam / Test, through the updater callback otherwise."""
        kind = _FUSED_KINDS.get(type(optimizer).__name__)
        fused_ok = kind is not None and os.environ.get('B200KV_FUSED_OPTIMIZER', '1') != '0' \
            and type(optimizer).__module__ == opt.__name__
        if fused_ok:
            self._set_fused(optimizer, kind)
        else:
            self._fused = None
            self._set_updater(opt.get_updater(optimizer))

    def _set_fused(self, optimizer, kind):
        o = optimizer
        params = {'learning_rate': o.learning_rate, 'wd': o.wd, 'rescale_grad': o.rescale_grad,
                  'clip_gradient': o.clip_gradient if o.clip_gradient else 0.0,
                  'multi_precision': bool(o.multi_precision),
                  'begin_num_update': o.begin_num_update}
        if kind == 'sgd':
            params['momentum'] = o.momentum
            params['lazy_update'] = bool(o.lazy_update)
        elif kind == 'adam':
            params.update(beta1=o.beta1, beta2=o.beta2, epsilon=o.epsilon,
                          lazy_update=bool(o.lazy_update))
        keys = list(params.keys())
        vals = [repr(float(v)) if isinstance(v, float) else str(v) for v in params.values()]
        check_call(_LIB.B200KVStoreSetOptimizer(self.handle, c_str(kind), mx_uint(len(keys)),
                                                c_str_array(keys), c_str_array(vals)))
        self._fused = o
        self._updater = None
        self._fused_sent = {'lr': o.learning_rate, 'rescale': o.rescale_grad}
        self._mult_sent = set()

    def _native_key(self, key):
        if isinstance(key, int):
            return key
        k = ctypes.c_int()
        check_call(_LIB.B200KVStoreLookupKey(self.handle, c_str(key), ctypes.byref(k)))
        return k.value

    def _sync_fused(self, key):
        """Hand the per-step scalars and the per-key multipliers of the Python Optimizer object
        (lr / lr_scheduler, rescale_grad, lr_mult, wd_mult) to the library when they changed."""
        o = self._fused
        if o is None:
            return
        keys = key if isinstance(key, (list, tuple)) else [key]
        new = [k for k in keys if k not in self._mult_sent]
        if new:
            nk = [self._native_key(k) for k in new]
            lrm = o._get_lrs(new)
            wdm = o._get_wds(new)
            base_lr = o.learning_rate
            lr_mult = [(l / base_lr) if base_lr != 0 else 1.0 for l in lrm]
            wd_mult = [(w / o.wd) if o.wd != 0 else 1.0 for w in wdm]
            # exact multipliers (not ratios) when they are directly available
            for i, k in enumerate(new):
                lm, wm = _exact_mults(o, k)
                if lm is not None:
                    lr_mult[i] = lm
                if wm is not None:
                    wd_mult[i] = wm
            check_call(_LIB.B200KVStoreSetKeyMultipliers(
                self.handle, mx_uint(len(nk)), (ctypes.c_int * len(nk))(*nk),
                (ctypes.c_double * len(nk))(*lr_mult), (ctypes.c_double * len(nk))(*wd_mult)))
            self._mult_sent.update(new)
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
        """Same pickle layout as Updater.get_states (optimizer.py:2155-2161): {index: state}."""
        if self._fused is not None:
            states = self._fused_states()
            payload = pickle.dumps((states, self._fused) if dump_optimizer else states)
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
                states = states[0]
            self._load_fused_states(states)
        else:
            assert self._updater is not None, "Cannot load states for distributed training"
            self._updater.set_states(data)

    def _fused_states(self):
        kind = _FUSED_KINDS[type(self._fused).__name__]
        out = {}
        for k in sorted(self._mult_sent, key=str):
            nk = self._native_key(k)

            def get(sid):
                h = NDArrayHandle()
                if _LIB.B200KVStoreGetOptimizerState(self.handle, ctypes.c_int(nk), ctypes.c_int(sid),
                                                     ctypes.byref(h)) != 0:
                    return None
                return NDArray(h)
            if kind == 'sgd':
                out[k] = get(0)
            elif kind == 'adam':
                out[k] = (get(0), get(1))
        return out

    def _load_fused_states(self, states):
        for k, st in states.items():
            nk = self._native_key(k)
            parts = st if isinstance(st, (tuple, list)) else (st,)
            for sid, s in enumerate(parts):
                if s is None:
                    continue
                check_call(_LIB.B200KVStoreSetOptimizerState(self.handle, ctypes.c_int(nk),
                                                             ctypes.c_int(sid), s.handle))

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


def _exact_mults(o, index):
    """Optimizer._get_lrs / _get_wds lookup order (optimizer.py:432-509), returning the multipliers
    themselves so the library multiplies exactly as the reference does (lr * mult in double)."""
    lm = wm = None
    if index in o.param_dict:
        lm = o.param_dict[index].lr_mult
        wm = o.param_dict[index].wd_mult
        return lm, wm
    if index in o.lr_mult:
        lm = o.lr_mult[index]
    elif index in o.idx2name:
        lm = o.lr_mult.get(o.idx2name[index], 1.0)
    else:
        lm = 1.0
    if index in o.wd_mult:
        wm = o.wd_mult[index]
    elif index in o.idx2name:
        wm = o.wd_mult.get(o.idx2name[index], 1.0)
    else:
        wm = 1.0
    return lm, wm
