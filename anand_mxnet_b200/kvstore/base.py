"""KVStoreBase, the plugin registry and ``create`` -- python/mxnet/kvstore/base.py:33-455 mirrored:
``create(name)`` consults the registry first (``KVStoreBase.register``) and otherwise asks the native
library (``MXKVStoreCreate``)."""
import ctypes
import warnings
from array import array as _pyarray
from itertools import chain, repeat
from operator import attrgetter

_HV = attrgetter('_hv')

from ..base import _LIB, check_call, c_str, c_str_array, c_array, c_handle_array, string_types, \
    KVStoreHandle
from ..ndarray import NDArray


def _ctype_key_value(keys, vals):
    """python/mxnet/kvstore/base.py:33-65: flatten (nested) key/value lists into C arrays."""
    if isinstance(keys, (tuple, list)):
        assert len(keys) == len(vals)
        # one pass, no per-key helper calls: a 157-key call is marshalled in ~25 us
        kinds = {type(k) for k in keys}
        assert len(kinds) <= 1, "inconsistent types of keys detected."
        use_str_keys = bool(kinds) and issubclass(next(iter(kinds)), string_types)
        assert not kinds or use_str_keys or issubclass(next(iter(kinds)), int), \
            "unexpected type for keys: " + str(kinds)
        try:
            if vals and isinstance(vals[0], NDArray):          # one value per key
                c_keys = keys
                hbuf = _pyarray('Q', map(_HV, vals))
            else:                                              # a list of values per key
                lens = list(map(len, vals))
                c_keys = list(chain.from_iterable(map(repeat, keys, lens)))
                hbuf = _pyarray('Q', map(_HV, chain.from_iterable(vals)))
        except (AttributeError, TypeError):                    # mixed nesting: general walk
            c_keys, flat = [], []
            for key, val in zip(keys, vals):
                if isinstance(val, NDArray):
                    c_keys.append(key)
                    flat.append(val)
                else:
                    for v in val:
                        assert isinstance(v, NDArray)
                    c_keys.extend([key] * len(val))
                    flat.extend(val)
            hbuf = _pyarray('Q', [v._hv for v in flat])
        handles = (ctypes.c_void_p * len(hbuf)).from_buffer(hbuf)
        if use_str_keys:
            c_keys_arr = c_str_array(c_keys)
        else:
            kbuf = _pyarray('i', c_keys)
            c_keys_arr = (ctypes.c_int * len(kbuf)).from_buffer(kbuf)
        return c_keys_arr, handles, use_str_keys
    c_keys, c_vals, use_str_keys = _flat_key_value(keys, vals)
    c_keys_arr = c_str_array(c_keys) if use_str_keys else c_array(ctypes.c_int, c_keys)
    return c_keys_arr, c_handle_array(c_vals), use_str_keys


def _flat_key_value(key, vals):
    assert isinstance(key, (int,) + string_types), "unexpected type for keys: " + str(type(key))
    use_str_keys = isinstance(key, string_types)
    if isinstance(vals, NDArray):
        return [key], [vals], use_str_keys
    for value in vals:
        assert isinstance(value, NDArray)
    return [key] * len(vals), list(vals), use_str_keys


def _ctype_dict(param_dict):
    assert isinstance(param_dict, dict), "unexpected type for param_dict: " + str(type(param_dict))
    return c_str_array(list(param_dict.keys())), c_str_array([str(v) for v in param_dict.values()])


class KVStoreBase(object):
    """An abstract key-value store interface for data parallel training (base.py:75-245)."""
    OPTIMIZER = 'optimizer'
    kv_registry = {}

    def broadcast(self, key, value, out, priority=0):
        raise NotImplementedError()

    def pushpull(self, key, value, out=None, priority=0):
        raise NotImplementedError()

    def set_optimizer(self, optimizer):
        raise NotImplementedError()

    @staticmethod
    def is_capable(capability):
        raise NotImplementedError()

    def save_optimizer_states(self, fname, dump_optimizer=False):
        raise NotImplementedError()

    def load_optimizer_states(self, fname):
        raise NotImplementedError()

    @property
    def type(self):
        raise NotImplementedError()

    @property
    def rank(self):
        raise NotImplementedError()

    @property
    def num_workers(self):
        raise NotImplementedError()

    @staticmethod
    def register(klass):
        """Registers a new KVStore class under its lower-cased name (base.py:221-245)."""
        assert isinstance(klass, type)
        name = klass.__name__.lower()
        if name in KVStoreBase.kv_registry:
            warnings.warn('WARNING: New kvstore %s.%s is overriding existing kvstore %s.%s' % (
                klass.__module__, klass.__name__, KVStoreBase.kv_registry[name].__module__,
                KVStoreBase.kv_registry[name].__name__))
        KVStoreBase.kv_registry[name] = klass
        return klass


@KVStoreBase.register
class TestStore(KVStoreBase):
    """The pure front-end store of the reference's plugin-API conformance test (base.py:247-404):
    broadcast copies, pushpull sums on the first value's context. No optimizer capability."""

    def broadcast(self, key, value, out, priority=0):
        out = out if isinstance(out, list) else [out]
        for o in out:
            o[:] = value

    def pushpull(self, key, value, out=None, priority=0):
        if isinstance(value, NDArray):
            if out is not None:
                out = out if isinstance(out, list) else [out]
                for o in out:
                    o[:] = value
            return
        ctx = value[0].context
        reduced = value[0].as_in_context(ctx).copy()
        for v in value[1:]:
            reduced += v.as_in_context(ctx)
        targets = value if out is None else (out if isinstance(out, list) else [out])
        for o in targets:
            o[:] = reduced

    @staticmethod
    def is_capable(capability):
        if capability.lower() == KVStoreBase.OPTIMIZER:
            return False
        raise ValueError('Unknown capability: {}'.format(capability))

    @property
    def type(self):
        return 'teststore'

    @property
    def rank(self):
        return 0

    @property
    def num_workers(self):
        return 1


def create(name='local'):
    """Creates a new KVStore (base.py:406-455): registry first, then the native factory."""
    if not isinstance(name, string_types):
        raise TypeError('name must be a string')
    name = name.lower()
    if name in KVStoreBase.kv_registry:
        return KVStoreBase.kv_registry[name]()
    handle = KVStoreHandle()
    check_call(_LIB.MXKVStoreCreate(c_str(name), ctypes.byref(handle)))
    from .kvstore import KVStore
    return KVStore(handle)
