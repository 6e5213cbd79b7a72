"""ctypes binding of libb200kv.so -- the counterpart of python/mxnet/base.py for this path.

The library is the product: it is loaded (and, when missing, built in-tree with nvcc) at import
time, and every call goes through ``check_call`` exactly as the reference's bindings do
(python/mxnet/base.py:246-273). There is no pure-Python or CPU fallback: without the CUDA
extension the import raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libb200kv.so")


class MXNetError(Exception):
    """Error raised by the native library (same name as the reference's, base.py:76)."""


def _load():
    if not os.path.exists(_LIB_PATH):
        from . import build as _build
        _build.build()
    lib = ctypes.CDLL(_LIB_PATH, ctypes.RTLD_LOCAL)
    lib.MXGetLastError.restype = ctypes.c_char_p
    lib.B200KVBuildInfo.restype = ctypes.c_char_p
    return lib


_LIB = _load()

mx_uint = ctypes.c_uint
NDArrayHandle = ctypes.c_void_p
KVStoreHandle = ctypes.c_void_p
OpHandle = ctypes.c_void_p
string_types = (str,)


def check_call(ret):
    """python/mxnet/base.py:246-273: non-zero return -> MXNetError(MXGetLastError())."""
    if ret != 0:
        raise MXNetError(_LIB.MXGetLastError().decode("utf-8", "replace"))


def c_str(s):
    return ctypes.c_char_p(s.encode("utf-8"))


def c_str_array(strings):
    arr = (ctypes.c_char_p * len(strings))()
    arr[:] = [s.encode("utf-8") for s in strings]
    return arr


def c_array(ctype, values):
    return (ctype * len(values))(*values)


def c_handle_array(objs):
    arr = (ctypes.c_void_p * len(objs))()
    arr[:] = [o._hv for o in objs]
    return arr


def py_str(x):
    return x.decode("utf-8")


def gpu_count():
    n = ctypes.c_int()
    check_call(_LIB.MXGetGPUCount(ctypes.byref(n)))
    return n.value


def kernel_launch_count():
    n = ctypes.c_uint64()
    check_call(_LIB.B200KVGetKernelLaunchCount(ctypes.byref(n)))
    return n.value


def reset_kernel_launch_count():
    check_call(_LIB.B200KVResetKernelLaunchCount())


def last_kernel_info():
    name = ctypes.c_char_p()
    nbytes = ctypes.c_uint64()
    check_call(_LIB.B200KVGetLastKernelInfo(ctypes.byref(name), ctypes.byref(nbytes)))
    return py_str(name.value), nbytes.value


def set_stream(dev_id, cuda_stream):
    """Issue all library work for GPU `dev_id` on a caller-owned cudaStream_t (integer handle, e.g.
    torch.cuda.current_stream().cuda_stream). 0 means the CUDA legacy default stream -- what torch
    reports for its default stream -- and is passed as cudaStreamLegacy; None restores the
    library's own stream."""
    if cuda_stream is None:
        h = None
    elif cuda_stream == 0:
        h = 1  # cudaStreamLegacy
    else:
        h = cuda_stream
    check_call(_LIB.B200KVEngineSetStream(ctypes.c_int(dev_id), ctypes.c_void_p(h)))


def get_stream(dev_id):
    s = ctypes.c_void_p()
    check_call(_LIB.B200KVEngineGetStream(ctypes.c_int(dev_id), ctypes.byref(s)))
    return s.value or 0


def flush_all():
    """Issue every queued (bucketed) KVStore call now, without waiting for the device."""
    check_call(_LIB.B200KVFlushAll())


def waitall():
    check_call(_LIB.MXNDArrayWaitAll())
