"""ctypes binding of ``libxinv_b200.so`` (the C-ABI in ``include/xinv.h``).

There is no CPU fallback: if the shared library is missing it is built with
nvcc; if that fails, or if no B200 is visible when a solve is requested, the
call raises.
"""
import ctypes as C
import os
import threading

import numpy as np

from . import build as _build

BC_CODES = {"fixed": 0, "extend": 1, "periodic": 2}
ORDER_CODES = {"colour": 0, "color": 0, "redblack": 0, "red-black": 0,
               "lexicographic": 1, "lex": 1}
ENGINE_CODES = {"auto": 0, "colour": 1, "color": 1, "fused": 2, "resident": 3, "cluster": 4}
ENGINE_NAMES = {0: "auto", 1: "colour", 2: "fused", 3: "resident", 4: "cluster"}
MEM_HOST, MEM_DEVICE = 0, 1


class XinvOpts(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("ordering", C.c_int32),
                ("mem_space", C.c_int32), ("engine", C.c_int32),
                ("check_every", C.c_int32), ("profile", C.c_int32),
                ("coef_stride", C.c_int64 * 8), ("accel", C.c_int32), ("io_f32", C.c_int32)]


class XinvFlowAxis(C.Structure):
    _fields_ = [("uniform", C.c_int32), ("edge", C.c_int32), ("den", C.c_double), ("lo", C.c_double), ("hi", C.c_double),
                ("w", C.c_void_p)]


class XinvFlowDesc(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("comb", C.c_int32), ("swap", C.c_int32), ("nrows", C.c_int32),
                ("s1", C.c_double), ("s2", C.c_double), ("deg2m", C.c_double),
                ("y", XinvFlowAxis), ("x", XinvFlowAxis), ("rows", C.c_void_p)]


EDGE_CODES = {None: 0, "onesided": 0, "fixed": 1, "extend": 2, "reflect": 3, "periodic": 4}
FLOW_GRAD, FLOW_GM_LL, FLOW_GM_CART = 0, 1, 2


class XinvStats(C.Structure):
    _fields_ = [("sweeps_launched", C.c_int64), ("kernel_launches", C.c_int64),
                ("cell_updates", C.c_int64), ("solve_ms", C.c_double),
                ("h2d_ms", C.c_double), ("d2h_ms", C.c_double),
                ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
                ("engine", C.c_int32), ("ncolours", C.c_int32),
                ("sweep_ms", C.c_double), ("dom_ms", C.c_double),
                ("dom_launches", C.c_int64), ("slow_strips", C.c_int64),
                ("iters_per_pass", C.c_int32), ("row_coeffs", C.c_int32)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_}
        d["engine"] = ENGINE_NAMES.get(d["engine"], d["engine"])
        return d


# every symbol include/xinv.h declares: (name, restype, argtypes)
_vp, _i64, _dbl, _int = C.c_void_p, C.c_int64, C.c_double, C.c_int
_P = C.POINTER
_STD2D = [_vp] * 6 + [_i64, _i64, _i64, _int, _int] + [_dbl] * 5 + [_vp, _i64, _dbl, _P(XinvOpts)]
_GEN2D = [_vp] * 9 + [_i64, _i64, _i64, _int, _int] + [_dbl] * 7 + [_vp, _i64, _dbl, _P(XinvOpts)]
_STD2D_ROWS = [_vp] * 6 + [_dbl, _dbl, _i64, _i64, _i64, _int, _int] + [_dbl] * 5 + [_vp, _i64, _dbl, _P(XinvOpts)]
_GEN2D_ROWS = [_vp] * 4 + [_int, _dbl, _dbl, _dbl, _dbl, _i64, _i64, _i64, _int, _int] + [_dbl] * 7 + [_vp, _i64, _dbl, _P(XinvOpts)]
_STD3D_ROWS = [_vp] * 5 + [_i64, _vp, _dbl, _dbl, _i64, _i64, _i64, _i64, _int, _int, _int] + [_dbl] * 5 + [_vp, _i64, _dbl, _P(XinvOpts)]
_STD3D = [_vp] * 6 + [_i64, _i64, _i64, _i64, _int, _int, _int] + [_dbl] * 5 + [_vp, _i64, _dbl, _P(XinvOpts)]
_STD2DT = [_vp] * 8 + [_i64, _i64, _i64, _int, _int] + [_dbl] * 5 + [_vp, _i64, _dbl, _P(XinvOpts)]
_GEN3D = [_vp] * 10 + [_i64, _i64, _i64, _i64, _int, _int, _int] + [_dbl] * 8 + [_vp, _i64, _dbl, _P(XinvOpts)]
_STD1D = [_vp] * 5 + [_i64, _i64, _int] + [_dbl] * 3 + [_vp, _i64, _dbl, _P(XinvOpts)]
_BIH2D = [_vp] * 12 + [_i64, _i64, _i64, _int, _int] + [_dbl] * 9 + [_vp, _i64, _dbl, _P(XinvOpts)]
_STD2D_FRONT = [_vp] * 6 + [_dbl, _dbl, _i64, _i64, _i64, _int, _int] + [_dbl] * 5 + [_vp, _i64, _dbl, _P(XinvOpts)]
SYMBOLS = [
    ("xinv_create", _int, [_P(_vp), _int]),
    ("xinv_create_on_stream", _int, [_P(_vp), _int, _vp]),
    ("xinv_destroy", None, [_vp]),
    ("xinv_last_error", C.c_char_p, []),
    ("xinv_version", _int, []),
    ("xinv_get_stats", _int, [_vp, _P(XinvStats)]),
    ("xinv_device_count", _int, [_P(_int)]),
    ("xinv_synchronize", _int, [_vp]),
    ("xinv_timer_start", _int, [_vp]),
    ("xinv_timer_stop", _int, [_vp, _P(_dbl)]),
    ("xinv_host_alloc", _int, [_P(_vp), _i64]),
    ("xinv_host_free", _int, [_vp]),
    ("xinv_host_is_pinned", _int, [_vp, _P(_int)]),
    ("xinv_dev_alloc", _int, [_vp, _P(_vp), _i64]),
    ("xinv_dev_free", _int, [_vp, _vp]),
    ("xinv_memcpy_h2d", _int, [_vp, _vp, _vp, _i64]),
    ("xinv_memcpy_d2h", _int, [_vp, _vp, _vp, _i64]),
    ("xinv_flow2d", _int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _P(XinvFlowDesc), _P(XinvOpts)]),
    ("xinv_std2d", _int, _STD2D),
    ("xinv_std2d_rows", _int, _STD2D_ROWS),
    ("xinv_std2d_front", _int, _STD2D_FRONT),
    ("xinv_gen2d_rows", _int, _GEN2D_ROWS),
    ("xinv_std3d_rows", _int, _STD3D_ROWS),
    ("xinv_gen2d", _int, _GEN2D),
    ("xinv_std3d", _int, _STD3D),
    ("xinv_std2d_test", _int, _STD2DT),
    ("xinv_gen3d", _int, _GEN3D),
    ("xinv_std1d", _int, _STD1D),
    ("xinv_bih2d", _int, _BIH2D),
    ("xinv_std2d_begin", _int, _STD2D),
    ("xinv_gen2d_begin", _int, _GEN2D),
    ("xinv_std3d_begin", _int, _STD3D),
    ("xinv_step", _int, [_vp, _i64, _P(_i64)]),
    ("xinv_end", _int, [_vp]),
    ("xinv_nccl_unique_id", _int, [_vp]),
    ("xinv_nccl_init", _int, [_vp, _vp, _int, _int]),
    ("xinv_nccl_allreduce_active", _int, [_vp, _i64, _P(_i64)]),
    ("xinv_nccl_finalize", _int, [_vp]),
]

_lib = None
_lock = threading.Lock()


# return codes of include/xinv.h
E_ARG, E_CUDA, E_STATE, E_NOMEM, E_UNSUPPORTED, E_NCCL = -1, -2, -3, -4, -5, -6


class XinvError(RuntimeError):
    """A negative return code of the C-ABI (argument, CUDA or NCCL error); ``.code`` holds it
    (``E_UNSUPPORTED`` = -5 is what the facade tests for before it falls back to the host-built path)."""

    def __init__(self, msg, code=None):
        super().__init__(msg)
        self.code = code


def load():
    """dlopen libxinv_b200.so (building it first if needed) and bind symbols."""
    global _lib
    with _lock:
        if _lib is None:
            path = _build.build()
            L = C.CDLL(path)
            for name, res, args in SYMBOLS:
                f = getattr(L, name)          # AttributeError if the symbol is missing
                f.restype = res
                f.argtypes = args
            _lib = L
    return _lib


def check(rc):
    if rc != 0:
        msg = load().xinv_last_error()
        raise XinvError(f"libxinv_b200 error {rc}: {msg.decode() if msg else ''}", code=int(rc))


def device_count():
    n = C.c_int(0)
    rc = load().xinv_device_count(C.byref(n))
    return n.value if rc == 0 else 0


class Context:
    """One ``xinv_ctx``: a device, a stream and reusable staging buffers."""

    def __init__(self, device=0, stream=None):
        L = load()
        h = _vp()
        if stream is None:
            check(L.xinv_create(C.byref(h), int(device)))
        else:
            check(L.xinv_create_on_stream(C.byref(h), int(device), _vp(int(stream))))
        self._h = h
        self.device = int(device)
        # a ctx holds ONE problem slot and shared staging buffers: every begin..end sequence on it is
        # serialised (ctypes releases the GIL for the whole solve, so two Python threads -- e.g. the
        # dask threaded scheduler -- could otherwise interleave on the same ctx)
        self.lock = threading.RLock()

    def close(self):
        if getattr(self, "_h", None):
            load().xinv_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        if not self._h:
            raise XinvError("context is closed")
        return self._h

    def stats(self):
        s = XinvStats()
        check(load().xinv_get_stats(self.handle, C.byref(s)))
        return s.as_dict()

    def synchronize(self):
        check(load().xinv_synchronize(self.handle))

    def timer_start(self):
        check(load().xinv_timer_start(self.handle))

    def timer_stop(self):
        ms = C.c_double(0)
        check(load().xinv_timer_stop(self.handle, C.byref(ms)))
        return ms.value


_default_ctx = {}
_default_ctx_lock = threading.Lock()


def default_context(device=0):
    """Process-wide context per device (created on first use).  Calls that share it are serialised by
    ``Context.lock``; threads that want to overlap their solves create their own ``Context``."""
    with _default_ctx_lock:
        ctx = _default_ctx.get(device)
        if ctx is None or not ctx._h:
            if device_count() <= device:
                raise XinvError(
                    f"no CUDA device {device} visible: xinvert_b200 has no CPU fallback "
                    "(the SOR path runs only as sm_100a CUDA)")
            ctx = Context(device)
            _default_ctx[device] = ctx
    return ctx


_pipe_ctx = {}


def pipeline_contexts(device, n):
    """``n`` contexts on ``device`` (the first is the default context): every context has its own stream and
    staging buffers, so calls on different contexts overlap on the device -- one context's host<->device copies
    run while another's kernels do (solvers._execute pipelines the chunks of a batch this way)."""
    out = [default_context(device)]
    with _default_ctx_lock:
        extra = _pipe_ctx.setdefault(device, [])
        extra[:] = [c for c in extra if c._h]
        while len(extra) < n - 1:
            extra.append(Context(device))
        out += extra[:n - 1]
    return out


ACCEL_CODES = {None: 0, "none": 0, "chebyshev": 1}


def make_opts(ordering="colour", mem_space=MEM_HOST, engine="auto", check_every=0,
              coef_strides=None, profile=False, accel=None, io_f32=0):
    o = XinvOpts()
    o.struct_size = C.sizeof(XinvOpts)
    o.ordering = ORDER_CODES[ordering]
    o.mem_space = mem_space
    o.engine = ENGINE_CODES[engine]
    o.check_every = int(check_every)
    o.profile = 1 if profile else 0
    o.accel = ACCEL_CODES[accel]
    o.io_f32 = int(io_f32)
    for m in range(8):
        o.coef_stride[m] = -1
    if coef_strides:
        for m, s in enumerate(coef_strides[:8]):         # (arrays beyond the eighth are always dense: xinv_bih2d)
            o.coef_stride[m] = int(s)
    return o


class _PinnedBlock:
    """Owner of one cudaHostAlloc block; numpy views keep it alive via .base.  When the last view
    goes away the block returns to a small pool instead of being unpinned: page-locking tens of
    megabytes costs milliseconds, and result arrays of repeated solves have the same size."""
    _pool = {}                     # nbytes -> [ptr, ...]
    _pooled_bytes = 0
    POOL_LIMIT = 2 << 30           # at most 2 GiB of idle pinned memory is kept

    def __init__(self, nbytes):
        nbytes = int(nbytes)
        free = _PinnedBlock._pool.get(nbytes)
        if free:
            self.ptr = free.pop()
            _PinnedBlock._pooled_bytes -= nbytes
        else:
            p = _vp()
            check(load().xinv_host_alloc(C.byref(p), nbytes))
            self.ptr = p.value
        self.nbytes = nbytes
        self.__array_interface__ = {"shape": (nbytes,), "typestr": "|u1",
                                    "data": (self.ptr, False), "version": 3}

    def __del__(self):
        ptr, self.ptr = getattr(self, "ptr", None), None
        if not ptr or _lib is None:
            return
        try:
            if _PinnedBlock._pooled_bytes + self.nbytes <= _PinnedBlock.POOL_LIMIT:
                _PinnedBlock._pool.setdefault(self.nbytes, []).append(ptr)
                _PinnedBlock._pooled_bytes += self.nbytes
            else:
                _lib.xinv_host_free(_vp(ptr))
        except Exception:
            pass


def is_pinned(arr):
    """True if the numpy array's buffer is page-locked memory (e.g. from ``pinned_empty``)."""
    out = C.c_int(0)
    check(load().xinv_host_is_pinned(_vp(arr.ctypes.data), C.byref(out)))
    return bool(out.value)


_copy_pool = None


def parallel_copy(dst, src, nthreads=4):
    """dst[...] = src for large C-contiguous arrays, split over a few threads (numpy releases the GIL while it
    copies): one core moves ~10 GB/s, a PCIe 5 link takes 50."""
    global _copy_pool
    d, s = dst.reshape(-1), src.reshape(-1)
    n = d.size
    if n * d.itemsize < (8 << 20) or nthreads <= 1:
        np.copyto(d, s)
        return
    if _copy_pool is None:
        from concurrent.futures import ThreadPoolExecutor
        _copy_pool = ThreadPoolExecutor(max_workers=8)
    step = -(-n // nthreads)
    list(_copy_pool.map(lambda k: np.copyto(d[k * step:(k + 1) * step], s[k * step:(k + 1) * step]), range(nthreads)))


def pinned_empty(shape, dtype=np.float64):
    """numpy array over cudaHostAlloc'ed (page-locked) memory."""
    dtype = np.dtype(dtype)
    shape = tuple(int(x) for x in np.atleast_1d(shape))
    n = int(np.prod(shape)) * dtype.itemsize
    block = _PinnedBlock(max(n, 16))
    return np.asarray(block)[:n].view(dtype).reshape(shape)
