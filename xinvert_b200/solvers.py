"""ndarray-level face of the CUDA SOR path.

Two layers:

* ``solve_standard_2D / solve_general_2D / solve_standard_3D`` -- batched: every
  leading (non-core) axis of ``S`` is flattened into the ``batch`` argument of
  ONE C-ABI call (``include/xinv.h``).  This is what replaces the reference's
  serial slice loop (``core.py:129-139``, ``:418-428``, ``:59-69``).
* ``invert_standard_2D / invert_general_2D / invert_standard_3D`` -- single
  slice, with the exact positional signatures of the reference's numba kernels
  (``numbas.py:216-219``, ``:988-991``, ``:16-19``) so parity tests read like
  calls into the reference.

Arrays may be numpy arrays (host; staged by the library) or CUDA tensors /
objects exposing ``data_ptr()`` (device; used in place).  There is no CPU
implementation behind these functions.
"""
import ctypes as C
import os

import numpy as np

from . import _lib

_UNDEF = -9.99e8


def _is_device_array(a):
    return hasattr(a, "data_ptr") and getattr(a, "is_cuda", False)


def _host_f64(a, name):
    """C-contiguous float64 view/copy of a host array (float32 inputs are
    promoted; the reference would iterate in mixed precision for those)."""
    arr = np.asarray(a)
    if arr.dtype != np.float64 or not arr.flags["C_CONTIGUOUS"]:
        arr = np.ascontiguousarray(arr, dtype=np.float64)
    return arr


class _Operands:
    """Pointer + batch stride of S and of every coefficient array."""

    def __init__(self, S, coefs, core_ndim):
        self.device = _is_device_array(S)
        self.keep = []                      # keep temporaries alive during the call
        if self.device:
            import torch
            if S.dtype != torch.float64 or not S.is_contiguous():
                raise ValueError("device S must be a contiguous float64 tensor")
            self.shape = tuple(S.shape)
            self.S_ptr = S.data_ptr()
            self.S_host = None
        else:
            if not isinstance(S, np.ndarray):
                raise TypeError("S must be a numpy array or a CUDA tensor (it is updated in place)")
            self.shape = S.shape
            Sc = _host_f64(S, "S")
            self.S_host = Sc
            self.S_orig = S
            self.S_ptr = Sc.ctypes.data
        if len(self.shape) < core_ndim:
            raise ValueError(f"S needs at least {core_ndim} dimensions")
        self.core = tuple(int(n) for n in self.shape[-core_ndim:])
        self.batch = int(np.prod(self.shape[:-core_ndim], dtype=np.int64)) if len(self.shape) > core_ndim else 1
        self.N = int(np.prod(self.core, dtype=np.int64))
        self.ptrs, self.strides = [], []
        for name, a in coefs:
            if a is None:
                self.ptrs.append(None)
                self.strides.append(0)
                continue
            if _is_device_array(a) != self.device:
                raise ValueError(f"{name}: S and coefficients must live in the same memory space")
            if self.device:
                import torch
                if a.dtype != torch.float64 or not a.is_contiguous():
                    raise ValueError(f"{name} must be a contiguous float64 tensor")
                shp, ptr = tuple(a.shape), a.data_ptr()
            else:
                a = _host_f64(a, name)
                self.keep.append(a)
                shp, ptr = a.shape, a.ctypes.data
            shp = tuple(int(n) for n in shp)
            if shp == tuple(self.shape):
                stride = self.N
            elif shp[-core_ndim:] == self.core and all(d == 1 for d in shp[:-core_ndim]):
                stride = 0                  # one slice shared by the whole batch
            else:
                raise ValueError(f"{name} has shape {shp}; expected {self.shape} or {self.core}")
            self.ptrs.append(ptr)
            self.strides.append(stride)

    def finish(self):
        """Copy the result back if S had to be converted for the call."""
        if not self.device and self.S_host is not self.S_orig:
            self.S_orig[...] = self.S_host


def _flags_array(flags, batch):
    """Host flags [batch][3], seeded from the caller's flags (numbas keeps the
    incoming values when a slice overflows in its first sweep)."""
    f = np.asarray(flags, dtype=np.float64)
    if f.shape == (3,):
        out = np.tile(f, (batch, 1))
    elif f.shape == (batch, 3):
        out = np.ascontiguousarray(f)
    else:
        raise ValueError(f"flags must have shape (3,) or ({batch}, 3)")
    return np.ascontiguousarray(out, dtype=np.float64)


def _zero_to_none(B):
    """``B`` identically zero selects the 5-point red-black path (value-identical:
    the cross terms are exactly 0)."""
    if B is None:
        return None
    if _is_device_array(B):
        return None if not bool((B != 0).any().item()) else B
    Bn = np.asarray(B)
    return None if not Bn.any() else B


MAX_BATCH = 65535                 # slices per C-ABI call (the slice index rides on a 16-bit grid dimension)


def _batch_chunks(batch):
    """(lo, hi) blocks of at most MAX_BATCH slices: the reference's serial loop takes any number of
    non-core slices (hourly data x levels easily exceeds 65535), so longer batches are cut here."""
    if batch <= 0:
        return [(0, 0)]
    return [(lo, min(lo + MAX_BATCH, batch)) for lo in range(0, batch, MAX_BATCH)]


def _merge_stats(acc, st):
    """Statistics of a call that was cut into several C-ABI calls: counters and times add up."""
    if acc is None:
        return dict(st)
    for k in ("sweeps_launched", "kernel_launches", "cell_updates", "solve_ms", "h2d_ms", "d2h_ms", "h2d_bytes",
              "d2h_bytes", "dom_ms", "dom_launches"):
        acc[k] += st[k]
    return acc


def _sync_torch_stream(device_array):
    """Device operands: the library works on the ctx's own (non-blocking) stream, which is not ordered
    against torch's current stream -- wait for whatever produced the tensors (S.copy_, computed
    coefficients ...) before the library reads them.  (The call itself returns synchronised.)"""
    import torch
    torch.cuda.current_stream(device_array.device).synchronize()


# ---------------------------------------------------------------------------
# Execution plan of one batched call: slices are independent solves (core.py:129-139), so the flattened batch
# may be cut anywhere.  It is cut (a) over the devices named in ``devices`` (contiguous blocks,
# distributed.shard_bounds -- what a multi-process launch does rank by rank, from one process: one context and
# one thread per GPU), and (b) on each device into chunks that are pipelined through two contexts (two streams,
# two sets of staging buffers): while one chunk is being solved the next one is copied to the device and the
# previous result is copied back.  Every slice still stops on its own test; results do not depend on the cut.
# ---------------------------------------------------------------------------
PIPE_MIN_CELLS = int(os.environ.get("XINV_PIPE_MIN_CELLS", 8 << 20))          # a chunk keeps at least this many cells (the fused kernels lose efficiency below)
PIPE_STREAMS = 2
PIPE_MAX_CHUNKS = int(os.environ.get("XINV_PIPE_MAX_CHUNKS", 2 * PIPE_STREAMS))


def _plan(batch, cells_per_slice, devices, pipelined):
    """[(device, worker, lo, hi), ...]: the blocks of slices and who runs them."""
    from .distributed import shard_bounds
    devices = list(devices) if devices else [0]
    items = []
    for r, dev in enumerate(devices):
        lo, hi = shard_bounds(batch, len(devices), r)
        n = hi - lo
        if n <= 0:
            continue
        nchunk = 1
        if pipelined:
            nchunk = max(1, min(n, (n * cells_per_slice) // PIPE_MIN_CELLS))
            if nchunk > 1:
                nchunk = max(nchunk, PIPE_STREAMS)
            nchunk = min(nchunk, PIPE_MAX_CHUNKS)           # a few chunks are enough to hide the copies
        nchunk = max(nchunk, -(-n // MAX_BATCH))
        for k in range(nchunk):
            clo, chi = shard_bounds(n, nchunk, k)
            if chi > clo:
                items.append((dev, k % PIPE_STREAMS if nchunk > 1 else 0, lo + clo, lo + chi))
    return items


def _execute(call, batch, cells_per_slice, ctx, devices, host, _contexts=None):
    """Run ``call(ctx, lo, hi)`` (one C-ABI call on slices [lo, hi), returns that call's stats) over the whole batch.
    An explicit ``ctx`` (and no ``devices``) pins everything to that context; otherwise the plan above is used."""
    if batch <= 0:
        c = ctx or (_contexts or _lib.pipeline_contexts)((list(devices) if devices else [0])[0], 1)[0]
        with c.lock:
            return call(c, 0, 0)
    if ctx is not None and not devices:
        stats = None
        with ctx.lock:
            for lo, hi in _batch_chunks(batch):
                stats = _merge_stats(stats, call(ctx, lo, hi))
        return stats
    if ctx is not None and devices and list(devices) == [ctx.device]:
        first = {ctx.device: ctx}
    else:
        first = {}
    items = _plan(batch, cells_per_slice, devices, pipelined=host)
    workers = {}
    for dev, w, lo, hi in items:
        workers.setdefault((dev, w), []).append((lo, hi))
    ctxs = {}
    for dev in sorted({d for d, _ in workers}):
        n = 1 + max(w for d, w in workers if d == dev)
        cs = (_contexts or _lib.pipeline_contexts)(dev, n)
        if dev in first:
            cs = [first[dev]] + [c for c in cs if c is not first[dev]][:n - 1]
        for w in range(n):
            ctxs[(dev, w)] = cs[w]

    def work(key):
        c, st = ctxs[key], None
        for lo, hi in workers[key]:
            with c.lock:
                st = _merge_stats(st, call(c, lo, hi))
        return st

    keys = sorted(workers)
    if len(keys) == 1:
        return work(keys[0])
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=len(keys)) as ex:       # ctypes releases the GIL for the whole C call
        results = list(ex.map(work, keys))
    stats = None
    for st in results:
        stats = _merge_stats(stats, st)
    stats["pipeline"] = {"devices": sorted({d for d, _ in keys}), "workers": len(keys), "chunks": len(items)}
    return stats


def _run(kind, ops, fn_args_tail, flags, ordering, engine, check_every, ctx, profile=False, S_dev=None, devices=None,
         accel=None):
    L = _lib.load()
    fl = _flags_array(flags, ops.batch)
    fn = {"std2d": L.xinv_std2d, "gen2d": L.xinv_gen2d, "std3d": L.xinv_std3d, "std2dt": L.xinv_std2d_test,
          "gen3d": L.xinv_gen3d, "std1d": L.xinv_std1d, "bih2d": L.xinv_bih2d}[kind]
    if ops.device:
        if S_dev is not None:
            _sync_torch_stream(S_dev)
        if ctx is None:
            ctx = _lib.default_context(getattr(getattr(S_dev, "device", None), "index", None) or 0)
        devices = None                           # the tensors live on one device

    def call(c, lo, hi):
        opts = _lib.make_opts(ordering=ordering, mem_space=_lib.MEM_DEVICE if ops.device else _lib.MEM_HOST,
                              engine=engine, check_every=check_every, coef_strides=ops.strides, profile=profile,
                              accel=accel)
        off = lambda p, stride: C.c_void_p(p + 8 * lo * stride) if p is not None else None
        ptr_args = [off(ops.S_ptr, ops.N)] + [off(p, st) for p, st in zip(ops.ptrs, ops.strides)]
        rc = fn(c.handle, *ptr_args, *fn_args_tail(fl[lo:hi], hi - lo), C.byref(opts))
        _lib.check(rc)
        return c.stats()

    stats = _execute(call, ops.batch, ops.N, ctx, devices, host=not ops.device)
    ops.finish()
    return fl, stats


def solve_standard_2D(S, A, B, C_, F, BCy, BCx, delxSqr, ratioQtr, ratioSqr, optArg,
                      undef=_UNDEF, flags=(0.0, 1.0, 0.0), mxLoop=5000, tolerance=1e-8,
                      ordering="colour", engine="auto", check_every=0, ctx=None, profile=False, devices=None,
                      accel=None):
    """Batched ``invert_standard_2D`` (numbas.py:215-416) over S[..., ny, nx], in place.

    ``devices``: GPUs to cut the batch over (one context and one thread per GPU, from this process).
    Returns ``(flags[batch, 3], stats)``."""
    B = _zero_to_none(B)
    ops = _Operands(S, [("A", A), ("B", B), ("C", C_), ("F", F)], 2)
    ny, nx = ops.core
    tail = lambda fl, nb: (nb, ny, nx, _lib.BC_CODES[BCy], _lib.BC_CODES[BCx],
                           float(delxSqr), float(ratioQtr), float(ratioSqr), float(optArg), float(undef),
                           C.c_void_p(fl.ctypes.data), int(mxLoop), float(tolerance))
    return _run("std2d", ops, tail, flags, ordering, engine, check_every, ctx, profile, S_dev=S, devices=devices, accel=accel)


def solve_standard_2D_rows(F_user, A_rows, C_rows, F_row_scale, user_undef, out_undef, BCy, BCx,
                           delxSqr, ratioQtr, ratioSqr, optArg, undef=_UNDEF, flags=(0.0, 1.0, 0.0),
                           mxLoop=5000, tolerance=1e-8, check_every=0, ctx=None, out=None, devices=None, accel=None):
    """Poisson-type front end (``xinv_std2d_rows``): the user's forcing ``F_user[..., ny, nx]`` (host
    numpy array or CUDA tensor; cells equal to ``user_undef`` -- any NaN when that is NaN -- are
    land), per-row coefficients ``A_rows[ny]``, ``C_rows[ny]`` and an optional per-row forcing scale;
    returns ``(S, flags[batch, 3], stats)`` with ``S`` solved from a zero initial guess and land set
    to ``out_undef``.  What apps.__mask_FS / __coeffs_Poisson / __template's de-masking do on the
    host (apps.py:2112-2159, :1397-1437, :1386-1392) happens on the device; results are identical.
    Raises ``XinvError`` (code -5) when the fused engine cannot take the problem."""
    L = _lib.load()
    device = _is_device_array(F_user)
    if device:
        ctx = ctx or _lib.default_context(F_user.device.index or 0)
        devices = None
        import torch
        if F_user.dtype != torch.float64 or not F_user.is_contiguous():
            raise ValueError("device forcing must be a contiguous float64 tensor")
        shape = tuple(F_user.shape)
        S = out if out is not None else torch.empty_like(F_user)
        rows = [torch.as_tensor(np.ascontiguousarray(v, dtype=np.float64)).to(F_user.device) if v is not None else None
                for v in (A_rows, C_rows, F_row_scale)]
        torch.cuda.current_stream(F_user.device).synchronize()    # the row vectors were copied on torch's stream
        ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        F_ptr, S_ptr = ptr(F_user), ptr(S)
        keep = rows
    else:
        # float32 forcing stays float32 on its way to the device and so does the result on its way back
        # (xinv_opts.io_f32: half the PCIe bytes; widened / narrowed on the device, the solve is float64 either way)
        f32 = isinstance(F_user, np.ndarray) and F_user.dtype == np.float32
        Fh = np.ascontiguousarray(F_user) if f32 else _host_f64(F_user, "F")
        shape = Fh.shape
        # (page-locked and pooled: the device-to-host copy of the result runs at full PCIe rate)
        S = out if out is not None else _lib.pinned_empty(shape, Fh.dtype)
        if S.dtype != Fh.dtype or not S.flags["C_CONTIGUOUS"] or S.shape != shape:
            raise ValueError(f"out must be a C-contiguous {Fh.dtype} array of the forcing's shape")
        rows = [np.ascontiguousarray(v, dtype=np.float64) if v is not None else None
                for v in (A_rows, C_rows, F_row_scale)]
        ptr = lambda a: C.c_void_p(a.ctypes.data) if a is not None else None
        F_ptr, S_ptr = ptr(Fh), ptr(S)
        keep = rows + [Fh]
    if len(shape) < 2:
        raise ValueError("forcing needs at least 2 dimensions")
    ny, nx = int(shape[-2]), int(shape[-1])
    batch = int(np.prod(shape[:-2], dtype=np.int64)) if len(shape) > 2 else 1
    for v, name in ((rows[0], "A_rows"), (rows[1], "C_rows"), (rows[2], "F_row_scale")):
        if v is not None and tuple(v.shape) != (ny,):
            raise ValueError(f"{name} must have shape ({ny},)")
    fl = _flags_array(flags, batch)
    stage_F = (not device) and not _lib.is_pinned(Fh)      # pageable forcing: through pinned buffers, chunk by chunk
    io32 = 3 if (not device and Fh.dtype == np.float32) else 0
    item = 4 if io32 else 8
    if io32:
        user_undef = float(np.float32(user_undef))         # the value the float32 cells actually hold
        out_undef = float(np.float32(out_undef))

    def call(c, lo, hi):
        opts = _lib.make_opts(mem_space=_lib.MEM_DEVICE if device else _lib.MEM_HOST, check_every=check_every, accel=accel,
                              io_f32=io32)
        off = lambda p: C.c_void_p(p.value + item * lo * ny * nx)
        Fc = off(F_ptr)
        if stage_F and hi > lo:
            buf = _lib.pinned_empty((hi - lo, ny, nx), Fh.dtype)      # pooled; filled by a few threads while the other chunk solves
            _lib.parallel_copy(buf, Fh.reshape(batch, ny, nx)[lo:hi])
            Fc = C.c_void_p(buf.ctypes.data)
        rc = L.xinv_std2d_rows(c.handle, off(S_ptr), ptr(rows[0]), ptr(rows[1]), Fc, ptr(rows[2]),
                               float(user_undef), float(out_undef), hi - lo, ny, nx, _lib.BC_CODES[BCy],
                               _lib.BC_CODES[BCx], float(delxSqr), float(ratioQtr), float(ratioSqr), float(optArg),
                               float(undef), C.c_void_p(fl[lo:hi].ctypes.data), int(mxLoop), float(tolerance),
                               C.byref(opts))
        _lib.check(rc)
        return c.stats()

    stats = _execute(call, batch, ny * nx, ctx, devices, host=not device)
    del keep
    return S, fl, stats


def solve_standard_2D_front(F_user, A, B, C_, user_undef, out_undef, BCy, BCx, delxSqr, ratioQtr, ratioSqr, optArg,
                            undef=_UNDEF, flags=(0.0, 1.0, 0.0), mxLoop=5000, tolerance=1e-8, engine="auto", check_every=0,
                            ctx=None, devices=None, accel=None):
    """Dense front end (``xinv_std2d_front``; invert_Eliassen): the user's forcing ``F_user[..., ny, nx]`` (host numpy
    array; cells equal to ``user_undef`` -- any NaN when that is NaN -- are land) and the coefficient arrays A, B
    (``None`` = 0), C as the caller holds them: S's shape, or one ``[ny, nx]`` slice shared by the batch.  Returns
    ``(S, flags[batch, 3], stats)``: solved from a zero initial guess, land set to ``out_undef`` (what apps.__mask_FS and
    apps.__template's de-masking do on full-size host arrays happens on the device)."""
    L = _lib.load()
    Fh = _host_f64(F_user, "F")
    S = _lib.pinned_empty(Fh.shape)
    B = _zero_to_none(B)
    ops = _Operands(S, [("A", A), ("B", B), ("C", C_), ("F", Fh)], 2)
    ny, nx = ops.core
    fl = _flags_array(flags, ops.batch)

    def call(c, lo, hi):
        opts = _lib.make_opts(mem_space=_lib.MEM_HOST, engine=engine, check_every=check_every, coef_strides=ops.strides, accel=accel)
        off = lambda p, stride: C.c_void_p(p + 8 * lo * stride) if p is not None else None
        ptr = [off(p, st) for p, st in zip(ops.ptrs, ops.strides)]
        rc = L.xinv_std2d_front(c.handle, off(ops.S_ptr, ops.N), ptr[0], ptr[1], ptr[2], ptr[3], float(user_undef),
                                float(out_undef), hi - lo, ny, nx, _lib.BC_CODES[BCy], _lib.BC_CODES[BCx], float(delxSqr),
                                float(ratioQtr), float(ratioSqr), float(optArg), float(undef),
                                C.c_void_p(fl[lo:hi].ctypes.data), int(mxLoop), float(tolerance), C.byref(opts))
        _lib.check(rc)
        return c.stats()

    stats = _execute(call, ops.batch, ops.N, ctx, devices, host=True)
    return S, fl, stats


def solve_general_2D_rows(G_user, rows, g_mode, g_p1, g_p2, user_undef, out_undef, BCy, BCx, delx, delxSqr, ratio,
                          ratioQtr, ratioSqr, optArg, undef=_UNDEF, flags=(0.0, 1.0, 0.0), mxLoop=5000,
                          tolerance=1e-8, check_every=0, ctx=None, devices=None, accel=None):
    """General-form front end (``xinv_gen2d_rows``): the user's forcing ``G_user[..., ny, nx]`` (host
    numpy array), ``rows[5, ny]`` = A, C, D, E, F of every row, and the forcing transform (``g_mode`` 0:
    G = forcing; 1: G = ((-forcing) / g_p1) / g_p2); returns ``(S, flags[batch, 3], stats)`` with ``S``
    solved from a zero initial guess and land set to ``out_undef`` (what apps.__mask_FS /
    __coeffs_GillMatsuno / __coeffs_Stommel / __template do on the host, on the device)."""
    L = _lib.load()
    f32 = isinstance(G_user, np.ndarray) and G_user.dtype == np.float32      # xinv_opts.io_f32: float32 both ways
    Gh = np.ascontiguousarray(G_user) if f32 else _host_f64(G_user, "G")
    item, io32 = (4, 3) if f32 else (8, 0)
    if f32:
        user_undef, out_undef = float(np.float32(user_undef)), float(np.float32(out_undef))
    shape = Gh.shape
    if len(shape) < 2:
        raise ValueError("forcing needs at least 2 dimensions")
    ny, nx = int(shape[-2]), int(shape[-1])
    batch = int(np.prod(shape[:-2], dtype=np.int64)) if len(shape) > 2 else 1
    rows = np.ascontiguousarray(rows, dtype=np.float64)
    if rows.shape != (5, ny):
        raise ValueError(f"rows must have shape (5, {ny})")
    S = _lib.pinned_empty(shape, Gh.dtype)
    fl = _flags_array(flags, batch)

    def call(c, lo, hi):
        opts = _lib.make_opts(mem_space=_lib.MEM_HOST, check_every=check_every, accel=accel, io_f32=io32)
        o = item * lo * ny * nx
        rc = L.xinv_gen2d_rows(c.handle, C.c_void_p(S.ctypes.data + o), C.c_void_p(rows.ctypes.data),
                               C.c_void_p(Gh.ctypes.data + o), int(g_mode), float(g_p1), float(g_p2),
                               float(user_undef), float(out_undef), hi - lo, ny, nx, _lib.BC_CODES[BCy],
                               _lib.BC_CODES[BCx], float(delx), float(delxSqr), float(ratio), float(ratioQtr),
                               float(ratioSqr), float(optArg), float(undef), C.c_void_p(fl[lo:hi].ctypes.data),
                               int(mxLoop), float(tolerance), C.byref(opts))
        _lib.check(rc)
        return c.stats()

    stats = _execute(call, batch, ny * nx, ctx, devices, host=True)
    return S, fl, stats


def solve_standard_3D_rows(F_user, rows, N2, n2_strides, user_undef, out_undef, BCz, BCy, BCx, delxSqr, ratio2Sqr,
                           ratio1Sqr, optArg, undef=_UNDEF, flags=(0.0, 1.0, 0.0), mxLoop=5000, tolerance=1e-8,
                           check_every=0, ctx=None, devices=None, accel=None):
    """invert_omega front end (``xinv_std3d_rows``): the user's forcing ``F_user[..., nz, ny, nx]`` (host numpy
    array), ``rows[4, ny]`` (A; the factor of B = N2 * rows[1]; the divisor of C = N2 / rows[2]; the forcing
    scale) and ``N2`` (a float64 buffer read through four element strides: batch, level, row, column; 0 =
    broadcast).  Returns ``(S, flags[batch, 3], stats)``: solved from a zero initial guess, land = ``out_undef``
    (what apps.__mask_FS / __coeffs_omega / __template do on the host, on the device)."""
    L = _lib.load()
    f32 = isinstance(F_user, np.ndarray) and F_user.dtype == np.float32      # xinv_opts.io_f32: float32 both ways
    Fh = np.ascontiguousarray(F_user) if f32 else _host_f64(F_user, "F")
    item, io32 = (4, 3) if f32 else (8, 0)
    if f32:
        user_undef, out_undef = float(np.float32(user_undef)), float(np.float32(out_undef))
    shape = Fh.shape
    if len(shape) < 3:
        raise ValueError("forcing needs at least 3 dimensions")
    nz, ny, nx = (int(n) for n in shape[-3:])
    batch = int(np.prod(shape[:-3], dtype=np.int64)) if len(shape) > 3 else 1
    rows = np.ascontiguousarray(rows, dtype=np.float64)
    if rows.shape != (4, ny):
        raise ValueError(f"rows must have shape (4, {ny})")
    n2 = np.ascontiguousarray(N2, dtype=np.float64).reshape(-1)
    st4 = (C.c_int64 * 4)(*[int(v) for v in n2_strides])
    S = _lib.pinned_empty(shape, Fh.dtype)
    fl = _flags_array(flags, batch)
    N = nz * ny * nx
    stage_F = not _lib.is_pinned(Fh)             # pageable operands: through pinned buffers (parallel copies)
    if n2.size * 8 >= (8 << 20) and not _lib.is_pinned(n2):
        n2p = _lib.pinned_empty(n2.shape)
        _lib.parallel_copy(n2p, n2)
        n2 = n2p

    def call(c, lo, hi):
        opts = _lib.make_opts(mem_space=_lib.MEM_HOST, check_every=check_every, accel=accel, io_f32=io32)
        o = item * lo * N
        n2_off = 8 * lo * int(n2_strides[0])
        Fc = Fh.ctypes.data + o
        if stage_F and hi > lo:
            buf = _lib.pinned_empty((hi - lo, nz, ny, nx), Fh.dtype)
            _lib.parallel_copy(buf, Fh.reshape(batch, nz, ny, nx)[lo:hi])
            Fc = buf.ctypes.data
        rc = L.xinv_std3d_rows(c.handle, C.c_void_p(S.ctypes.data + o), C.c_void_p(rows.ctypes.data),
                               C.c_void_p(n2.ctypes.data + n2_off), st4, int(n2.size - lo * int(n2_strides[0])),
                               C.c_void_p(Fc), float(user_undef), float(out_undef), hi - lo, nz, ny, nx,
                               _lib.BC_CODES[BCz], _lib.BC_CODES[BCy], _lib.BC_CODES[BCx], float(delxSqr),
                               float(ratio2Sqr), float(ratio1Sqr), float(optArg), float(undef),
                               C.c_void_p(fl[lo:hi].ctypes.data), int(mxLoop), float(tolerance), C.byref(opts))
        _lib.check(rc)
        return c.stats()

    stats = _execute(call, batch, N, ctx, devices, host=True)
    return S, fl, stats


def solve_general_2D(S, A, B, C_, D, E, F, G, BCy, BCx, delx, delxSqr, ratio, ratioQtr,
                     ratioSqr, optArg, undef=_UNDEF, flags=(0.0, 1.0, 0.0), mxLoop=5000,
                     tolerance=1e-8, ordering="colour", engine="auto", check_every=0, ctx=None, profile=False, devices=None,
                      accel=None):
    """Batched ``invert_general_2D`` (numbas.py:987-1201) over S[..., ny, nx], in place."""
    B = _zero_to_none(B)
    ops = _Operands(S, [("A", A), ("B", B), ("C", C_), ("D", D), ("E", E), ("F", F), ("G", G)], 2)
    ny, nx = ops.core
    tail = lambda fl, nb: (nb, ny, nx, _lib.BC_CODES[BCy], _lib.BC_CODES[BCx],
                           float(delx), float(delxSqr), float(ratio), float(ratioQtr), float(ratioSqr),
                           float(optArg), float(undef), C.c_void_p(fl.ctypes.data), int(mxLoop), float(tolerance))
    return _run("gen2d", ops, tail, flags, ordering, engine, check_every, ctx, profile, S_dev=S, devices=devices, accel=accel)


def solve_standard_3D(S, A, B, C_, F, BCz, BCy, BCx, delxSqr, ratio2Sqr, ratio1Sqr, optArg,
                      undef=_UNDEF, flags=(0.0, 1.0, 0.0), mxLoop=5000, tolerance=1e-8,
                      ordering="colour", engine="auto", check_every=0, ctx=None, profile=False, devices=None,
                      accel=None):
    """Batched ``invert_standard_3D`` (numbas.py:15-212) over S[..., nz, ny, nx], in place."""
    ops = _Operands(S, [("A", A), ("B", B), ("C", C_), ("F", F)], 3)
    nz, ny, nx = ops.core
    tail = lambda fl, nb: (nb, nz, ny, nx, _lib.BC_CODES[BCz], _lib.BC_CODES[BCy], _lib.BC_CODES[BCx],
                           float(delxSqr), float(ratio2Sqr), float(ratio1Sqr), float(optArg), float(undef),
                           C.c_void_p(fl.ctypes.data), int(mxLoop), float(tolerance))
    return _run("std3d", ops, tail, flags, ordering, engine, check_every, ctx, profile, S_dev=S, devices=devices, accel=accel)


# ---------------------------------------------------------------------------
# cal_flow epilogue (xinv_flow2d)
# ---------------------------------------------------------------------------
# ---- the remaining kernels of numbas.py (generic colour engine) ----
def solve_standard_2D_test(S, A, B, C_, D, E, F, BCy, BCx, delxSqr, ratioQtr, ratioSqr, optArg, undef=_UNDEF,
                           flags=(0.0, 1.0, 0.0), mxLoop=5000, tolerance=1e-8, check_every=0, ctx=None, devices=None):
    """Batched ``invert_standard_2D_test`` (numbas.py:420-629) over S[..., ny, nx], in place."""
    ops = _Operands(S, [("A", A), ("B", B), ("C", C_), ("D", D), ("E", E), ("F", F)], 2)
    ny, nx = ops.core
    tail = lambda fl, nb: (nb, ny, nx, _lib.BC_CODES[BCy], _lib.BC_CODES[BCx], float(delxSqr), float(ratioQtr), float(ratioSqr),
                           float(optArg), float(undef), C.c_void_p(fl.ctypes.data), int(mxLoop), float(tolerance))
    return _run("std2dt", ops, tail, flags, "colour", "auto", check_every, ctx, False, S_dev=S, devices=devices)


def solve_general_3D(S, A, B, C_, D, E, F, G, H, BCz, BCy, BCx, delx, delxSqr, ratio2, ratio1, ratio2Sqr, ratio1Sqr, optArg,
                     undef=_UNDEF, flags=(0.0, 1.0, 0.0), mxLoop=5000, tolerance=1e-8, check_every=0, ctx=None, devices=None):
    """Batched ``invert_general_3D`` (numbas.py:745-984) over S[..., nz, ny, nx], in place."""
    ops = _Operands(S, [(k, v) for k, v in zip("ABCDEFGH", (A, B, C_, D, E, F, G, H))], 3)
    nz, ny, nx = ops.core
    tail = lambda fl, nb: (nb, nz, ny, nx, _lib.BC_CODES[BCz], _lib.BC_CODES[BCy], _lib.BC_CODES[BCx], float(delx),
                           float(delxSqr), float(ratio2), float(ratio1), float(ratio2Sqr), float(ratio1Sqr), float(optArg),
                           float(undef), C.c_void_p(fl.ctypes.data), int(mxLoop), float(tolerance))
    return _run("gen3d", ops, tail, flags, "colour", "auto", check_every, ctx, False, S_dev=S, devices=devices)


def solve_standard_1D(S, A, B, F, BCx, delxSqr, optArg, undef=_UNDEF, flags=(0.0, 1.0, 0.0), mxLoop=5000, tolerance=1e-8,
                      check_every=0, ctx=None, devices=None):
    """Batched ``invert_standard_1D`` (numbas.py:632-742) over S[..., nx], in place."""
    ops = _Operands(S, [("A", A), ("B", B), ("F", F)], 1)
    (nx,) = ops.core
    tail = lambda fl, nb: (nb, nx, _lib.BC_CODES[BCx], float(delxSqr), float(optArg), float(undef),
                           C.c_void_p(fl.ctypes.data), int(mxLoop), float(tolerance))
    return _run("std1d", ops, tail, flags, "colour", "auto", check_every, ctx, False, S_dev=S, devices=devices)


def solve_general_bih_2D(S, A, B, C_, D, E, F, G, H, I, J, BCy, BCx, delxSSr, delxTr, delxSqr, ratio, ratioSSr, ratioQtr,
                         ratioSqr, optArg, undef=_UNDEF, flags=(0.0, 1.0, 0.0), mxLoop=5000, tolerance=1e-8, check_every=0,
                         ctx=None, devices=None):
    """Batched ``invert_general_bih_2D`` (numbas.py:1204-1586) over S[..., ny, nx], in place."""
    ops = _Operands(S, [(k, v) for k, v in zip("ABCDEFGHIJ", (A, B, C_, D, E, F, G, H, I, J))], 2)
    if any(st == 0 for st in ops.strides[8:]):
        raise ValueError("I and J must have S's shape (one slice per batch entry)")
    ny, nx = ops.core
    tail = lambda fl, nb: (nb, ny, nx, _lib.BC_CODES[BCy], _lib.BC_CODES[BCx], float(delxSSr), float(delxTr), float(delxSqr),
                           float(ratio), float(ratioSSr), float(ratioQtr), float(ratioSqr), float(optArg), float(undef),
                           C.c_void_p(fl.ctypes.data), int(mxLoop), float(tolerance))
    return _run("bih2d", ops, tail, flags, "colour", "auto", check_every, ctx, False, S_dev=S, devices=devices)


def axis_diff(coord, edge=None, fill=(0.0, 0.0)):
    """How ``numpy.gradient`` differentiates along an axis with these coordinate values -- decided here exactly as
    numpy decides it -- for an unpadded line (``edge=None``: DataArray.differentiate, one-sided ends) or for the line
    padded by finitediffs.padBCs (``edge`` in fixed / extend / reflect / periodic; finitediffs.py:548-606: the two
    extra coordinate values are linear extrapolations).  Returns a dict for ``flow_2d``."""
    x = np.asarray(coord, dtype=np.float64)
    n = x.size
    if edge is not None:
        x = np.concatenate([[0.0], x, [0.0]])
        x[0] = x[1] * 2 - x[2]
        x[-1] = x[-2] * 2 - x[-3]
    dx = np.diff(x)
    uniform = bool((dx == dx[0]).all())
    d = dict(uniform=uniform, edge=edge, den=2. * dx[0], lo=float(dx[0]), hi=float(dx[-1]), w=None)
    if edge == "fixed":
        d["lo"], d["hi"] = float(fill[0]), float(fill[1])
    if not uniform:
        dx1, dx2 = dx[0:-1], dx[1:]
        a = -(dx2) / (dx1 * (dx1 + dx2))
        b = (dx2 - dx1) / (dx1 * dx2)
        c = dx1 / (dx2 * (dx1 + dx2))
        w = np.zeros((3, n))
        sl = slice(1, n - 1) if edge is None else slice(0, n)
        w[0, sl], w[1, sl], w[2, sl] = a, b, c
        d["w"] = w
    return d


def flow_2d(S, ydiff, xdiff, comb, rows, swap=False, signs=(1.0, 1.0), deg2m=1.0, ctx=None):
    """Two flow components from ``S[..., ny, nx]`` (host float64) on the device: centred differences along the last
    two axes as described by ``axis_diff`` and the combination ``comb`` of include/xinv.h (XINV_FLOW_*)."""
    L = _lib.load()
    ctx = ctx or _lib.default_context()
    Sh = _host_f64(S, "S")
    shape = Sh.shape
    ny, nx = int(shape[-2]), int(shape[-1])
    batch = int(np.prod(shape[:-2], dtype=np.int64)) if len(shape) > 2 else 1
    rows = np.ascontiguousarray(rows, dtype=np.float64)
    o1, o2 = _lib.pinned_empty(shape), _lib.pinned_empty(shape)
    keep = []

    def axis(d):
        ax = _lib.XinvFlowAxis()
        ax.uniform, ax.edge, ax.den, ax.lo, ax.hi = int(d["uniform"]), _lib.EDGE_CODES[d["edge"]], d["den"], d["lo"], d["hi"]
        if d["w"] is not None:
            w = np.ascontiguousarray(d["w"], dtype=np.float64)
            keep.append(w)
            ax.w = w.ctypes.data
        return ax

    desc = _lib.XinvFlowDesc()
    desc.struct_size = C.sizeof(_lib.XinvFlowDesc)
    desc.comb, desc.swap, desc.nrows = int(comb), int(bool(swap)), int(rows.shape[0])
    desc.s1, desc.s2, desc.deg2m = float(signs[0]), float(signs[1]), float(deg2m)
    desc.y, desc.x = axis(ydiff), axis(xdiff)
    desc.rows = rows.ctypes.data
    opts = _lib.make_opts(mem_space=_lib.MEM_HOST)
    N = ny * nx
    with ctx.lock:
        for lo, hi in _batch_chunks(batch):
            o = 8 * lo * N
            _lib.check(L.xinv_flow2d(ctx.handle, C.c_void_p(o1.ctypes.data + o), C.c_void_p(o2.ctypes.data + o),
                                     C.c_void_p(Sh.ctypes.data + o), hi - lo, ny, nx, C.byref(desc), C.byref(opts)))
    del keep
    return o1, o2


# ---------------------------------------------------------------------------
# single-slice shims with the reference's numba signatures
# ---------------------------------------------------------------------------
def _store_flags(flags, fl):
    flags[0], flags[1], flags[2] = fl[0, 0], fl[0, 1], fl[0, 2]


def invert_standard_2D(S, A, B, C_, F, yc, xc, dely, delx, BCy, BCx, delxSqr,
                       ratioQtr, ratioSqr, optArg, undef, flags, mxLoop, tolerance, **kw):
    """Same positional signature as ``numbas.invert_standard_2D`` (numbas.py:216-219)."""
    if tuple(S.shape) != (yc, xc):
        raise ValueError(f"S.shape {tuple(S.shape)} != (yc, xc) = {(yc, xc)}")
    fl, _ = solve_standard_2D(S, A, B, C_, F, BCy, BCx, delxSqr, ratioQtr, ratioSqr, optArg,
                              undef, flags, mxLoop, tolerance, **kw)
    _store_flags(flags, fl)
    return S


def invert_general_2D(S, A, B, C_, D, E, F, G, yc, xc, dely, delx, BCy, BCx,
                      delxSqr, ratio, ratioQtr, ratioSqr, optArg, undef, flags,
                      mxLoop, tolerance, **kw):
    """Same positional signature as ``numbas.invert_general_2D`` (numbas.py:988-991)."""
    if tuple(S.shape) != (yc, xc):
        raise ValueError(f"S.shape {tuple(S.shape)} != (yc, xc) = {(yc, xc)}")
    fl, _ = solve_general_2D(S, A, B, C_, D, E, F, G, BCy, BCx, delx, delxSqr, ratio, ratioQtr,
                             ratioSqr, optArg, undef, flags, mxLoop, tolerance, **kw)
    _store_flags(flags, fl)
    return S


def invert_standard_3D(S, A, B, C_, F, zc, yc, xc, delz, dely, delx, BCz, BCy, BCx,
                       delxSqr, ratio2Sqr, ratio1Sqr, optArg, undef, flags, mxLoop,
                       tolerance, **kw):
    """Same positional signature as ``numbas.invert_standard_3D`` (numbas.py:16-19)."""
    if tuple(S.shape) != (zc, yc, xc):
        raise ValueError(f"S.shape {tuple(S.shape)} != (zc, yc, xc) = {(zc, yc, xc)}")
    fl, _ = solve_standard_3D(S, A, B, C_, F, BCz, BCy, BCx, delxSqr, ratio2Sqr, ratio1Sqr, optArg,
                              undef, flags, mxLoop, tolerance, **kw)
    _store_flags(flags, fl)
    return S


def invert_standard_2D_test(S, A, B, C_, D, E, F, yc, xc, dely, delx, BCy, BCx, delxSqr,
                            ratioQtr, ratioSqr, optArg, undef, flags, mxLoop, tolerance, **kw):
    """Same positional signature as ``numbas.invert_standard_2D_test`` (numbas.py:421-424)."""
    if tuple(S.shape) != (yc, xc):
        raise ValueError(f"S.shape {tuple(S.shape)} != (yc, xc) = {(yc, xc)}")
    fl, _ = solve_standard_2D_test(S, A, B, C_, D, E, F, BCy, BCx, delxSqr, ratioQtr, ratioSqr, optArg, undef, flags,
                                   mxLoop, tolerance, **kw)
    _store_flags(flags, fl)
    return S


def invert_general_3D(S, A, B, C_, D, E, F, G, H, zc, yc, xc, delz, dely, delx, BCz, BCy, BCx, delxSqr,
                      ratio2, ratio1, ratio2Sqr, ratio1Sqr, optArg, undef, flags, mxLoop, tolerance, **kw):
    """Same positional signature as ``numbas.invert_general_3D`` (numbas.py:746-749)."""
    if tuple(S.shape) != (zc, yc, xc):
        raise ValueError(f"S.shape {tuple(S.shape)} != (zc, yc, xc) = {(zc, yc, xc)}")
    fl, _ = solve_general_3D(S, A, B, C_, D, E, F, G, H, BCz, BCy, BCx, delx, delxSqr, ratio2, ratio1, ratio2Sqr, ratio1Sqr,
                             optArg, undef, flags, mxLoop, tolerance, **kw)
    _store_flags(flags, fl)
    return S


def invert_standard_1D(S, A, B, F, xc, delx, BCx, delxSqr, optArg, undef, flags, mxLoop, tolerance, **kw):
    """Same positional signature as ``numbas.invert_standard_1D`` (numbas.py:633-635)."""
    if tuple(S.shape) != (xc,):
        raise ValueError(f"S.shape {tuple(S.shape)} != (xc,) = {(xc,)}")
    fl, _ = solve_standard_1D(S, A, B, F, BCx, delxSqr, optArg, undef, flags, mxLoop, tolerance, **kw)
    _store_flags(flags, fl)
    return S


def invert_general_bih_2D(S, A, B, C_, D, E, F, G, H, I, J, yc, xc, dely, delx, BCy, BCx, delxSSr, delxTr, delxSqr,
                          ratio, ratioSSr, ratioQtr, ratioSqr, optArg, undef, flags, mxLoop, tolerance, **kw):
    """Same positional signature as ``numbas.invert_general_bih_2D`` (numbas.py:1205-1210)."""
    if tuple(S.shape) != (yc, xc):
        raise ValueError(f"S.shape {tuple(S.shape)} != (yc, xc) = {(yc, xc)}")
    fl, _ = solve_general_bih_2D(S, A, B, C_, D, E, F, G, H, I, J, BCy, BCx, delxSSr, delxTr, delxSqr, ratio, ratioSSr,
                                 ratioQtr, ratioSqr, optArg, undef, flags, mxLoop, tolerance, **kw)
    _store_flags(flags, fl)
    return S
