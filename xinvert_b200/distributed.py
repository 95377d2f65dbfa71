"""Multi-GPU driver: independent slices sharded across ranks (one process per
GPU), no halo exchange, one scalar all-reduce of the active-slice count per
chunk of sweeps.

The reference has no multi-device path at all (core.py:129 is a serial loop);
slices are independent solves, so every rank keeps the exact per-slice
semantics (each slice stops on its own test) and the collective only decides
when *all* ranks may leave the loop together.

Host logic (partitioning, the termination protocol) is backend-agnostic and is
tested on CPU with gloo and a mock stepper; on the GPU box the scalar goes over
NCCL -- either the library's own communicator (``backend='xinv-nccl'``,
``include/xinv.h: xinv_nccl_*``) or ``torch.distributed``'s.
"""
import ctypes as C

import numpy as np

from . import _lib


def shard_bounds(n_items, world, rank):
    """Contiguous block partition of the flattened batch axis: the first
    ``n_items % world`` ranks get one extra slice."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, extra = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


class TorchAllReduce:
    """Sum-all-reduce of one int64 through torch.distributed (nccl or gloo)."""

    def __init__(self, group=None, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.device = device if device is not None else (
            torch.device("cuda", torch.cuda.current_device())
            if dist.get_backend(group) == "nccl" else torch.device("cpu"))

    def __call__(self, value):
        t = self.torch.tensor([int(value)], dtype=self.torch.int64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return int(t.item())


class XinvNcclAllReduce:
    """Sum-all-reduce through the library's own NCCL communicator.  The 128-byte
    unique id is created on rank 0 and shipped with torch.distributed."""

    def __init__(self, ctx, rank, world, group=None):
        import torch
        import torch.distributed as dist
        L = _lib.load()
        buf = (C.c_char * 128)()
        if rank == 0:
            _lib.check(L.xinv_nccl_unique_id(C.cast(buf, C.c_void_p)))
        obj = [bytes(buf.raw)]
        dist.broadcast_object_list(obj, src=0, group=group)
        idbuf = C.create_string_buffer(obj[0], 128)
        _lib.check(L.xinv_nccl_init(ctx.handle, C.cast(idbuf, C.c_void_p), int(rank), int(world)))
        self.ctx, self.L = ctx, L

    def __call__(self, value):
        out = C.c_int64(0)
        _lib.check(self.L.xinv_nccl_allreduce_active(self.ctx.handle, int(value), C.byref(out)))
        return out.value


def run_until_all_done(step, allreduce, sweeps_per_chunk=0, max_chunks=10 ** 9):
    """Termination protocol shared by all ranks.

    ``step(k)`` advances the local problem by up to ``k`` sweeps (0 = library
    default) and returns the number of local slices still active; ranks whose
    slices are all frozen keep calling it (a no-op) so that every rank issues
    the same number of collectives.  Returns the number of chunks executed."""
    chunks = 0
    while chunks < max_chunks:
        local = step(sweeps_per_chunk)
        chunks += 1
        if allreduce(local) == 0:
            break
    return chunks


class Stepper:
    """begin/step/end protocol of the C-ABI for one rank's shard."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.L = _lib.load()

    def step(self, sweeps=0):
        na = C.c_int64(0)
        _lib.check(self.L.xinv_step(self.ctx.handle, int(sweeps), C.byref(na)))
        return na.value

    def end(self):
        _lib.check(self.L.xinv_end(self.ctx.handle))


def solve_standard_2D_sharded(S, A, B, C_, F, BCy, BCx, delxSqr, ratioQtr, ratioSqr, optArg,
                              undef=-9.99e8, flags=(0.0, 1.0, 0.0), mxLoop=5000, tolerance=1e-8,
                              ctx=None, allreduce=None, sweeps_per_chunk=0, engine="auto", profile=False):
    """Solve this rank's shard ``S[local_batch, ny, nx]`` (host arrays or CUDA
    tensors) while staying in lock-step with the other ranks.  Arguments as
    ``solvers.solve_standard_2D``; returns ``(flags, stats, chunks)``."""
    from . import solvers
    ctx = ctx or _lib.default_context()
    B = solvers._zero_to_none(B)
    ops = solvers._Operands(S, [("A", A), ("B", B), ("C", C_), ("F", F)], 2)
    ny, nx = ops.core
    opts = _lib.make_opts(mem_space=_lib.MEM_DEVICE if ops.device else _lib.MEM_HOST, engine=engine,
                          coef_strides=ops.strides, profile=profile)
    fl = solvers._flags_array(flags, ops.batch)
    ptrs = [C.c_void_p(ops.S_ptr)] + [C.c_void_p(p) if p is not None else None for p in ops.ptrs]
    L = _lib.load()
    if ops.device:
        solvers._sync_torch_stream(S)           # operands produced on torch's stream (S.copy_ ...) are complete
    ctx.lock.acquire()                           # one begin..end sequence at a time on a ctx
    try:
        return _sharded_locked(L, ctx, ptrs, ops, ny, nx, BCy, BCx, delxSqr, ratioQtr, ratioSqr, optArg, undef, fl,
                               mxLoop, tolerance, opts, allreduce, sweeps_per_chunk)
    finally:
        ctx.lock.release()


def _sharded_locked(L, ctx, ptrs, ops, ny, nx, BCy, BCx, delxSqr, ratioQtr, ratioSqr, optArg, undef, fl, mxLoop,
                    tolerance, opts, allreduce, sweeps_per_chunk):
    _lib.check(L.xinv_std2d_begin(ctx.handle, *ptrs, ops.batch, ny, nx, _lib.BC_CODES[BCy], _lib.BC_CODES[BCx],
                                  float(delxSqr), float(ratioQtr), float(ratioSqr), float(optArg), float(undef),
                                  C.c_void_p(fl.ctypes.data), int(mxLoop), float(tolerance), C.byref(opts)))
    st = Stepper(ctx)
    try:
        chunks = run_until_all_done(st.step, allreduce or (lambda v: v), sweeps_per_chunk)
    finally:
        st.end()
    ops.finish()
    return fl, ctx.stats(), chunks
