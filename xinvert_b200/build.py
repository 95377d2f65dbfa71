"""Build libxinv_b200.so (hand-written sm_100a CUDA + the C-ABI) in-tree.

    python -m xinvert_b200.build            # build if stale
    python -m xinvert_b200.build --force

nvcc cross-compiles without a GPU.  Flags that matter:
  -gencode arch=compute_100a,code=sm_100a   B200 only; no other targets, no PTX JIT
  -fmad=false     no FMA contraction: every +,-,*,/ is one IEEE binary64 op in
                  the reference's evaluation order (numba fastmath=False), which
                  is what makes bit-exact parity with the oracle possible.  The
                  kernels are HBM-bound, so this costs nothing measurable.
  -lineinfo       ncu source-page mapping
  -cudart static  the .so loads (dlopen) on machines without a GPU or libcuda
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libxinv_b200.so")
SOURCES = ["xinv_api.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
    "-Xcompiler", "-fPIC",
    "-shared", "-cudart", "static",
]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, "..", "include", "xinv.h"))
    deps.append(os.path.abspath(__file__))
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    """Compile the library if it is missing or older than its sources.  Safe when several processes
    (one rank per GPU) get here at once: an fcntl lock serialises them, the compiler writes to a
    temporary file and the finished library is renamed into place, so nobody dlopens a half-written
    file and only the first process compiles."""
    if not force and not _stale():
        return LIB
    import fcntl
    try:
        lockf = open(LIB + ".lock", "w")
    except OSError:                              # read-only install: nothing to build into
        if os.path.exists(LIB):
            return LIB
        raise
    try:
        fcntl.flock(lockf, fcntl.LOCK_EX)
        if not force and not _stale():           # another process built it while we waited
            return LIB
        tmp = f"{LIB}.tmp.{os.getpid()}"
        extra = os.environ.get("XINV_NVCC_EXTRA", "").split()      # e.g. -DX3_BARRIER_NOINLINE=1 for a synccheck run
        cmd = [NVCC, *NVCC_FLAGS, *extra, *(["-Xptxas", "-v"] if verbose else []),
               *[os.path.join(CSRC, s) for s in SOURCES], "-o", tmp, "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            if os.path.exists(tmp):
                os.unlink(tmp)
            raise RuntimeError("nvcc failed building libxinv_b200.so:\n" + r.stderr[-4000:])
        os.replace(tmp, LIB)
    finally:
        fcntl.flock(lockf, fcntl.LOCK_UN)
        lockf.close()
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
