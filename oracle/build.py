"""Build the C oracle (oracle/sor_oracle.c -> oracle/libsor_oracle.so).

TEST INFRASTRUCTURE: see the header of sor_oracle.c.  gcc only; no FMA
contraction and no fast-math so each operation is one IEEE binary64 op in the
order the reference (numba, fastmath=False) evaluates it.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "sor_oracle.c")
LIB = os.path.join(HERE, "libsor_oracle.so")

CFLAGS = ["-O2", "-ffp-contract=off", "-fno-fast-math", "-fno-unsafe-math-optimizations",
          "-shared", "-fPIC", "-Wall", "-Wextra"]


def build(force=False):
    if (not force and os.path.exists(LIB)
            and os.path.getmtime(LIB) >= os.path.getmtime(SRC)):
        return LIB
    cmd = ["gcc", *CFLAGS, SRC, "-o", LIB, "-lm"]
    subprocess.run(cmd, check=True)
    return LIB


REF_SRC = "/root/reference/xinvert/numbas.py"
REF_DIR = os.path.join(HERE, "_ref")
REF_COPY = os.path.join(REF_DIR, "numbas.py")


def build_ref():
    """oracle/_ref/: the reference's own kernel file, taken UNMODIFIED from where it lies under
    /root/reference (present in the authoring container only).  The directory is git-ignored -- no
    reference source enters the history -- but it travels to the GPU box with the snapshot, so that
    ``bench.py --impl reference`` and the cpu_baseline leg can time the reference itself (numba is in
    the image) instead of the C port.  Returns the path, or None when the reference is not here."""
    if not os.path.exists(REF_SRC):
        return REF_COPY if os.path.exists(REF_COPY) else None
    os.makedirs(REF_DIR, exist_ok=True)
    with open(REF_SRC, "rb") as f:
        data = f.read()
    if not os.path.exists(REF_COPY) or open(REF_COPY, "rb").read() != data:
        with open(REF_COPY, "wb") as f:
            f.write(data)
    return REF_COPY


if __name__ == "__main__":
    print(build(force=True))
    print(build_ref())
