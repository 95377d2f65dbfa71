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


if __name__ == "__main__":
    print(build(force=True))
