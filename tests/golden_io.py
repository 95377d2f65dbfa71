"""Loader of the committed fixtures under tests/golden/ (made by
tests/golden/make_golden.py from the unmodified reference kernels)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

_INT_KEYS = ("gc1", "gc2", "gc3")


def load(tag):
    """Return (case dict usable with tests.cases.run_*, dict of stored outputs)."""
    z = np.load(os.path.join(GOLDEN, tag + ".npz"))
    c, out, p = {}, {}, {}
    for k in z.files:
        if k.startswith("p_"):
            v = z[k].item()
            p[k[2:]] = int(v) if k[2:] in _INT_KEYS else float(v)
        elif k.startswith("S_") or k.startswith("fl_") or k in ("S", "fl"):
            out[k] = z[k]
        else:
            c[k] = np.ascontiguousarray(z[k])
    c["p"] = p
    return c, out
