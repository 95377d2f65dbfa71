"""Load the UNMODIFIED reference kernels (``/root/reference/xinvert/numbas.py``)
by file path.  TEST INFRASTRUCTURE: used only to pin the C oracle and to
generate the fixtures under ``tests/golden/``.  ``/root/reference`` exists only
in the authoring container, never on the GPU box, so everything that calls this
must be skippable (``available()``).
"""
import importlib.util
import os

REF_NUMBAS = "/root/reference/xinvert/numbas.py"
_mod = None


def available():
    if not os.path.exists(REF_NUMBAS):
        return False
    try:
        import numba  # noqa: F401
    except Exception:
        return False
    return True


def ref_numbas():
    """The reference's ``xinvert.numbas`` module, imported stand-alone."""
    global _mod
    if _mod is None:
        spec = importlib.util.spec_from_file_location("ref_xinvert_numbas", REF_NUMBAS)
        _mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_mod)
    return _mod
