"""Load the UNMODIFIED reference kernels (``/root/reference/xinvert/numbas.py``)
by file path.  TEST INFRASTRUCTURE: used only to pin the C oracle and to
generate the fixtures under ``tests/golden/``.  ``/root/reference`` exists only
in the authoring container, never on the GPU box, so everything that calls this
must be skippable (``available()``).
"""
import importlib.util
import os

REF_NUMBAS = "/root/reference/xinvert/numbas.py"
# the same file, copied unmodified by oracle/build.py:build_ref() into the git-ignored oracle/_ref/
# (it travels to the GPU box with the snapshot; used there by bench.py's CPU legs only)
REF_COPY = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "numbas.py")
_mod = None


def _path():
    return REF_NUMBAS if os.path.exists(REF_NUMBAS) else (REF_COPY if os.path.exists(REF_COPY) else None)


def available():
    if _path() is None:
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
        spec = importlib.util.spec_from_file_location("ref_xinvert_numbas", _path())
        _mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_mod)
    return _mod
