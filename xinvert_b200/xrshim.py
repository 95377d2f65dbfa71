"""A tiny named-array stand-in for ``xarray.DataArray``.

xarray is not installed in this image, so the xarray-in /
xarray-out surface of ``apps.py`` is exercised with this class; when xarray *is*
importable, real ``xarray.DataArray`` objects are accepted and returned instead
(``wrap_like``).  Only what the hot-path facade needs is implemented: named
dims, 1-D coordinates, ``.values``, ``.rename``, coordinate lookup by name,
``isel`` and element-wise arithmetic between arrays of identical dims.
"""
import numpy as np

try:                                    # pragma: no cover - not available in this image
    import xarray as _xr
except Exception:                       # noqa: BLE001
    _xr = None


class DataArray:
    """``DataArray(values, dims=[...], coords={dim: 1-D array}, name=None)``."""

    def __init__(self, values, dims=None, coords=None, name=None):
        self.values = np.asarray(values)
        if dims is None:
            dims = [f"dim_{k}" for k in range(self.values.ndim)]
        if isinstance(dims, str):
            dims = [dims]
        self.dims = tuple(dims)
        if len(self.dims) != self.values.ndim:
            raise ValueError(f"{len(self.dims)} dims for a {self.values.ndim}-D array")
        self.coords = {}
        for d, n in zip(self.dims, self.values.shape):
            c = None if coords is None else coords.get(d)
            if c is None:
                c = np.arange(n)
            c = np.asarray(getattr(c, "values", c))
            if c.shape != (n,):
                raise ValueError(f"coordinate {d!r} has shape {c.shape}, expected ({n},)")
            self.coords[d] = c
        self.name = name

    # -- introspection -----------------------------------------------------
    @property
    def shape(self):
        return self.values.shape

    @property
    def ndim(self):
        return self.values.ndim

    @property
    def dtype(self):
        return self.values.dtype

    def __getitem__(self, key):
        if isinstance(key, str):
            return DataArray(self.coords[key], dims=[key], coords={key: self.coords[key]}, name=key)
        out = self.values[key]
        if np.ndim(out) == self.values.ndim:        # plain slicing keeps the dims
            idx = key if isinstance(key, tuple) else (key,)
            idx = idx + (slice(None),) * (self.ndim - len(idx))
            coords = {d: self.coords[d][i] for d, i in zip(self.dims, idx)}
            return DataArray(out, self.dims, coords, self.name)
        return out

    def __setitem__(self, key, value):
        self.values[key] = getattr(value, "values", value)

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.values, dtype=dtype)

    def __repr__(self):
        return f"<xrshim.DataArray {self.name!r} {dict(zip(self.dims, self.shape))}>"

    # -- the few xarray methods the facade and its tests use -----------------
    def rename(self, name):
        return DataArray(self.values, self.dims, self.coords, name)

    def copy(self):
        return DataArray(self.values.copy(), self.dims, dict(self.coords), self.name)

    def isel(self, **indexers):
        idx = tuple(indexers.get(d, slice(None)) for d in self.dims)
        keep = [d for d, i in zip(self.dims, idx) if not np.isscalar(i)]
        coords = {d: self.coords[d][i] for d, i in zip(self.dims, idx) if not np.isscalar(i)}
        return DataArray(self.values[idx], keep, coords, self.name)

    def sum(self):
        return self.values.sum()

    def max(self):
        return self.values.max()

    def min(self):
        return self.values.min()

    def all(self):
        return self.values.all()

    def _binary(self, other, op):
        if isinstance(other, DataArray):
            if other.dims != self.dims:
                raise ValueError("xrshim only combines arrays with identical dims")
            other = other.values
        return DataArray(op(self.values, other), self.dims, self.coords, self.name)

    def __add__(self, o): return self._binary(o, np.add)
    def __radd__(self, o): return self._binary(o, lambda a, b: b + a)
    def __sub__(self, o): return self._binary(o, np.subtract)
    def __rsub__(self, o): return self._binary(o, lambda a, b: b - a)
    def __mul__(self, o): return self._binary(o, np.multiply)
    def __rmul__(self, o): return self._binary(o, lambda a, b: b * a)
    def __truediv__(self, o): return self._binary(o, np.divide)
    def __pow__(self, o): return self._binary(o, np.power)
    def __neg__(self): return DataArray(-self.values, self.dims, self.coords, self.name)
    def __abs__(self): return DataArray(np.abs(self.values), self.dims, self.coords, self.name)
    def __le__(self, o): return self._binary(o, np.less_equal)
    def __lt__(self, o): return self._binary(o, np.less)
    def __ge__(self, o): return self._binary(o, np.greater_equal)
    def __gt__(self, o): return self._binary(o, np.greater)


def is_xarray(obj):
    return _xr is not None and isinstance(obj, _xr.DataArray)


def coord_values(obj, dim):
    """1-D coordinate values of ``dim`` for an xarray or xrshim DataArray."""
    return np.asarray(obj[dim].values)


def wrap_like(template, values, name=None):
    """New array of the template's kind (xarray if the template is xarray)."""
    dims = tuple(template.dims)
    coords = {d: coord_values(template, d) for d in dims}
    if is_xarray(template):             # pragma: no cover - xarray absent here
        return _xr.DataArray(values, dims=dims, coords=coords, name=name)
    return DataArray(values, dims, coords, name)
