"""A minimal stand-in for the parts of xarray that the reference's `Model.to_xr`, `diagnostics.band` and
`diagnostics.compare_ebal` use (ref model.py:338-447; diagnostics.py:19-108, 476-530).  xarray is not installed in this
image; with this module registered as `sys.modules["xarray"]` those reference functions run UNMODIFIED and produce the
fixtures `ref_ebal.npz` / `ref_to_xr.json` (make_golden_ebal.py), and the product's `to_xr` / `sensitivity_to_xr` can be
materialised in the tests.  TEST INFRASTRUCTURE ONLY: named dimensions, broadcasting by dimension name, `sum(dim)`,
positional indexing along the first dimension, attrs.  Nothing else of xarray is imitated."""
import numpy as np


def _dims(d):
    if d is None:
        return ()
    return (d,) if isinstance(d, str) else tuple(d)


class DataArray:
    def __init__(self, data=None, dims=None, attrs=None, name=None, coords=None):
        self.values = np.asarray(data)
        self.dims = _dims(dims)
        if self.values.ndim != len(self.dims):
            raise ValueError(f"{name}: data has {self.values.ndim} dimensions, dims = {self.dims}")
        self.attrs = dict(attrs or {})
        self.name = name
        self._ds = None  # owning Dataset: coordinate look-up by attribute (da.wl)

    # -- arithmetic with broadcasting by dimension name (attrs dropped, like xarray's default)
    def _binary(self, other, op):
        if not isinstance(other, DataArray):
            return DataArray(op(self.values, other), self.dims)
        dims = list(self.dims) + [d for d in other.dims if d not in self.dims]

        def expand(a):
            src = [dims.index(d) for d in a.dims]
            v = a.values.reshape(a.values.shape + (1,) * (len(dims) - a.values.ndim))
            order = list(src) + [i for i in range(len(dims)) if i not in src]
            return np.transpose(v, np.argsort(order))

        out = DataArray(op(expand(self), expand(other)), dims)
        out._ds = self._ds or other._ds
        return out

    def __mul__(self, o):
        return self._binary(o, np.multiply)

    __rmul__ = __mul__

    def __add__(self, o):
        return self._binary(o, np.add)

    def __sub__(self, o):
        return self._binary(o, np.subtract)

    def __truediv__(self, o):
        return self._binary(o, np.divide)

    def __rtruediv__(self, o):
        return DataArray(np.divide(o, self.values), self.dims)

    def __neg__(self):
        return DataArray(-self.values, self.dims)

    def __array__(self, dtype=None, copy=None):
        return self.values if dtype is None else self.values.astype(dtype)

    def sum(self, dim=None):
        if dim is None:
            return DataArray(self.values.sum(), ())
        ax = self.dims.index(dim)
        out = DataArray(self.values.sum(axis=ax), self.dims[:ax] + self.dims[ax + 1:])
        out._ds = self._ds
        return out

    def __getitem__(self, i):
        return DataArray(self.values[i], self.dims[1:] if np.ndim(i) == 0 else self.dims)

    def __getattr__(self, k):  # coordinate of the owning dataset, e.g. `da.wl`
        ds = self.__dict__.get("_ds")
        if ds is not None and k in ds._vars:
            return ds._vars[k]
        raise AttributeError(k)

    @property
    def shape(self):
        return self.values.shape

    def copy(self):
        out = DataArray(self.values.copy(), self.dims, self.attrs, self.name)
        out._ds = self._ds
        return out


class Dataset:
    def __init__(self, data_vars=None, coords=None, attrs=None):
        self._vars = {}
        self.coord_names = []
        for k, v in (coords or {}).items():
            self[k] = v
            self.coord_names.append(k)
        for k, v in (data_vars or {}).items():
            self[k] = v
        self.attrs = dict(attrs or {})

    def __setitem__(self, k, v):
        if isinstance(v, tuple):
            v = DataArray(v[1], v[0], v[2] if len(v) > 2 else None, name=k)
        elif not isinstance(v, DataArray):
            v = DataArray(v, (), name=k)
        else:
            v = DataArray(v.values, v.dims, v.attrs, name=k)
        v._ds = self
        self._vars[k] = v

    def __getitem__(self, k):
        return self._vars[k]

    def __contains__(self, k):
        return k in self._vars

    def __getattr__(self, k):
        v = self.__dict__.get("_vars", {})
        if k in v:
            return v[k]
        raise AttributeError(k)

    @property
    def variables(self):
        return dict(self._vars)

    @property
    def data_vars(self):
        return {k: v for k, v in self._vars.items() if k not in self.coord_names}

    @property
    def coords(self):
        return {k: self._vars[k] for k in self.coord_names}

    @property
    def dims(self):
        out = {}
        for v in self._vars.values():
            out.update(dict(zip(v.dims, v.values.shape)))
        return out

    def copy(self):
        out = Dataset(attrs=self.attrs)
        out.coord_names = list(self.coord_names)
        for k, v in self._vars.items():
            out._vars[k] = v.copy()
            out._vars[k]._ds = out
        return out

    def drop_vars(self, names):
        out = self.copy()
        for n in names:
            out._vars.pop(n, None)
            if n in out.coord_names:
                out.coord_names.remove(n)
        return out


def install():
    """Register this module as `xarray` (only when the real one is absent)."""
    import sys

    try:
        import xarray  # noqa: F401

        return sys.modules["xarray"]
    except ImportError:
        sys.modules["xarray"] = sys.modules[__name__]
        return sys.modules[__name__]
