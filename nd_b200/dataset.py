"""
A minimal stand-in for the handful of `xarray.Dataset` features `Filter.apply` touches
(reference nd/filters.py:105-191): named dims, coords, attrs, data variables with `.dims` /
`.values`, deep copy, item assignment / deletion, `isel`, `equals`.

xarray is not installed in this image; when it is, real `xarray.Dataset`s go
through exactly the same code path in `nd_b200.filters` (which only uses this subset).
"""
from collections import OrderedDict

import numpy as np


class Variable:
    def __init__(self, dims, values):
        values = np.asarray(values)
        dims = tuple(dims)
        if values.ndim != len(dims):
            raise ValueError('dims %r do not match data of shape %r' % (dims, values.shape))
        self.dims = dims
        self.values = values

    @property
    def shape(self):
        return self.values.shape

    @property
    def dtype(self):
        return self.values.dtype

    @property
    def sizes(self):
        return OrderedDict(zip(self.dims, self.values.shape))

    def copy(self, deep=True, data=None):
        if data is not None:                       # xarray's `copy(data=...)`: same dims, new values
            return Variable(self.dims, data)
        return Variable(self.dims, self.values.copy() if deep else self.values)


class DataArray:
    """A named array with dims, coords and attrs -- the subset of `xarray.DataArray` the change detectors return
    (reference nd/change.py:69-75)."""

    def __init__(self, values, dims, coords=None, attrs=None, name=None):
        self.values = np.asarray(values)
        self.dims = tuple(dims)
        if self.values.ndim != len(self.dims):
            raise ValueError('dims %r do not match data of shape %r' % (self.dims, self.values.shape))
        self.coords = OrderedDict((k, np.asarray(v)) for k, v in (coords or {}).items())
        self.attrs = OrderedDict(attrs or {})
        self.name = name

    @property
    def shape(self):
        return self.values.shape

    @property
    def dtype(self):
        return self.values.dtype

    def isel(self, **indexers):
        idx = tuple(indexers.get(d, slice(None)) for d in self.dims)
        dims = tuple(d for d in self.dims if not isinstance(indexers.get(d, slice(None)), (int, np.integer)))
        coords = OrderedDict((k, (c[indexers[k]] if k in indexers else c)) for k, c in self.coords.items())
        return DataArray(self.values[idx], dims, coords, self.attrs, self.name)

    def sum(self, dim=None):
        if dim is None:
            return self.values.sum()
        axis = self.dims.index(dim)
        return DataArray(self.values.sum(axis=axis), self.dims[:axis] + self.dims[axis + 1:],
                         OrderedDict((k, c) for k, c in self.coords.items() if k != dim), self.attrs, self.name)

    def all(self):
        return bool(self.values.all())

    def __eq__(self, other):
        return DataArray(self.values == (other.values if isinstance(other, DataArray) else other), self.dims,
                         self.coords, self.attrs, self.name)


class Dataset:
    def __init__(self, data_vars=None, coords=None, attrs=None):
        self.data_vars = OrderedDict()
        self.coords = OrderedDict((k, np.asarray(v)) for k, v in (coords or {}).items())
        self.attrs = OrderedDict(attrs or {})
        for k, v in (data_vars or {}).items():
            self[k] = v

    # -- mapping protocol ----------------------------------------------------------------
    def __getitem__(self, name):
        if name in self.data_vars:
            return self.data_vars[name]
        if name in self.coords:
            return Variable((name,), self.coords[name])
        raise KeyError(name)

    def __setitem__(self, name, value):
        if isinstance(value, Variable):
            var = value
        else:
            dims, values = value
            var = Variable(dims, values)
        for d, n in zip(var.dims, var.shape):
            if d in self.sizes and self.sizes[d] != n:
                raise ValueError('conflicting sizes for dimension %r' % d)
        self.data_vars[name] = var

    def __delitem__(self, name):
        del self.data_vars[name]

    def __contains__(self, name):
        return name in self.data_vars or name in self.coords

    # -- dims ----------------------------------------------------------------------------
    @property
    def sizes(self):
        out = OrderedDict()
        for name, c in self.coords.items():
            if c.ndim == 1:
                out[name] = c.shape[0]
        for var in self.data_vars.values():
            for d, n in zip(var.dims, var.shape):
                out.setdefault(d, n)
        return out

    @property
    def dims(self):
        # xarray sorts Dataset.dims alphabetically (noted at reference nd/filters.py:126-127)
        return OrderedDict(sorted(self.sizes.items()))

    # -- the operations Filter.apply needs -------------------------------------------------
    def copy(self, deep=True):
        new = Dataset(coords={k: (v.copy() if deep else v) for k, v in self.coords.items()}, attrs=self.attrs)
        for k, v in self.data_vars.items():
            new.data_vars[k] = v.copy(deep=deep)
        return new

    def isel(self, **indexers):
        new = Dataset(attrs=self.attrs)
        for k, c in self.coords.items():
            new.coords[k] = c[indexers[k]] if k in indexers else c
            if np.ndim(new.coords[k]) == 0:
                new.coords[k] = np.asarray(new.coords[k])
        for k, v in self.data_vars.items():
            idx = tuple(indexers.get(d, slice(None)) for d in v.dims)
            dims = tuple(d for d in v.dims if not isinstance(indexers.get(d, slice(None)), (int, np.integer)))
            new.data_vars[k] = Variable(dims, v.values[idx])
        return new

    def equals(self, other):
        if set(self.data_vars) != set(other.data_vars):
            return False
        for k, v in self.data_vars.items():
            o = other.data_vars[k]
            if v.dims != o.dims or v.shape != o.shape:
                return False
            if not np.array_equal(v.values, o.values, equal_nan=True):
                return False
        for k, c in self.coords.items():
            if k not in other.coords or not np.array_equal(c, other.coords[k]):
                return False
        return True


def save_dataset(ds, path):
    """Write `ds` to ONE file (NumPy .npz: variables, coords, and a JSON header with dims / attrs).
    Stand-in for the reference's `to_netcdf` (nd/io.py): NetCDF libraries are absent in this image."""
    import json
    meta = {'vars': {k: list(v.dims) for k, v in ds.data_vars.items()}, 'coords': list(ds.coords),
            'attrs': {k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in ds.attrs.items()}}
    arrays = {'v__' + k: v.values for k, v in ds.data_vars.items()}
    arrays.update({'c__' + k: c for k, c in ds.coords.items()})
    with open(path, 'wb') as f:
        np.savez(f, __meta__=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), **arrays)


def open_dataset(path):
    """Inverse of `save_dataset` (stand-in for the reference's `open_netcdf`)."""
    import json
    with np.load(path, allow_pickle=False) as z:
        meta = json.loads(bytes(z['__meta__']).decode())
        ds = Dataset(coords=OrderedDict((k, z['c__' + k]) for k in meta['coords']), attrs=meta['attrs'])
        for k, dims in meta['vars'].items():
            ds[k] = (tuple(dims), z['v__' + k])
    return ds


def concat(datasets, dim):
    """Concatenate datasets along `dim` (variables and the coordinate of that dimension)."""
    first = datasets[0]
    coords = OrderedDict(first.coords)
    if dim in coords:
        coords[dim] = np.concatenate([d.coords[dim] for d in datasets])
    out = Dataset(coords=coords, attrs=first.attrs)
    for k, v in first.data_vars.items():
        if dim in v.dims:
            out[k] = (v.dims, np.concatenate([d[k].values for d in datasets], axis=v.dims.index(dim)))
        else:
            out[k] = (v.dims, v.values)
    return out


def generate_test_dataset(dims=None, var=('C11', 'C12__im', 'C12__re', 'C22'), mean=0, sigma=1,
                          random_seed=42, dtype=np.float64):
    """NumPy-only restatement of the reference fixture generator (nd/testing.py:34-70):
    seeded N(mean, sigma) variables over (y, x, time) with coords and attrs."""
    if dims is None:
        dims = OrderedDict([('y', 20), ('x', 20), ('time', 10)])
    dims = OrderedDict(dims)
    np.random.seed(random_seed)
    extent = (-10.0, 50.0, 0.0, 60.0)
    coords = OrderedDict()
    for name, size in dims.items():
        if name == 'y':
            coords[name] = np.linspace(extent[3], extent[1], size)
        elif name == 'x':
            coords[name] = np.linspace(extent[0], extent[2], size)
        elif name == 'time':
            coords[name] = (np.datetime64('2017-01-01') +
                            (np.arange(size) * (365 * 86400 // max(size - 1, 1))).astype('timedelta64[s]'))
        else:
            coords[name] = np.arange(size)
    ds = Dataset(coords=coords, attrs={'attr1': 1, 'attr2': 2, 'attr3': 3})
    if isinstance(mean, (int, float)):
        mean = [mean] * len(var)
    for v, m in zip(var, mean):
        ds[v] = (tuple(dims.keys()), np.random.normal(m, sigma, tuple(dims.values())).astype(dtype))
    return ds
