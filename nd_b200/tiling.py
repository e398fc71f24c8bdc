"""
Out-of-core tiling with buffers -- the reference's nd/tiling.py (`tile` :18-106, `map_over_tiles` :109-179,
`sort_into_array` :210-240, `debuffer` :243-285, `auto_merge` :342-422) for cubes that do
not fit in host memory at once: cut a Dataset into tiles with `buffer` overlapping rows (for the NLM filter the
buffer is its halo `r + f`, `NLMeansFilter._buffer`), filter tile by tile on the GPU, strip the buffers, merge.

Differences from the reference, all forced by this image: tiles are NumPy `.npz` files written by
`nd_b200.dataset.save_dataset` instead of NetCDF (h5netcdf / netCDF4 / xarray are not installed), chunk sizes must
be given (there are no dask chunks to default to), and `map_over_tiles` runs the tiles one after the other on the
GPU instead of building a dask.delayed graph (every tile already fills the device).  Names, argument meaning, file
naming (`{prefix}.{dim}_{start}_{stop}...`), skip-existing / `.part` rename semantics and the buffer arithmetic
are the reference's.
"""
import glob
import itertools
import os
from collections import OrderedDict

import numpy as np

from .dataset import Dataset, concat, open_dataset, save_dataset

EXT = '.npz'


def dict_product(d):
    """Like itertools.product, but for a dict of lists (reference nd/utils.py `dict_product`)."""
    return (dict(zip(d, x)) for x in itertools.product(*d.values()))


def get_dims(ds):
    """The dimensions that carry a coordinate, in a fixed order (reference nd/utils.py `get_dims`)."""
    return tuple(d for d in ds.dims if d in ds.coords)


def tile(ds, path, prefix='part', chunks=None, buffer=0):
    """Split dataset into tiles and write them to `path` (reference nd/tiling.py:18-106).

    chunks : dict  chunk size for every dimension along which to split
    buffer : int or dict  number of overlapping pixels stored around each tile
    """
    if os.path.isfile(path):
        raise ValueError("`path` cannot be a file!")
    elif not os.path.isdir(path):
        os.makedirs(path)
    if isinstance(ds, str):
        ds = open_dataset(ds)
    if chunks is None:
        raise ValueError('`chunks` must be given (no dask chunks to default to)')

    # 1. Convert chunk sizes into slice objects.
    slices = OrderedDict()
    for dim, size in chunks.items():
        n = ds.sizes[dim]
        if isinstance(buffer, int):
            _buf = buffer
        elif isinstance(buffer, dict) and dim in buffer:
            _buf = buffer[dim]
        else:
            _buf = 0
        slices[dim] = []
        start = 0
        while start < n:
            length = min(int(size), n - start)
            slices[dim].append(slice(max(0, start - _buf), start + length + _buf))
            start += length

    def _write_tile(slice_dict):
        subset = ds.isel(**slice_dict)
        suffix = '.'.join('{}_{}_{}'.format(dim, s.start, s.stop) for dim, s in slice_dict.items())
        tile_path = os.path.join(path, '{}.{}{}'.format(prefix, suffix, EXT))
        if not os.path.isfile(tile_path):                      # skip existing files (resume)
            temp_tile_path = tile_path + '.part'
            save_dataset(subset, temp_tile_path)
            os.rename(temp_tile_path, tile_path)

    # 2. Then apply itertools to the slices.
    for slice_dict in dict_product(slices):
        _write_tile(slice_dict)


def map_over_tiles(files, fn, args=(), kwargs={}, path=None, suffix='', merge=True, overwrite=False):
    """Apply `fn` to each tile, write every result next to its input (or into `path`) and return the merged
    result (reference nd/tiling.py:109-179)."""
    if isinstance(files, str):
        files = sorted(glob.glob(files))
    if path is not None:
        os.makedirs(path, exist_ok=True)
    results = []
    for f in files:
        data = open_dataset(f)
        result = fn(data, *args, **kwargs)
        root, name = os.path.split(f)
        stem, ext = os.path.splitext(name)
        out_file = os.path.join(root if path is None else path, '{}{}{}'.format(stem, suffix, ext))
        if not overwrite and os.path.exists(out_file):
            out_file = '{}_new{}'.format(*os.path.splitext(out_file))
        save_dataset(result, out_file)
        results.append(out_file)
    return auto_merge(results) if merge else results


def sort_into_array(datasets, dims=None):
    """An object array laid out like the tile grid (reference nd/tiling.py:210-240).  `dims` is accepted and ignored
    exactly as in the reference, whose first statement overwrites it with the dims of the first dataset (:214)."""
    dims = get_dims(datasets[0])
    initials = {dim: np.unique([d.coords[dim][0] for d in datasets]) for dim in dims}
    grid = np.empty(tuple(len(initials[dim]) for dim in dims), dtype=object)

    def _idx(ds):
        result = []
        for dim in dims:
            vals = ds.coords[dim]
            order = 1 if (len(vals) < 2 or vals[-1] >= vals[0]) else -1
            result.append(int(np.argmax(initials[dim][::order] == vals[0])))
        return tuple(result)

    for d in datasets:
        grid[_idx(d)] = d
    return grid


def debuffer(datasets, flat=True):
    """Remove the buffer from tiled datasets (reference nd/tiling.py:243-285): neighbours along a dimension share
    `o` coordinates; the first of them keeps the leading o - o//2, the second drops its first o//2."""
    dims = get_dims(datasets[0])
    grid = sort_into_array(datasets)

    def _remove_buffer(data, dim):
        overlap = [len(np.intersect1d(a.coords[dim], b.coords[dim])) for a, b in zip(data[:-1], data[1:])]
        buf_start = [o // 2 for o in overlap]
        buf_stop = [-(o - b) if b > 0 else None for b, o in zip(buf_start, overlap)]
        out = np.empty(len(data), dtype=object)
        for i, (d, start, stop) in enumerate(zip(data, [None] + buf_start, buf_stop + [None])):
            out[i] = d.isel(**{dim: slice(start, stop)})
        return out

    for axis, dim in enumerate(dims):
        moved = np.moveaxis(grid, axis, -1)
        for idx in np.ndindex(*moved.shape[:-1]):
            moved[idx] = _remove_buffer(list(moved[idx]), dim)
    return list(grid.flatten()) if flat else grid


def _get_common_attrs(datasets):
    """All attributes that are the same in every dataset (reference nd/tiling.py:316-339)."""
    attrs, not_equal = {}, []
    for d in datasets:
        for key, val in d.attrs.items():
            if key not in attrs:
                attrs[key] = val
            elif not np.array_equal(val, attrs[key]):
                not_equal.append(key)
    return {k: v for k, v in attrs.items() if k not in not_equal}


def auto_merge(datasets, buffer=True):
    """Merge a tiled Dataset along all of its split dimensions (reference nd/tiling.py:342-422)."""
    if isinstance(datasets, str):
        datasets = sorted(glob.glob(datasets))
    if len(datasets) == 0:
        raise ValueError("No files found!")
    if isinstance(datasets[0], str):
        datasets = [open_dataset(p) for p in datasets]
    attrs = _get_common_attrs(datasets)
    dims = get_dims(datasets[0])
    grid = debuffer(datasets, flat=False) if buffer else sort_into_array(datasets)
    # concatenate the grid, last dimension first
    for axis in reversed(range(grid.ndim)):
        dim = dims[axis]
        merged = np.empty(grid.shape[:-1], dtype=object)
        for idx in np.ndindex(*grid.shape[:-1]):
            row = list(grid[idx])
            merged[idx] = row[0] if len(row) == 1 else concat(row, dim)
        grid = merged
    result = grid[()]
    result.attrs = OrderedDict(attrs)
    return result
