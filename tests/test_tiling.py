"""
Out-of-core tiling with buffers, mirroring nd/tests/test_tiling.py:33-127 on the Dataset
stand-in with .npz tiles (NetCDF / xarray / dask are absent in this image).  CPU part: split / buffer / merge logic.
GPU part: a filter mapped over buffered tiles equals the filter on the whole cube -- the `buffer` IS the halo.
"""
import os

import numpy as np
import pytest

from nd_b200 import tiling
from nd_b200.dataset import generate_test_dataset, open_dataset, save_dataset

ny, nx, ntime = 20, 20, 10
ds = generate_test_dataset(dims={'y': ny, 'x': nx, 'time': ntime})
slices = dict(y=[slice(None, 10), slice(10, None)], x=[slice(None, 10), slice(10, None)],
              time=[slice(None, 5), slice(5, None)])
parts = [ds.isel(**sl) for sl in tiling.dict_product(slices)]
buffered_slices = dict(y=[slice(None, 12), slice(8, None)], x=[slice(None, 11), slice(9, None)],
                       time=[slice(None, 5), slice(5, None)])
buffered_parts = [ds.isel(**sl) for sl in tiling.dict_product(buffered_slices)]


def test_dataset_file_round_trip(tmp_path):
    p = str(tmp_path / 'a.npz')
    save_dataset(ds, p)
    back = open_dataset(p)
    assert back.equals(ds) and dict(back.attrs) == dict(ds.attrs)
    assert [back[v].dims for v in back.data_vars] == [ds[v].dims for v in ds.data_vars]


def test_auto_merge():
    """nd/tests/test_tiling.py:33-43"""
    assert ds.equals(tiling.auto_merge(parts))
    merged = tiling.auto_merge(buffered_parts)
    assert ds.equals(merged) and dict(merged.attrs) == dict(ds.attrs)
    with pytest.raises(ValueError):
        tiling.auto_merge([])


def test_debuffer_keeps_every_voxel_once():
    flat = tiling.debuffer(buffered_parts)
    assert sum(int(np.prod(d['C11'].shape)) for d in flat) == ny * nx * ntime
    grid = tiling.debuffer(buffered_parts, flat=False)
    assert grid.shape == (2, 2, 2)


@pytest.mark.parametrize('buffer', [0, 2, {'x': 3}])
@pytest.mark.parametrize('chunks', [{'time': 2}, {'x': 4}, {'y': 10, 'x': 10}, {'y': 100, 'x': 100},
                                    {'y': 100, 'x': 8, 'time': 3}])
def test_tile(tmp_path, chunks, buffer):
    """nd/tests/test_tiling.py:59-100"""
    tile_path = str(tmp_path / 'tiles')
    tiling.tile(ds, tile_path, chunks=chunks, buffer=buffer)
    buffer_dict = ({dim: buffer for dim in ds.dims} if isinstance(buffer, int)
                   else {dim: buffer.get(dim, 0) for dim in chunks})
    nchunks = np.prod([int(np.ceil(ds.sizes[dim] / n)) for dim, n in chunks.items()])
    tile_files = [os.path.join(tile_path, f) for f in os.listdir(tile_path)]
    assert len(tile_files) == nchunks
    for f in tile_files:
        t = open_dataset(f)
        assert dict(t.attrs) == dict(ds.attrs)
        for dim, val in chunks.items():
            assert t.sizes[dim] <= val + 2 * buffer_dict[dim]
    assert ds.equals(tiling.auto_merge(os.path.join(tile_path, '*' + tiling.EXT)))


def test_tile_skips_existing_files_and_rejects_a_file_path(tmp_path):
    tile_path = str(tmp_path / 'tiles')
    tiling.tile(ds, tile_path, chunks={'y': 10})
    files = sorted(os.listdir(tile_path))
    assert files == ['part.y_0_10.npz', 'part.y_10_20.npz']                 # reference naming scheme
    stamp = [os.path.getmtime(os.path.join(tile_path, f)) for f in files]
    tiling.tile(ds, tile_path, chunks={'y': 10})                            # resume: nothing is rewritten
    assert stamp == [os.path.getmtime(os.path.join(tile_path, f)) for f in files]
    assert not [f for f in os.listdir(tile_path) if f.endswith('.part')]
    with pytest.raises(ValueError):
        tiling.tile(ds, os.path.join(tile_path, files[0]), chunks={'y': 10})


@pytest.mark.parametrize('fn', [lambda x: x, lambda x: _scaled(x, 2)])
def test_map_over_tiles_cpu(tmp_path, fn):
    """nd/tests/test_tiling.py:113-127 (the functions that need no filter)."""
    tile_path = str(tmp_path / 'tiles')
    tiling.tile(ds, tile_path, chunks={'y': 10, 'x': 10}, buffer=0)
    mapped = tiling.map_over_tiles(os.path.join(tile_path, '*' + tiling.EXT), fn, path=str(tmp_path / 'out'))
    assert mapped.equals(fn(ds))
    files = tiling.map_over_tiles(os.path.join(tile_path, '*' + tiling.EXT), fn, suffix='_f', merge=False)
    assert len(files) == 4 and all(f.endswith('_f' + tiling.EXT) for f in files)


def _scaled(d, k):
    out = d.copy(deep=True)
    for v in out.data_vars:
        out[v].values[...] *= k
    return out


@pytest.mark.gpu
def test_map_boxcar_over_buffered_tiles_equals_whole(tmp_path):
    """nd/tests/test_tiling.py:116 -- BoxcarFilter(w=3) over tiles with buffer 1."""
    from nd_b200.filters import BoxcarFilter
    fn = BoxcarFilter(w=3, dims=('x', 'y')).apply
    tile_path = str(tmp_path / 'tiles')
    tiling.tile(ds, tile_path, chunks={'y': 10, 'x': 10}, buffer=1)
    mapped = tiling.map_over_tiles(os.path.join(tile_path, '*' + tiling.EXT), fn, path=str(tmp_path / 'out'))
    whole = fn(ds)
    for v in ds.data_vars:        # a buffer of 1 keeps the interior exact; the summation order is the same
        assert np.array_equal(mapped[v].values, whole[v].values)


@pytest.mark.gpu
def test_map_nlmeans_over_buffered_tiles_equals_whole(tmp_path):
    """The out-of-core NLM workflow: tile with buffer = NLMeansFilter._buffer = r + f, filter tile by tile on the
    GPU, debuffer, merge -- equals the filter on the whole cube."""
    from nd_b200.filters import NLMeansFilter
    big = generate_test_dataset(dims={'y': 64, 'x': 70, 'time': 6}, dtype=np.float32)
    flt = NLMeansFilter(dims=('y', 'x', 'time'), r=(3, 3, 1), f=1, sigma=0.5, h=1.0)
    buffer = {d: flt._buffer(d) for d in ('y', 'x')}
    assert buffer == {'y': 4, 'x': 4}
    tile_path = str(tmp_path / 'tiles')
    tiling.tile(big, tile_path, chunks={'y': 32, 'x': 35}, buffer=buffer)
    mapped = tiling.map_over_tiles(os.path.join(tile_path, '*' + tiling.EXT), flt.apply, path=str(tmp_path / 'out'))
    whole = flt.apply(big)
    for v in big.data_vars:
        assert mapped[v].values.shape == whole[v].values.shape
        assert np.abs(mapped[v].values - whole[v].values).max() <= 1e-6 * np.abs(whole[v].values).max()
