"""
CPU tests of the host side (run with -m "not gpu"): the C-ABI library loads and exports every symbol
include/ndnlm.h declares, plans are geometry-only and can be created without a GPU, the Python mirror
keeps the reference interface (names, argument meaning, errors), and NOTHING computes on the CPU.
"""
import inspect
import os
import re

import numpy as np
import pytest

from nd_b200 import _lib, device
from nd_b200.dataset import Dataset, generate_test_dataset
from nd_b200.filters import Filter, NLMeansFilter, nlmeans
from nd_b200._filters import _pixelwise_nlmeans_3d, find_weight
from nd_b200.shard import ShardPlan

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- C ABI -------------------------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "ndnlm.h")).read()
    declared = set(re.findall(r"\b(ndnlm_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS)
    L = _lib.lib()
    for sym in declared:
        assert hasattr(L, sym), sym
    assert b"sm_100a" in L.ndnlm_version()


def test_no_oracle_import_in_product():
    """The product path must never route through the oracle (or any CPU fallback)."""
    for root, _, files in os.walk(os.path.join(ROOT, "nd_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, fn)).read()
                assert "import oracle" not in text and "from oracle" not in text, fn


def test_plan_roles_and_tiles_cfg3():
    p = device.Plan((4096, 4096, 32, 4), (5, 5, 2), (1, 1, 1), 0.25, 0.5)
    d = p.describe()
    assert p.is_tiled and "nlm_tiled" in d["kernel"]
    assert d["role_axis_WRX"] == [0, 2, 1]            # W = y, R = time, X = x
    assert d["pad_WRX"] == [6, 3, 6]
    assert d["n_offsets"] == 604
    assert d["smem_bytes"] <= 232448
    assert d["box_WRX"][1] % 2 == 1                   # odd R pitch (bank-conflict-free LDS.128)
    assert d["padded_bytes"] == (4096 + 12) * (4096 + 12) * (32 + 6) * 16


@pytest.mark.parametrize("shape,r,f,n_eff,F", [
    ((1, 206, 500, 4), (0, 3, 3), (0, 1, 1), -1, 1455),            # cfg1  
    ((1024, 1024, 24, 4), (5, 5, 1), (1, 1, 1), -1, 11599),        # cfg2
    ((4096, 4096, 32, 4), (5, 5, 2), (1, 1, 1), -1, 19343),        # cfg3
    ((16384, 16384, 64, 4), (7, 7, 2), (2, 2, 2), -1, 35983),      # cfg4
    ((32768, 32768, 128, 6), (10, 10, 3), (2, 2, 2), -1, 129633),  # cfg5
    ((4096, 4096, 32, 4), (5, 5, 2), (1, 1, 1), 50, 19343 + 2 * 604),
])
def test_algorithmic_flops_match_baseline_table(shape, r, f, n_eff, F):
    p = device.Plan(shape, r, f, 0.25, 0.5, n_eff)
    assert p.flops_per_voxel == F
    assert p.voxels == shape[0] * shape[1] * shape[2]


def test_plan_kernel_selection():
    # float64 data: the float64 instantiation of the tiled kernel where one exists (32-byte voxels), else generic
    pd = device.Plan((40, 50, 30, 4), (2, 2, 1), (1, 1, 1), 1, 1, dtype=np.float64)
    assert pd.is_tiled and "double" in pd.kernel_name and pd.kernel_request == "tiled64"
    assert pd.padded_bytes == (40 + 6) * (50 + 6) * (30 + 4) * 32 and pd.out_bytes == 40 * 50 * 30 * 32
    assert "double" in device.Plan((40, 50, 30, 4), (2, 2, 1), (2, 2, 2), 1, 1, dtype=np.float64).kernel_name   # f = 2 too
    assert device.Plan((40, 50, 30, 4), (2, 2, 1), (3, 3, 3), 1, 1, dtype=np.float64).kernel_name == "nlm_generic<double>"
    assert device.Plan((40, 50, 30, 6), (2, 2, 1), (1, 1, 1), 1, 1, dtype=np.float64).kernel_name == "nlm_generic<double>"
    assert device.Plan((40, 50, 30, 4), (2, 2, 1), (1, 1, 1), 1, 1, dtype=np.float64, kernel="tiled64").is_tiled
    with pytest.raises(ValueError):
        device.Plan((40, 50, 30, 4), (2, 2, 1), (1, 1, 1), 1, 1, dtype=np.float32, kernel="tiled64")
    pb = device.Plan((40, 50, 30, 4), (2, 2, 1), (1, 1, 1), 1, 1, semantics="reference_compiled")
    assert "boxmean" in pb.kernel_name and pb.scratch_bytes == (40 + 4) * 50 * 30 * 16 and pb.kernel_request == "tiled"
    assert device.Plan((40, 50, 30, 4), (2, 2, 1), (1, 1, 1), 1, 1, n_eff=5, semantics="reference_compiled").kernel_name \
        == "nlm_generic<float>[zero_dist]"
    assert "zero_dist" in device.Plan((40, 50, 30, 4), (2, 2, 1), (1, 1, 1), 1, 1, semantics="reference_compiled").kernel_name
    assert device.Plan((40, 50, 30, 4), (2, 2, 1), (0, 0, 0), 1, 1, semantics="reference_compiled").is_tiled
    assert device.Plan((1, 206, 500, 4), (0, 3, 3), (0, 1, 1), 1, 1).is_tiled
    assert not device.Plan((40, 50, 30, 4), (2, 2, 1), (1, 1, 1), 1, 1, kernel="generic").is_tiled
    # float64 data on the fp32 tiled kernel only on explicit request
    p64 = device.Plan((40, 50, 30, 4), (2, 2, 1), (1, 1, 1), 1, 1, dtype=np.float64, kernel="tiled")
    assert p64.is_tiled and p64.padded_bytes == (40 + 6) * (50 + 6) * (30 + 4) * 16
    with pytest.raises(ValueError):                                  # V > 8 has no tiled instantiation
        device.Plan((40, 50, 30, 9), (2, 2, 1), (1, 1, 1), 1, 1, kernel="tiled")
    # ... or globally with ND_NLM_FLOAT64_COMPUTE=float32 (configurations without an instantiation stay generic)
    os.environ["ND_NLM_FLOAT64_COMPUTE"] = "float32"
    try:
        assert device.Plan((40, 50, 30, 4), (2, 2, 1), (1, 1, 1), 1, 1, dtype=np.float64).is_tiled
        assert not device.Plan((40, 50, 30, 9), (2, 2, 1), (1, 1, 1), 1, 1, dtype=np.float64).is_tiled
        assert device.Plan((40, 50, 30, 4), (2, 2, 1), (1, 1, 1), 1, 1, dtype=np.float64, kernel="generic").kernel_name == "nlm_generic<double>"
    finally:
        del os.environ["ND_NLM_FLOAT64_COMPUTE"]


def test_double_duty_halo_warp_instantiations_are_chosen_for_f2(monkeypatch):
    """instances_g6.inc (ndnlm.cu): with four patch-halo rows per CTA tile (f_W = 2) the halo rows are served two per
    warp -- 10 valid rows of 12 warps instead of 8 -- by default; with two (f_W = 1) only on request."""
    monkeypatch.delenv("NDNLM_DH", raising=False)
    cfg4 = device.Plan((256, 16384, 64, 4), (7, 7, 2), (2, 2, 2), 0.25, 0.5)
    assert "(dh)" in cfg4.kernel_name and list(cfg4.info.tile)[0] == 10 and cfg4.info.smem_bytes <= 232448
    cfg3 = device.Plan((4096, 4096, 32, 4), (5, 5, 2), (1, 1, 1), 0.25, 0.5)
    assert "(dh)" not in cfg3.kernel_name and list(cfg3.info.tile)[0] == 14
    monkeypatch.setenv("NDNLM_DH", "0")
    plain = device.Plan((256, 16384, 64, 4), (7, 7, 2), (2, 2, 2), 0.25, 0.5)
    assert "(dh)" not in plain.kernel_name and list(plain.info.tile)[0] == 8
    assert plain.roles == cfg4.roles and plain.padded_bytes == cfg4.padded_bytes     # same staged layout
    monkeypatch.setenv("NDNLM_DH", "1")
    p = device.Plan((4096, 4096, 32, 4), (5, 5, 2), (1, 1, 1), 0.25, 0.5)
    assert "(dh)" in p.kernel_name and list(p.info.tile)[0] == 15 and p.info.smem_bytes <= 232448
    # float64 data: the default for f = 1 and f = 2 (7 / 6 valid rows of 8 warps instead of 6 / 4)
    monkeypatch.delenv("NDNLM_DH", raising=False)
    for f, rows in (((1, 1, 1), 7), ((2, 2, 2), 6)):
        p64 = device.Plan((210, 2048, 32, 4), (5, 5, 2), f, 0.25, 0.5, dtype=np.float64)
        assert "double" in p64.kernel_name and "(dh)" in p64.kernel_name and list(p64.info.tile)[0] == rows
    # configurations without such an instantiation are unaffected (2-D: no W patch axis; V = 6)
    assert "(dh)" not in device.Plan((1, 206, 500, 4), (0, 3, 3), (0, 1, 1), 1, 1).kernel_name
    assert "(dh)" not in device.Plan((64, 256, 128, 6), (10, 10, 3), (2, 2, 2), 1, 1).kernel_name


def test_output_is_native_rule():
    """ndnlm_output_is_native: a caller array that already has the internal output layout ([W][X][R] vectors of the 4
    variables) is written by the kernels directly (no unstage pass, no second copy of the cube)."""
    from nd_b200 import _lib
    L = _lib.lib()
    def native(plan, shape, order):
        strides, acc = [0] * 4, 1
        for ax in reversed(order):
            strides[ax] = acc
            acc *= shape[ax]
        return bool(L.ndnlm_output_is_native(plan._h, _lib.i64(strides)))
    shape = (64, 80, 16, 4)
    p = device.Plan(shape, (3, 3, 1), (1, 1, 1), 1, 1)
    assert p.roles == (0, 2, 1)                                   # W = axis 0, R = axis 2, X = axis 1
    assert native(p, shape, (0, 1, 2, 3))                         # C-ordered (y, x, time, V)
    assert not native(p, shape, (3, 0, 1, 2))                     # variable-major (what Filter.apply hands over)
    assert not native(p, shape, (0, 2, 1, 3))                     # time before x
    assert not native(device.Plan((64, 80, 16, 3), (3, 3, 1), (1, 1, 1), 1, 1), (64, 80, 16, 3), (0, 1, 2, 3))   # V != 4
    assert not native(device.Plan((64, 80, 16, 8), (3, 3, 1), (1, 1, 1), 1, 1), (64, 80, 16, 8), (0, 1, 2, 3))   # two groups
    assert native(device.Plan(shape, (3, 3, 1), (1, 1, 1), 1, 1, dtype=np.float64), shape, (0, 1, 2, 3))         # float64 kernel
    assert not native(device.Plan(shape, (3, 3, 1), (1, 1, 1), 1, 1, dtype=np.float64, kernel="tiled"), shape, (0, 1, 2, 3))
    assert not native(device.Plan(shape, (3, 3, 1), (1, 1, 1), 1, 1, kernel="generic"), shape, (0, 1, 2, 3))
    assert native(device.Plan(shape, (3, 3, 1), (1, 1, 1), 1, 1, semantics="reference_compiled"), shape, (0, 1, 2, 3))  # box mean
    # 2-D image behind a singleton axis: the stride of an extent-1 axis is irrelevant
    p2 = device.Plan((1, 206, 500, 4), (0, 3, 3), (0, 1, 1), 1, 1)
    ok = [native(p2, (1, 206, 500, 4), o) for o in ((0, 1, 2, 3), (0, 2, 1, 3))]
    assert ok.count(True) == 1                                    # exactly the order that matches its X / R roles


def test_staging_kernel_emulated_on_the_host(tmp_path):
    """tools/emu_stage.cu runs every thread of the row-blocked staging kernel on the CPU and compares the staged cube
    with the rule of the one-thread-per-element kernel it replaced (8960 layouts / edge modes / V / dtypes)."""
    import shutil, subprocess
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc is not on PATH")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "emu_stage")
    subprocess.run(["nvcc", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-w", "-o", exe,
                    os.path.join(root, "tools", "emu_stage.cu")], check=True, capture_output=True, timeout=600)
    p = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and " 0 mismatching elements" in p.stdout, p.stdout + p.stderr


def test_plan_errors_map_to_reference_exceptions():
    with pytest.raises(TypeError):                                   # fused `floating` dispatch failure
        device.Plan((4, 5, 6, 2), (1, 1, 1), (0, 0, 0), 1, 1, dtype=np.int32)
    with pytest.raises(ValueError, match="reflection"):              # r+f > N-1 is UB in the reference
        device.Plan((4, 5, 6, 2), (3, 1, 1), (1, 1, 1), 1, 1)
    with pytest.raises(ValueError):
        device.Plan((4, 5, 6), (1, 1, 1), (0, 0, 0), 1, 1)
    with pytest.raises(ValueError):
        device.Plan((4, 5, 6, 2), (1, 1, 1), (0, 0, 0), 1, 1, semantics="bogus")
    with pytest.raises(ValueError):
        device.Plan((4, 5, 6, 2), (1, 1, 1), (0, 0, 0), 1, 0.0)


def test_halo_message_size():
    p = device.Plan((64, 80, 16, 4), (3, 3, 1), (1, 1, 1), 1, 1)
    assert p.halo_bytes(0) == 4 * (80 + 8) * (16 + 4) * 16           # (r+f) rows x padded X x padded T x float4
    assert device.Plan((64, 80, 16, 4), (0, 3, 1), (0, 1, 1), 1, 1).halo_bytes(0) == 0


# ---- the reference's Python interface ------------------------------------------------------------
def test_filter_signature_is_the_reference_plugin_boundary():
    # nd/tests/test_filters_common.py:37-41
    assert list(inspect.signature(NLMeansFilter._filter).parameters) == ['self', 'arr', 'axes', 'output']
    assert issubclass(NLMeansFilter, Filter)
    assert NLMeansFilter.per_variable is False and NLMeansFilter.supports_complex is False
    params = inspect.signature(NLMeansFilter.apply).parameters
    assert list(params)[:3] == ['self', 'ds', 'inplace'] and 'njobs' in params
    sig = inspect.signature(NLMeansFilter.__init__).parameters
    assert [sig[k].default for k in ('dims', 'r', 'sigma', 'h', 'f', 'n_eff')] == [('y', 'x'), 1, 1, 1, 1, -1]
    assert list(inspect.signature(_pixelwise_nlmeans_3d).parameters)[:7] == ['arr', 'output', 'r', 'f', 'sigma', 'h', 'n_eff']
    assert 'dims' in inspect.signature(nlmeans).parameters and 'ds' in inspect.signature(nlmeans).parameters


def test_constructor_matches_reference():
    # nd/filters.py:414-422: scalar r is broadcast; f_i = f if r_i > 0 else 0; uint32 arrays
    flt = NLMeansFilter(dims=('y', 'x', 'time'), r=(2, 3, 0), f=2)
    assert flt.r.dtype == np.uint32 and flt.f.dtype == np.uint32
    assert list(flt.r) == [2, 3, 0] and list(flt.f) == [2, 2, 0]
    assert list(NLMeansFilter(dims=('y', 'x'), r=3).r) == [3, 3]
    assert flt._buffer('y') == 4 and flt._buffer('x') == 5 and flt._buffer('time') == 0 and flt._buffer('band') == 0


def test_parallel_dimension_rule():
    ds = generate_test_dataset(dims={'y': 20, 'x': 30, 'time': 10})
    assert NLMeansFilter(dims=('y', 'x'))._parallel_dimension(ds) == 'time'      # largest non-filter dim
    assert NLMeansFilter(dims=('y', 'x', 'time'))._parallel_dimension(ds) == 'x'  # else largest dim


def test_zero_radius_and_empty_dims_are_exact_identity():
    # nd/tests/test_nlmeans_filter.py:17-25 -- exact, float64 in -> float64 out, no device needed
    ds = generate_test_dataset(dims={'y': 20, 'x': 20, 'time': 10})
    out = NLMeansFilter(dims=('y', 'x'), r=0, f=1, sigma=1, h=1).apply(ds)
    assert isinstance(out, Dataset) and ds.equals(out) and out is not ds
    assert out.attrs == ds.attrs and list(out.coords) == list(ds.coords)
    assert ds.equals(NLMeansFilter(dims=(), r=1, f=1, sigma=1, h=1).apply(ds))
    assert ds.equals(nlmeans(ds, dims=('y', 'x'), r=0))
    with pytest.raises(NotImplementedError):
        NLMeansFilter(dims=('y', 'x'), r=0).apply(ds, inplace=True)


def test_compute_path_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    a = np.zeros((4, 5, 6, 2), np.float32)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _pixelwise_nlmeans_3d(a, np.empty_like(a), np.array([1, 1, 1], np.uint32), np.array([0, 0, 0], np.uint32), 1.0, 1.0)
    ds = generate_test_dataset(dims={'y': 8, 'x': 8, 'time': 3})
    with pytest.raises(RuntimeError):
        NLMeansFilter(dims=('y', 'x'), r=1).apply(ds)


def test_entry_point_argument_errors():
    a = np.zeros((4, 5, 6, 2), np.float32)
    u = lambda *v: np.array(v, np.uint32)
    with pytest.raises(TypeError):
        _pixelwise_nlmeans_3d(a.astype(np.int32), np.empty_like(a, dtype=np.int32), u(1, 1, 1), u(0, 0, 0), 1.0, 1.0)
    with pytest.raises(TypeError):
        _pixelwise_nlmeans_3d(a, np.empty_like(a, dtype=np.float64), u(1, 1, 1), u(0, 0, 0), 1.0, 1.0)
    with pytest.raises(ValueError, match="Buffer dtype mismatch"):
        _pixelwise_nlmeans_3d(a, np.empty_like(a), np.array([1, 1, 1], np.int64), u(0, 0, 0), 1.0, 1.0)
    with pytest.raises(ValueError):
        _pixelwise_nlmeans_3d(a[0], np.empty_like(a[0]), u(1, 1, 1), u(0, 0, 0), 1.0, 1.0)


def test_find_weight_closed_form():
    # nd/_filters.pyx:297-314
    S, Q, n = 26.0, 26.0, 5.0
    w = find_weight(S, Q, n)
    assert abs((S + w) ** 2 / (Q + w * w) - n) < 1e-12
    with pytest.raises(ValueError, match="No solution"):
        find_weight(1.0, 1.0, 20.0)


# ---- shard plan == xr_split chunking (nd/utils.py:305-310) ---------------------------------------
@pytest.mark.parametrize("n,chunks,halo", [(20, 2, 2), (21, 2, 3), (4096, 8, 6), (10, 3, 1), (7, 8, 0)])
def test_shard_plan_matches_xr_split(n, chunks, halo):
    sp = ShardPlan(n, chunks, halo)
    chunksize = int(np.ceil(n / chunks))
    expected = [(i * chunksize, min((i + 1) * chunksize, n)) for i in range(chunks) if i * chunksize < n]
    assert sp.ranges == expected
    assert sp.ranges[0][0] == 0 and sp.ranges[-1][1] == n
    for i, (lo, hi) in enumerate(sp.ranges):
        assert sp.buffered_range(i) == (max(lo - halo, 0), min(hi + halo, n))
        assert sp.edges(i) == ('reflect' if i == 0 else 'halo', 'reflect' if i == sp.nshards - 1 else 'halo')


@pytest.mark.parametrize("n,chunks,halo,expect", [(16, 8, 6, 2), (13, 8, 6, 1), (40, 8, 4, 8), (39, 8, 4, 5), (7, 3, 6, 1)])
def test_shard_plan_degrades_to_fewer_shards_when_they_cannot_carry_the_halo(n, chunks, halo, expect):
    """njobs larger than the cube can carry: fewer shards, like `xr_split` handing out fewer chunks, never an
    opaque radius error later (every shard keeps halo + 1 rows: `halo` for its neighbour, one to reflect about)."""
    sp = ShardPlan(n, chunks, halo)
    assert sp.nshards == expect
    assert sp.ranges[0][0] == 0 and sp.ranges[-1][1] == n
    assert sp.nshards == 1 or all(hi - lo >= halo + 1 for lo, hi in sp.ranges)


def test_choose_shard_axis():
    from nd_b200._filters import choose_shard_axis
    assert choose_shard_axis((1, 300, 200, 4), (0, 3, 3), (0, 1, 1), 4) == 1       # 2-D: the free axis has one row
    assert choose_shard_axis((24, 300, 200, 4), (0, 3, 3), (0, 1, 1), 4) == 0      # a real free axis (time first)
    assert choose_shard_axis((4, 300, 200, 4), (0, 3, 3), (0, 1, 1), 8) == 1       # free axis shorter than njobs
    assert choose_shard_axis((100, 300, 20, 4), (2, 2, 1), (1, 1, 1), 2) == 1      # all filtered: the largest


@pytest.mark.parametrize("n,rows,halo", [(4096, 126, 6), (4108, 126, 6), (100, 48, 4), (200, 64, 0), (130, 64, 6), (50, 100, 3),
                                         (37, 5, 6)])
def test_slab_plan_from_rows(n, rows, halo):
    """Slabs of the host pipeline: contiguous, complete, every slab tall enough to carry its halo, buffers clipped."""
    sp = ShardPlan.from_rows(n, rows, halo)
    assert sp.ranges[0][0] == 0 and sp.ranges[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(sp.ranges[:-1], sp.ranges[1:]))
    if sp.nshards > 1:
        assert all(hi - lo >= halo + 1 for lo, hi in sp.ranges)
    for i, (lo, hi) in enumerate(sp.ranges):
        blo, bhi = sp.buffered_range(i)
        assert blo == max(lo - halo, 0) and bhi == min(hi + halo, n)
        if i > 0:
            assert lo - blo == halo                      # interior edges always have the full halo available
        if i < sp.nshards - 1:
            assert bhi - hi == halo


def test_dataset_standin_roundtrip():
    ds = generate_test_dataset(dims={'y': 6, 'x': 7, 'time': 3})
    assert list(ds.data_vars) == ['C11', 'C12__im', 'C12__re', 'C22']
    assert ds['C11'].dims == ('y', 'x', 'time') and ds['C11'].values.dtype == np.float64
    assert list(ds.dims) == ['time', 'x', 'y']                      # alphabetical like xarray
    t0 = ds.isel(time=0)
    assert t0['C11'].dims == ('y', 'x') and t0['C11'].shape == (6, 7)
    c = ds.copy(deep=True)
    c['C11'].values[0, 0, 0] += 1
    assert not ds.equals(c)


# ---- round 2: host-side pieces of the streaming layer (no GPU needed) ----------------------------------
def test_parallel_staging_copy_and_variable_list():
    import torch
    from nd_b200 import stream
    a = torch.arange(3 * 70 * 1024 * 64, dtype=torch.float32).reshape(3 * 70, 1024, 64)      # ~55 MB: below the split size
    b = torch.empty_like(a)
    stream.par_copy(b, a)
    assert torch.equal(a, b)
    big = torch.ones(5, 1 << 22, dtype=torch.float32)                                           # 80 MB: split over threads
    out = torch.zeros_like(big)
    stream.par_copy(out, big)
    assert torch.equal(out, big)
    v = stream._VarList([np.full((4, 5, 6), k, np.float32) for k in range(3)])
    assert v[(2,)].shape == (4, 5, 6) and float(v[(2,)][0, 0, 0]) == 2.0 and not v.is_pinned()


def test_nlmeans_variables_declines_what_it_cannot_stream():
    from nd_b200._filters import nlmeans_variables
    u = lambda *x: np.array(x, dtype=np.uint32)
    a = [np.zeros((64, 20, 6), np.float32) for _ in range(3)]
    o = [np.empty_like(x) for x in a]
    assert nlmeans_variables(a[:2] + [a[2][:, ::2]], o[:2] + [o[2][:, ::2]], u(2, 2, 1), u(1, 1, 1), 1, 1) is False   # not contiguous
    assert nlmeans_variables(a, o[:2], u(2, 2, 1), u(1, 1, 1), 1, 1) is False                                      # one output missing
    assert nlmeans_variables([x[:10] for x in a], [x[:10] for x in o], u(2, 2, 1), u(1, 1, 1), 1, 1) is False      # too few rows
    assert nlmeans_variables([x.astype(np.int16) for x in a], o, u(2, 2, 1), u(1, 1, 1), 1, 1) is False
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            nlmeans_variables(a, o, u(2, 2, 1), u(1, 1, 1), 1, 1)


def test_apply_streams_per_variable_arrays_and_falls_back(monkeypatch):
    """Host plumbing of `Filter.apply` for per_variable=False filters (no GPU: the two kernels entry points are
    replaced by NumPy stand-ins): variables that already have the kernel's layout go to `_filter_variables` one array
    each, everything else is gathered into the (dims..., variable) block of the reference and goes to `_filter`;
    dims / coords / attrs / untouched variables come out as the reference's `apply` leaves them (nd/filters.py:105-191)."""
    from nd_b200.dataset import generate_test_dataset
    calls = []

    class Doubler(NLMeansFilter):
        def _filter_variables(self, arrays, axes, outputs):
            calls.append(("vars", len(arrays), arrays[0].shape, axes))
            for a, o in zip(arrays, outputs):
                o[...] = 2 * a
            return True

        def _filter(self, arr, axes, output):
            calls.append(("block", arr.shape, axes))
            output[...] = 3 * arr

    ds = generate_test_dataset(dims={'y': 12, 'x': 9, 'time': 4})
    ds['mask'] = (('y', 'x'), np.ones((12, 9)))
    ds.attrs['crs'] = 'EPSG:4326'
    out = Doubler(dims=('y', 'x', 'time'), r=1).apply(ds)
    assert calls == [("vars", 4, (12, 9, 4), (0, 1, 2))]
    for v in ('C11', 'C22', 'C12__re', 'C12__im'):
        assert np.array_equal(out[v].values, 2 * ds_values(ds, v)) and out[v].dims == ('y', 'x', 'time')
    assert np.array_equal(out['mask'].values, ds['mask'].values) and out['mask'].values is not ds['mask'].values
    assert dict(out.attrs) == dict(ds.attrs) and all(np.array_equal(out.coords[c], ds.coords[c]) for c in ds.coords)
    assert list(ds.data_vars) == ['C11', 'C12__im', 'C12__re', 'C22', 'mask']       # the input is left as it was

    # filter dims in another order than the variables' dims: the reference's gather / transpose path
    calls.clear()
    del ds['mask']                       # (a 2-D variable beside 3-D ones would be filtered too with dims=('x','y'))
    out2 = Doubler(dims=('x', 'y'), r=1).apply(ds)
    assert calls and calls[0][0] == "block" and calls[0][1] == (9, 12, 4, 4)
    assert np.array_equal(out2['C11'].values, 3 * ds['C11'].values) and out2['C11'].dims == ('y', 'x', 'time')

    # a filter that declines (returns False) falls back to the block path too
    calls.clear()
    monkeypatch.setattr(Doubler, "_filter_variables", lambda self, a, ax, o: False)
    out3 = Doubler(dims=('y', 'x', 'time'), r=1).apply(ds)
    assert [c[0] for c in calls] == ["block"] and np.array_equal(out3['C22'].values, 3 * ds['C22'].values)


def ds_values(ds, name):
    """Values of a (possibly complex-split) variable of the stand-in Dataset after `apply` reassembled the input."""
    if name in ds.data_vars:
        return ds[name].values
    stem, part = name.rsplit('__', 1)
    return ds[stem].values.real if part == 're' else ds[stem].values.imag
