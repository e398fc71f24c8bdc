"""
GPU parity tests (run on the B200 box with `-m gpu`): the CUDA path, called through the C ABI
(libndnlm.so via ctypes -- `nd_b200.device.Plan` / `nd_b200._filters._pixelwise_nlmeans_3d`), against
  * the committed golden vectors = outputs of the reference's own compiled kernel (tests/golden/),
  * the C oracle on seeded inputs (bit-exact twin of the reference, tests/test_oracle.py),
  * size-independent properties at larger sizes (shard invariance, identity, box mean, loader equality).
Tolerance: the north_star's max relative error <= 1e-4, measured per variable as
max|out-ref| / max|ref_v|; float64 data must agree to 1e-12.
Nothing here reads /root/reference.
"""
import itertools
import os

import numpy as np
import pytest

from helpers import sar_like, scaled_err

pytestmark = pytest.mark.gpu

TOL32 = 1e-4       # north_star tolerance (fp32 compute)
TOL64 = 1e-12


@pytest.fixture(scope="module")
def dev():
    import torch
    from nd_b200 import device
    assert torch.cuda.is_available()
    return device


def run_plan(device, a, r, f, sigma, h, n_eff=-1, semantics="as_written", kernel="auto"):
    import torch
    plan = device.Plan(a.shape, r, f, sigma, h, n_eff, semantics=semantics, dtype=a.dtype, kernel=kernel)
    out = plan.apply(torch.from_numpy(a).cuda())
    return out.cpu().numpy(), plan


# ---- golden vectors: outputs of the reference's own binary ---------------------------------------
@pytest.mark.parametrize("semantics,key", [("as_written", "__out_as_written"), ("reference_compiled", "__out_compiled")])
def test_golden_vectors(dev, golden, semantics, key):
    z, meta = golden
    launches0 = dev.launch_count()
    for name, m in meta.items():
        a = z[name + "__in"]
        out, plan = run_plan(dev, a, m["r"], m["f"], m["sigma"], m["h"], m["n_eff"], semantics)
        tol = TOL64 if a.dtype == np.float64 else TOL32
        err = scaled_err(out, z[name + key])
        assert err < tol, (name, semantics, plan.kernel_name, err)
        assert out.dtype == a.dtype
    assert dev.launch_count() > launches0           # the CUDA library really ran


def test_golden_vectors_tiled_kernel_is_used_and_tight(dev, golden):
    """float32 as-written cases with V<=4 must go through the TMA-tiled kernel and agree far better than 1e-4."""
    z, meta = golden
    for name in ("3d_f1_f32", "2d_f1_slc", "3d_f0_f32", "3d_batch_t", "1d_f1_V1"):
        m = meta[name]
        out, plan = run_plan(dev, z[name + "__in"], m["r"], m["f"], m["sigma"], m["h"], m["n_eff"], kernel="tiled")
        assert plan.is_tiled
        assert scaled_err(out, z[name + "__out_as_written"]) < 5e-6, name


# ---- seeded inputs vs the C oracle ----------------------------------------------------------------
CASES = [
    # shape, r, f  (as-written, float32)
    ((20, 45, 14, 4), (2, 3, 1), (1, 1, 1)),
    ((30, 70, 16, 4), (5, 5, 2), (1, 1, 1)),       # cfg3 parameters
    ((33, 61, 9, 4), (3, 3, 2), (1, 1, 1)),        # sizes that are not multiples of the tile
    ((7, 7, 4, 4), (3, 3, 1), (1, 1, 1)),          # minimal extents: r+f = N-1 on two axes ... reflection at both ends
    ((1, 40, 70, 4), (0, 3, 3), (0, 1, 1)),        # 2-D (cfg1 parameters)
    ((20, 33, 5, 4), (2, 2, 0), (1, 1, 0)),        # time is a batch axis
    ((6, 20, 15, 4), (1, 2, 1), (0, 0, 0)),        # f = 0
    ((5, 6, 50, 2), (0, 0, 4), (0, 0, 1)),         # 1-D, V=2
    ((14, 37, 8, 3), (1, 2, 2), (1, 1, 1)),        # V=3
    ((14, 37, 8, 1), (1, 2, 2), (1, 1, 1)),        # V=1
    ((12, 18, 9, 4), (2, 2, 1), (2, 2, 2)),        # f = 2
    ((9, 12, 8, 6), (1, 1, 1), (1, 1, 1)),         # V=6 (cfg5 variable count)
    ((8, 9, 7, 9), (1, 1, 1), (1, 1, 1)),          # V > 8: several passes of the generic kernel
    ((10, 12, 6, 4), (0, 2, 0), (1, 1, 1)),        # patch axes that are not search axes (direct C-ABI use)
    ((13, 31, 8, 4), (7, 7, 2), (2, 2, 2)),        # cfg4 parameters (tiled f = 2, several W passes)
    ((14, 29, 9, 6), (10, 10, 3), (2, 2, 2)),      # cfg5 parameters (V = 6, f = 2: two variable groups, one W offset per pass)
    ((1, 24, 40, 6), (0, 4, 4), (0, 2, 2)),        # 2-D, V = 6, f = 2
    ((10, 33, 8, 5), (2, 2, 1), (1, 1, 1)),        # V = 5: the second variable group is a "half" group (one real variable)
    ((10, 33, 8, 7), (2, 2, 1), (1, 1, 1)),        # V = 7: full second group
    ((16, 40, 9, 6), (5, 5, 2), (1, 1, 1)),        # V = 6 with the cfg3 radii (12-warp half-group instantiation)
]


@pytest.mark.parametrize("shape,r,f", CASES)
def test_matches_oracle_float32(dev, c_oracle, shape, r, f):
    a = sar_like(shape, seed=sum(shape), dtype=np.float32)
    ref = c_oracle.nlmeans(a, r, f, 0.3, 0.6)
    out, plan = run_plan(dev, a, r, f, 0.3, 0.6)
    assert scaled_err(out, ref) < TOL32, plan.kernel_name
    assert not np.isnan(out).any()


def test_half_variable_group_instantiations_are_selected(dev):
    """V = 5, 6: the second float4 group carries at most two variables; its upper lanes are skipped (HALF
    instantiations).  V = 7, 8 and V <= 4 must not take them."""
    for V, half in ((4, False), (5, True), (6, True), (7, False), (8, False)):
        for r, f in (((2, 2, 1), (1, 1, 1)), ((3, 3, 1), (2, 2, 2)), ((0, 3, 3), (0, 1, 1))):
            plan = dev.Plan((12, 40, 9, V), r, f, 0.3, 0.6)
            assert plan.is_tiled and ("(half)" in plan.kernel_name) == half, (V, r, plan.kernel_name)


@pytest.mark.parametrize("shape,r,f", CASES[:3] + CASES[4:7])
def test_generic_kernel_matches_oracle_closely(dev, c_oracle, shape, r, f):
    """The generic kernel mirrors the reference's arithmetic (fp64 weights): agreement ~1e-6."""
    a = sar_like(shape, seed=1 + sum(shape), dtype=np.float32)
    ref = c_oracle.nlmeans(a, r, f, 0.3, 0.6)
    out, plan = run_plan(dev, a, r, f, 0.3, 0.6, kernel="generic")
    assert "generic" in plan.kernel_name and scaled_err(out, ref) < 3e-6


@pytest.mark.parametrize("shape,r,f", [CASES[0], CASES[4], CASES[5], CASES[10]])
def test_float64_matches_oracle(dev, c_oracle, shape, r, f):
    a = sar_like(shape, seed=3, dtype=np.float64)
    for sem in ("as_written", "reference_compiled"):
        ref = c_oracle.nlmeans(a, r, f, 0.3, 0.6, semantics=sem)
        out, plan = run_plan(dev, a, r, f, 0.3, 0.6, semantics=sem)
        assert out.dtype == np.float64 and scaled_err(out, ref) < TOL64, (plan.kernel_name, sem)


@pytest.mark.parametrize("shape,r,f", [CASES[0], CASES[1], CASES[2], CASES[3], CASES[4], CASES[5], CASES[6], CASES[7],
                                       CASES[8], CASES[9], CASES[10], CASES[14], ((1, 24, 40, 4), (0, 4, 4), (0, 2, 2)),
                                       ((40, 70, 33, 4), (5, 5, 2), (1, 1, 1))])
def test_float64_tiled_kernel_matches_oracle_to_1e12(dev, c_oracle, shape, r, f):
    """float64 data (every fixture of the reference, nd/testing.py:68-69) runs on the float64 instantiations of the
    tiled kernel: the reference's own float64 arithmetic (nd/_filters.pyx:320-321), <= 1e-12 from the oracle."""
    a = sar_like(shape, seed=17 + sum(shape), dtype=np.float64)
    ref = c_oracle.nlmeans(a, r, f, 0.3, 0.6, threads=8)
    out, plan = run_plan(dev, a, r, f, 0.3, 0.6)
    assert plan.is_tiled and "double" in plan.kernel_name, plan.kernel_name
    assert out.dtype == np.float64 and scaled_err(out, ref) < TOL64, plan.kernel_name
    out_g, plan_g = run_plan(dev, a, r, f, 0.3, 0.6, kernel="generic")
    assert "generic" in plan_g.kernel_name and scaled_err(out_g, ref) < TOL64


def test_float64_tiled_neff_strided_and_sharded(dev, c_oracle):
    from nd_b200._filters import _pixelwise_nlmeans_3d
    a = sar_like((60, 40, 9, 4), seed=23, dtype=np.float64)
    r, f = (2, 2, 1), (1, 1, 1)
    ref = c_oracle.nlmeans(a, r, f, 0.3, 1.5, n_eff=6.0, threads=8)
    out, plan = run_plan(dev, a, r, f, 0.3, 1.5, n_eff=6.0)
    assert "double" in plan.kernel_name and scaled_err(out, ref) < TOL64
    with pytest.raises(ValueError, match="No solution"):
        run_plan(dev, a, (1, 1, 1), (0, 0, 0), 0.01, 0.01, n_eff=20.0)
    vm = np.moveaxis(np.ascontiguousarray(np.moveaxis(a, -1, 0)), 0, -1)           # variable-major view
    whole = np.empty_like(vm)
    _pixelwise_nlmeans_3d(vm, whole, np.array(r, np.uint32), np.array(f, np.uint32), 0.3, 0.6)
    ref2 = c_oracle.nlmeans(a, r, f, 0.3, 0.6, threads=8)
    assert scaled_err(whole, ref2) < TOL64
    parts = np.full_like(vm, np.nan)
    _pixelwise_nlmeans_3d(vm, parts, np.array(r, np.uint32), np.array(f, np.uint32), 0.3, 0.6, devices=[0, 0], pipeline=False)
    assert np.array_equal(parts, whole)                                           # halo rows as 32-byte voxels
    piped = np.full_like(vm, np.nan)
    _pixelwise_nlmeans_3d(vm, piped, np.array(r, np.uint32), np.array(f, np.uint32), 0.3, 0.6, pipeline=True, slab_rows=24)
    assert np.array_equal(piped, whole)


@pytest.mark.parametrize("shape,r,f", [CASES[1], CASES[4], CASES[10]])
def test_float64_data_on_the_fp32_tiled_kernel_on_request(dev, c_oracle, shape, r, f):
    """kernel='tiled' with float64 data: staged as float32, result widened back -- north_star's fp32 compute, far
    inside its 1e-4 (the default for float64 data stays the float64 generic kernel, see above)."""
    a = sar_like(shape, seed=5, dtype=np.float64)
    ref = c_oracle.nlmeans(a, r, f, 0.3, 0.6)
    out, plan = run_plan(dev, a, r, f, 0.3, 0.6, kernel="tiled")
    assert plan.is_tiled and out.dtype == np.float64
    assert scaled_err(out, ref) < 5e-6, plan.kernel_name
    out32, _ = run_plan(dev, a.astype(np.float32), r, f, 0.3, 0.6, kernel="tiled")
    assert np.array_equal(out.astype(np.float32), out32)          # exactly the float32 path


@pytest.mark.parametrize("shape,r,f", [CASES[0], CASES[1], CASES[3], CASES[4], CASES[5], CASES[7], CASES[10], CASES[11],
                                       CASES[14], CASES[15], ((40, 300, 70, 4), (5, 5, 2), (1, 1, 1)),
                                       ((9, 40, 11, 4), (4, 6, 4), (1, 1, 1))])
def test_reference_compiled_semantics(dev, c_oracle, shape, r, f):
    """Bug-for-bug the LP64 binary: a reflect box mean.  float32 data with the default self weight
    runs on the separable box-mean fast path (nlm_boxmean.cuh), whose float32 sums are ordered differently from
    the reference's sequential float32 accumulation: agreement to rounding noise, not bitwise."""
    a = sar_like(shape, seed=4, dtype=np.float32)
    ref = c_oracle.nlmeans(a, r, f, 0.3, 0.6, semantics="reference_compiled", threads=8)
    out, plan = run_plan(dev, a, r, f, 0.3, 0.6, semantics="reference_compiled")
    assert "boxmean" in plan.kernel_name, plan.kernel_name
    assert scaled_err(out, ref) < 1e-5, plan.kernel_name
    out_g, plan_g = run_plan(dev, a, r, f, 0.3, 0.6, semantics="reference_compiled", kernel="generic")
    assert "generic" in plan_g.kernel_name and scaled_err(out_g, ref) < 3e-6


def test_reference_compiled_neff_stays_on_the_generic_kernel(dev, c_oracle):
    a = sar_like((10, 16, 7, 4), seed=8, dtype=np.float32)
    ref = c_oracle.nlmeans(a, (2, 2, 1), (1, 1, 1), 0.3, 0.6, n_eff=6.0, semantics="reference_compiled")
    out, plan = run_plan(dev, a, (2, 2, 1), (1, 1, 1), 0.3, 0.6, n_eff=6.0, semantics="reference_compiled")
    assert "generic" in plan.kernel_name and scaled_err(out, ref) < 3e-6


def test_neff(dev, c_oracle):
    a = sar_like((10, 16, 7, 4), seed=5, dtype=np.float32)
    for r, f in (((2, 2, 1), (1, 1, 1)), ((0, 2, 2), (0, 1, 1)), ((2, 2, 1), (2, 2, 2)), ((0, 3, 2), (0, 2, 2))):
        ref = c_oracle.nlmeans(a, r, f, 0.3, 1.5, n_eff=6.0)
        out, plan = run_plan(dev, a, r, f, 0.3, 1.5, n_eff=6.0)
        assert plan.is_tiled and scaled_err(out, ref) < TOL32, plan.kernel_name


def test_neff_no_solution_raises_value_error(dev):
    a = sar_like((8, 9, 6, 2), seed=6, dtype=np.float32)
    with pytest.raises(ValueError, match="No solution"):      # nd/_filters.pyx:310-311
        run_plan(dev, a, (1, 1, 1), (0, 0, 0), 0.01, 0.01, n_eff=20.0)


def test_weights_below_the_fp32_range_warn_instead_of_nan(dev):
    """with h so small that every neighbour weight flushes to zero in fp32 the voxel used to come out as
    0/0 (n_eff >= 0) or silently unfiltered.  Now: unfiltered AND a RuntimeWarning (flag bit 1), never a NaN."""
    a = sar_like((12, 40, 9, 4), seed=9, dtype=np.float32)
    for n_eff in (-1, 6.0):
        with pytest.warns(RuntimeWarning, match="float32 range"):
            out, plan = run_plan(dev, a, (2, 2, 1), (1, 1, 1), 0.001, 0.005, n_eff=n_eff, kernel="tiled")
        assert plan.is_tiled and np.array_equal(out, a)


def test_nan_footprint_matches_reference(dev, c_oracle):
    a = sar_like((1, 15, 15, 3), seed=2, dtype=np.float32)
    a[0, 7, 7, 0] = np.nan
    for f, sem in (((0, 0, 0), "as_written"), ((0, 1, 1), "as_written"), ((0, 1, 1), "reference_compiled")):
        ref = c_oracle.nlmeans(a, (0, 2, 2), f, 0.3, 0.6, semantics=sem)
        out, plan = run_plan(dev, a, (0, 2, 2), f, 0.3, 0.6, semantics=sem)
        assert np.array_equal(np.isnan(out), np.isnan(ref)), (plan.kernel_name, f, sem)
        assert scaled_err(np.nan_to_num(out), np.nan_to_num(ref)) < TOL32


def test_strided_variable_major_input(dev, c_oracle):
    """What Filter.apply hands over: a (dims..., variable) VIEW of a variable-major block (nd/filters.py:170)."""
    import torch
    a = sar_like((12, 20, 8, 4), seed=7, dtype=np.float32)
    block = np.ascontiguousarray(np.moveaxis(a, -1, 0))
    view = np.moveaxis(block, 0, -1)
    ref = c_oracle.nlmeans(a, (2, 2, 1), (1, 1, 1), 0.3, 0.6)
    plan = dev.Plan(a.shape, (2, 2, 1), (1, 1, 1), 0.3, 0.6)
    t = torch.from_numpy(view).cuda()
    assert not t.is_contiguous()
    out = torch.empty_like(t)
    plan.apply(t, out)
    assert out.stride() == t.stride()
    assert scaled_err(out.cpu().numpy(), ref) < TOL32


def test_tma_loader_equals_plain_loader_bitwise(dev):
    a = sar_like((25, 64, 12, 4), seed=8, dtype=np.float32)
    out_tma, plan = run_plan(dev, a, (3, 3, 2), (1, 1, 1), 0.3, 0.6, kernel="tiled")
    os.environ["NDNLM_LOADER"] = "ldg"
    try:
        out_ldg, _ = run_plan(dev, a, (3, 3, 2), (1, 1, 1), 0.3, 0.6, kernel="tiled")
    finally:
        del os.environ["NDNLM_LOADER"]
    assert np.array_equal(out_tma, out_ldg)


# ---- properties at larger sizes --------------------------------------------------------------------
def test_constant_cube_identity_and_huge_sigma_box_mean(dev):
    a = np.full((40, 70, 10, 4), 2.5, dtype=np.float32)
    out, _ = run_plan(dev, a, (3, 3, 2), (1, 1, 1), 0.1, 0.2)
    assert np.array_equal(out, a)
    b = sar_like((16, 40, 8, 4), seed=9, dtype=np.float32)
    r = (2, 3, 1)
    out, _ = run_plan(dev, b, r, (1, 1, 1), 1e3, 1.0)
    P = np.pad(b.astype(np.float64), [(k, k) for k in r] + [(0, 0)], mode="reflect")
    box = np.zeros(b.shape)
    for t in itertools.product(*[range(2 * k + 1) for k in r]):
        box += P[t[0]:t[0] + 16, t[1]:t[1] + 40, t[2]:t[2] + 8]
    box /= np.prod([2 * k + 1 for k in r])
    assert scaled_err(out, box) < 2e-6


def test_deterministic(dev):
    a = sar_like((64, 100, 32, 4), seed=10, dtype=np.float32)
    o1, _ = run_plan(dev, a, (3, 3, 2), (1, 1, 1), 0.25, 0.5)
    o2, _ = run_plan(dev, a, (3, 3, 2), (1, 1, 1), 0.25, 0.5)
    assert np.array_equal(o1, o2)


@pytest.mark.parametrize("dtype,r,f,shape", [
    (np.float32, (3, 3, 2), (1, 1, 1), (37, 70, 11, 4)),       # f = 1: 15 valid rows per CTA instead of 14
    (np.float32, (4, 3, 2), (2, 2, 2), (33, 40, 10, 3)),       # f = 2: 10 instead of 8 (the default there)
    (np.float32, (3, 3, 1), (2, 2, 2), (23, 36, 9, 4)),
    (np.float64, (2, 3, 1), (1, 1, 1), (19, 40, 9, 2)),        # float64: 7 instead of 6
    (np.float64, (2, 2, 1), (2, 2, 2), (17, 36, 9, 4)),        # float64, f = 2: 6 instead of 4
])
def test_double_duty_halo_warps_equal_plain_instantiations_bitwise(dev, monkeypatch, dtype, r, f, shape):
    """instances_g6.inc: the halo rows of a CTA tile are served two per warp (no accumulators in those warps).  The
    arithmetic per voxel is that of the plain instantiations, so the results must agree bit for bit -- and with the
    oracle.  NDNLM_DH is read at plan creation."""
    a = sar_like(shape, seed=23, dtype=dtype)
    monkeypatch.setenv("NDNLM_DH", "0")
    plain, p0 = run_plan(dev, a, r, f, 0.3, 0.6)
    monkeypatch.setenv("NDNLM_DH", "1")
    dh, p1 = run_plan(dev, a, r, f, 0.3, 0.6)
    assert "(dh)" in p1.kernel_name and "(dh)" not in p0.kernel_name, (p0.kernel_name, p1.kernel_name)
    assert p1.info.tile[0] > p0.info.tile[0]
    assert np.array_equal(plain, dh), (p1.kernel_name, float(np.abs(plain - dh).max()))
    from oracle import c_port
    assert scaled_err(dh, c_port.nlmeans(a, r, f, 0.3, 0.6)) < (TOL64 if dtype == np.float64 else TOL32)


@pytest.mark.parametrize("dtype,sem", [(np.float32, "as_written"), (np.float64, "as_written"), (np.float32, "reference_compiled")])
def test_kernels_write_a_native_output_in_place(dev, dtype, sem):
    """A C-ordered (y, x, time, 4) output already has the kernels' own output layout (ndnlm_output_is_native): it is
    written directly, bitwise equal to the route through the internal buffer + unstage (taken for any other layout),
    for the one-call ndnlm_apply, Plan.run_into and the device slab streaming."""
    import torch
    from nd_b200 import stream
    a = sar_like((41, 50, 12, 4), seed=31, dtype=dtype)
    r, f = (3, 2, 1), (1, 1, 1)
    d_in = torch.from_numpy(a).cuda()
    plan = dev.Plan(a.shape, r, f, 0.3, 0.6, semantics=sem, dtype=dtype)
    out_native = torch.full_like(d_in, float("nan"))
    assert plan.output_is_native(out_native)
    plan.apply(d_in, out_native)
    vmajor = torch.empty((4,) + a.shape[:3], dtype=d_in.dtype, device="cuda").permute(1, 2, 3, 0)     # not native
    assert not plan.output_is_native(vmajor)
    plan.apply(d_in, vmajor)
    assert torch.equal(out_native, vmajor)
    assert not plan.output_is_native(torch.empty((a.shape[0], a.shape[1], a.shape[2] + 1, 4), dtype=d_in.dtype, device="cuda")[:, :, 1:])  # shape ok, rows not dense
    # run_into: direct and through the internal buffer
    padded = plan.new_padded("cuda"); flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    plan.stage(d_in, padded)
    o2 = torch.full_like(d_in, float("nan"))
    assert plan.run_into(padded, o2, flag) is True and torch.equal(o2, out_native)
    o3 = torch.empty((4,) + a.shape[:3], dtype=d_in.dtype, device="cuda").permute(1, 2, 3, 0)
    assert plan.run_into(padded, o3, flag) is False and torch.equal(o3, out_native)
    # device slab streaming writes its (contiguous) output slabs in place
    got = torch.full_like(d_in, float("nan"))
    stream.apply_device_streamed(lambda lo, hi, dst: dst.copy_(d_in[lo:hi]), lambda lo, hi, res: got[lo:hi].copy_(res),
                                 a.shape[0], a.shape[1:], r, f, 0.3, 0.6, semantics=sem, slab_rows=16, dtype=dtype)
    if sem == "reference_compiled":
        # the box-mean kernels update their running sums along march segments that depend on the slab height: a slab
        # agrees with the whole cube to rounding, not bitwise
        assert scaled_err(got.cpu().numpy(), out_native.cpu().numpy()) < 1e-5
    else:
        assert torch.equal(got, out_native)
    from oracle import c_port
    tol = TOL64 if dtype == np.float64 else (1e-5 if sem == "reference_compiled" else TOL32)
    assert scaled_err(out_native.cpu().numpy(), c_port.nlmeans(a, r, f, 0.3, 0.6, semantics=sem)) < tol


def test_double_duty_halo_warps_n_eff_and_default_choice(dev, monkeypatch):
    from oracle import c_port
    a = sar_like((31, 40, 10, 4), seed=29, dtype=np.float32)
    monkeypatch.delenv("NDNLM_DH", raising=False)
    out, plan = run_plan(dev, a, (3, 3, 1), (2, 2, 2), 0.4, 0.8, n_eff=30)
    assert "(dh)" in plan.kernel_name and "neff=true" in plan.kernel_name        # the default for f = 2
    assert scaled_err(out, c_port.nlmeans(a, (3, 3, 1), (2, 2, 2), 0.4, 0.8, 30)) < TOL32
    _, p1 = run_plan(dev, a, (3, 3, 1), (1, 1, 1), 0.4, 0.8)
    assert "(dh)" not in p1.kernel_name                                          # f = 1 keeps one warp per row


@pytest.mark.parametrize("nshards", [2, 3])
def test_sharded_equals_unsharded_bitwise(dev, nshards):
    """The halo layer on ONE GPU (devices=[0,0,..]): y-shards + neighbour halo rows == unsharded, bit for bit
    (the reference's test_parallelized_filter, nd/tests/test_filters_common.py:54-60, asks rtol 1e-5)."""
    from nd_b200._filters import _pixelwise_nlmeans_3d
    a = sar_like((50, 64, 10, 4), seed=11, dtype=np.float32)
    r, f = np.array([3, 3, 1], np.uint32), np.array([1, 1, 1], np.uint32)
    whole = np.empty_like(a)
    _pixelwise_nlmeans_3d(a, whole, r, f, 0.3, 0.6)
    parts = np.empty_like(a)
    _pixelwise_nlmeans_3d(a, parts, r, f, 0.3, 0.6, devices=[0] * nshards, shard_axis=0)
    assert np.array_equal(whole, parts)
    # sharding along a non-filtered axis needs no halo at all
    r2, f2 = np.array([2, 2, 0], np.uint32), np.array([1, 1, 0], np.uint32)
    _pixelwise_nlmeans_3d(a, whole, r2, f2, 0.3, 0.6)
    _pixelwise_nlmeans_3d(a, parts, r2, f2, 0.3, 0.6, devices=[0] * nshards, shard_axis=2)
    assert np.array_equal(whole, parts)


def test_sharded_threaded_host_pipelines_equal_unsharded_bitwise(dev):
    """njobs > 1 over a host array: one slab pipeline per shard, each in its own thread, reading its neighbours'
    r+f buffer rows from the host array (two shards on ONE GPU here; two GPUs in test_njobs_two_gpus...)."""
    from nd_b200._filters import _pixelwise_nlmeans_3d
    a = sar_like((230, 40, 8, 4), seed=14, dtype=np.float32)
    r, f = np.array([3, 3, 1], np.uint32), np.array([1, 1, 1], np.uint32)
    whole = np.empty_like(a)
    _pixelwise_nlmeans_3d(a, whole, r, f, 0.3, 0.6, pipeline=False)
    for vmajor in (False, True):
        b = np.moveaxis(np.ascontiguousarray(np.moveaxis(a, -1, 0)), 0, -1) if vmajor else a
        parts = np.full_like(b, np.nan)
        _pixelwise_nlmeans_3d(b, parts, r, f, 0.3, 0.6, devices=[0, 0, 0])
        assert np.array_equal(whole, parts), vmajor
        resident = np.full_like(b, np.nan)
        _pixelwise_nlmeans_3d(b, resident, r, f, 0.3, 0.6, devices=[0, 0, 0], pipeline=False)    # halo rows over peer copies
        assert np.array_equal(whole, resident), vmajor


def test_uneven_shards_along_axis_1_keep_the_roles_of_the_whole_array(dev):
    """the X / R roles are chosen from the extents of axes 1 and 2, so shards cut along axis 1 whose
    extent drops below shape[2] used to stage another layout than their neighbours (scrambled halo rows).  Every
    shard now takes the roles of the whole array (ndnlm_plan_create_roles)."""
    from nd_b200._filters import _pixelwise_nlmeans_3d
    a = sar_like((12, 37, 20, 4), seed=15, dtype=np.float32)            # whole: 37 > 20; shards of 19 / 18 < 20
    r, f = np.array([1, 2, 2], np.uint32), np.array([1, 1, 1], np.uint32)
    whole = np.empty_like(a)
    _pixelwise_nlmeans_3d(a, whole, r, f, 0.3, 0.6)
    parts = np.full_like(a, np.nan)
    _pixelwise_nlmeans_3d(a, parts, r, f, 0.3, 0.6, devices=[0, 0], shard_axis=1)
    assert scaled_err(parts, whole) < 1e-6
    p_whole = dev.Plan(a.shape, r, f, 0.3, 0.6)
    p_shard = dev.Plan((12, 19, 20, 4), r, f, 0.3, 0.6)
    assert p_whole.roles != p_shard.roles                                 # the situation the fix is for
    assert dev.Plan((12, 19, 20, 4), r, f, 0.3, 0.6, roles=p_whole.roles).roles == p_whole.roles
    with pytest.raises(ValueError):
        dev.Plan(a.shape, r, f, 0.3, 0.6, roles=(0, 0, 1))


def test_large_cube_sampled_against_oracle(dev, c_oracle):
    """cfg3 parameters on a cube that is many tiles in every direction; sub-cubes (corner, edge, interior)
    are checked against the oracle run on the same values (with halo r+f around the sample)."""
    import torch
    ny, nx, nt = 96, 256, 32
    cube = dev.synth_cube(ny, nx, nt, 4)
    r, f = (5, 5, 2), (1, 1, 1)
    plan = dev.Plan(cube.shape, r, f, 0.25, 0.5)
    out = plan.apply(cube).cpu().numpy()
    a = cube.cpu().numpy()
    assert plan.is_tiled and np.isfinite(out).all()
    pad = (6, 6, 3)
    for (y0, x0) in ((0, 0), (ny - 12, nx - 20), (40, 100), (0, 117)):
        ys, xs = slice(max(y0 - pad[0], 0), min(y0 + 12 + pad[0], ny)), slice(max(x0 - pad[1], 0), min(x0 + 20 + pad[1], nx))
        sub = np.ascontiguousarray(a[ys, xs])
        ref = c_oracle.nlmeans(sub, r, f, 0.25, 0.5)
        # voxels whose whole window lies inside `sub` (or is cut by a TRUE cube edge, where both reflect alike)
        iy = slice(y0 - ys.start, y0 - ys.start + 12)
        ix = slice(x0 - xs.start, x0 - xs.start + 20)
        assert scaled_err(out[y0:y0 + 12, x0:x0 + 20], ref[iy, ix]) < TOL32, (y0, x0)


# ---- the reference's own filter tests, through the Dataset API -------------------------------------
def test_reference_filter_tests_through_dataset_api(dev):
    from nd_b200.dataset import generate_test_dataset
    from nd_b200.filters import NLMeansFilter, nlmeans
    ds = generate_test_dataset(dims={'y': 20, 'x': 20, 'time': 10})
    # nd/tests/test_nlmeans_filter.py:28-32 (reduce std)
    ds_nlm = NLMeansFilter(dims=('y', 'x', 'time'), r=(1, 1, 0), sigma=2, h=2).apply(ds)
    assert all(ds_nlm[v].values.std() < ds[v].values.std() for v in ds.data_vars)
    assert ds_nlm.attrs == ds.attrs and all(ds_nlm[v].dims == ds[v].dims for v in ds.data_vars)
    # :35-43 (r_time = 0 == per-slice 2-D call, to 1e-8)
    t0 = ds.isel(time=0)
    t0_nlm = NLMeansFilter(dims=('y', 'x'), r=1, sigma=2, h=2).apply(t0)
    for v in ds.data_vars:
        assert np.abs(ds_nlm.isel(time=0)[v].values - t0_nlm[v].values).max() < 1e-8
    # nd/tests/test_filters_common.py:44-51 (dims order does not matter, rtol 1e-5)
    a = NLMeansFilter(dims=('y', 'x')).apply(ds)
    b = nlmeans(ds, dims=('x', 'y'))
    for v in ds.data_vars:
        assert np.allclose(a[v].values, b[v].values, rtol=1e-5, atol=1e-8)
    # float32 Dataset goes through the tiled kernel and returns float32
    ds32 = generate_test_dataset(dims={'y': 24, 'x': 40, 'time': 6}, dtype=np.float32)
    out32 = NLMeansFilter(dims=('y', 'x', 'time'), r=(2, 2, 1), sigma=0.5, h=1.0).apply(ds32)
    assert all(out32[v].values.dtype == np.float32 for v in ds32.data_vars)


def test_complex_variables_are_split_and_reassembled(dev):
    from nd_b200.dataset import Dataset
    from nd_b200.filters import NLMeansFilter
    rng = np.random.default_rng(0)
    ds = Dataset({'C12': (('y', 'x'), (rng.normal(size=(12, 14)) + 1j * rng.normal(size=(12, 14)))),
                  'C11': (('y', 'x'), rng.gamma(4, 0.25, size=(12, 14)))}, coords={'y': np.arange(12), 'x': np.arange(14)})
    out = NLMeansFilter(dims=('y', 'x'), r=1, sigma=1, h=1).apply(ds)
    assert 'C12' in ds.data_vars and np.iscomplexobj(ds['C12'].values)          # input reassembled (nd/filters.py:188-189)
    assert set(out.data_vars) == {'C11', 'C12__re', 'C12__im'}                   # result keeps the split parts (quirk)


# ---- host slab pipeline ---------------------------------------------------------
@pytest.mark.parametrize("vmajor", [False, True])
def test_host_slab_pipeline_equals_monolithic_bitwise(dev, vmajor):
    """H2D / kernels / D2H overlapped over y-slabs with an r+f buffer == one monolithic call, bit for bit,
    for C-contiguous arrays and for the variable-major views Filter.apply produces."""
    from nd_b200._filters import _pixelwise_nlmeans_3d
    a = sar_like((150, 48, 8, 4), seed=12, dtype=np.float32)
    if vmajor:
        a = np.moveaxis(np.ascontiguousarray(np.moveaxis(a, -1, 0)), 0, -1)
    r, f = np.array([3, 3, 1], np.uint32), np.array([1, 1, 1], np.uint32)
    mono = np.empty_like(a)
    _pixelwise_nlmeans_3d(a, mono, r, f, 0.3, 0.6, pipeline=False)
    piped = np.full_like(a, np.nan)
    _pixelwise_nlmeans_3d(a, piped, r, f, 0.3, 0.6, pipeline=True, slab_rows=40)
    assert np.array_equal(mono, piped)
    piped2 = np.full_like(a, np.nan)
    _pixelwise_nlmeans_3d(a, piped2, r, f, 0.3, 0.6, pipeline=True, slab_rows=64)
    assert np.array_equal(mono, piped2)


def test_host_slab_pipeline_pinned_and_neff_error(dev):
    import torch
    from nd_b200._filters import _pixelwise_nlmeans_3d
    h_in = torch.empty((160, 40, 6, 4), dtype=torch.float32, pin_memory=True)
    h_in.copy_(torch.from_numpy(sar_like((160, 40, 6, 4), seed=13, dtype=np.float32)))
    h_out = torch.empty_like(h_in, pin_memory=True)
    a, o = h_in.numpy(), h_out.numpy()
    r, f = np.array([2, 2, 1], np.uint32), np.array([1, 1, 1], np.uint32)
    _pixelwise_nlmeans_3d(a, o, r, f, 0.3, 0.6, pipeline=True, slab_rows=48)
    mono = np.empty_like(a)
    _pixelwise_nlmeans_3d(a, mono, r, f, 0.3, 0.6, pipeline=False)
    assert np.array_equal(o, mono)
    with pytest.raises(ValueError, match="No solution"):
        _pixelwise_nlmeans_3d(a, o, np.array([1, 1, 1], np.uint32), np.array([0, 0, 0], np.uint32), 0.01, 0.01, 20.0,
                              pipeline=True, slab_rows=48)


def test_device_slab_streaming_equals_monolithic_bitwise(dev):
    """Cubes produced / consumed slab by slab on the device (BASELINE configs[3] / [4] do not fit in HBM with their
    staged copy and result): `apply_device_streamed` == one monolithic call, bit for bit -- also for a y-shard whose
    neighbour rows arrive as separate tensors (what the ranks exchange over NCCL in bench.py)."""
    import torch
    from nd_b200 import stream
    a = sar_like((140, 36, 8, 4), seed=16, dtype=np.float32)
    r, f = (3, 3, 1), (1, 1, 1)
    whole, _ = run_plan(dev, a, r, f, 0.3, 0.6)
    d = torch.from_numpy(a).cuda()

    def run(lo, hi, slab_rows, lo_rows, hi_rows):
        out = torch.full((hi - lo,) + a.shape[1:], float("nan"), device="cuda")
        seen = []

        def source(s0, s1, dst):
            dst.copy_(d[lo + s0:lo + s1])

        def sink(s0, s1, res):
            out[s0:s1] = res
            seen.append((s0, s1))
        n = stream.apply_device_streamed(source, sink, hi - lo, a.shape[1:], r, f, 0.3, 0.6, slab_rows=slab_rows,
                                         lo_rows=lo_rows, hi_rows=hi_rows)
        assert n == len(seen) and seen[0][0] == 0 and seen[-1][1] == hi - lo
        return out.cpu().numpy(), n

    out, n = run(0, 140, 40, None, None)
    assert n >= 3 and np.array_equal(out, whole)
    out, n = run(0, 140, 1000, None, None)                      # one slab
    assert n == 1 and np.array_equal(out, whole)
    halo = 4
    out, n = run(50, 110, 28, d[50 - halo:50].contiguous(), d[110:110 + halo].contiguous())     # a shard of the cube
    assert n >= 2 and np.array_equal(out, whole[50:110])
    out, n = run(0, 60, 28, None, d[60:60 + halo].contiguous())                                  # the first shard
    assert np.array_equal(out, whole[:60])
    with pytest.raises(ValueError):
        run(50, 110, 28, d[50 - 2:50].contiguous(), None)        # neighbour rows of the wrong height


# ---- real multi-GPU (skipped on a 1-GPU box) ----------------------------------------------------------
def test_njobs_two_gpus_equals_one_gpu(dev):
    """The reference's test_parallelized_filter (nd/tests/test_filters_common.py:54-60) with njobs = GPUs:
    y-shards on two devices, halo rows over peer copies, == single-GPU result bit for bit."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from nd_b200.dataset import generate_test_dataset
    from nd_b200.filters import NLMeansFilter
    ds = generate_test_dataset(dims={'y': 40, 'x': 30, 'time': 10}, dtype=np.float32)
    for dims in (('x', 'y'), ('x', 'y', 'time')):
        one = NLMeansFilter(dims=dims, r=2, sigma=1, h=1).apply(ds)
        two = NLMeansFilter(dims=dims, r=2, sigma=1, h=1).apply(ds, njobs=2)
        for v in ds.data_vars:
            assert np.array_equal(one[v].values, two[v].values), (dims, v)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_dataset_apply_streams_the_variables_and_equals_the_block_path(dev, dtype):
    """`Filter.apply` on a Dataset whose variables already have the kernel's layout streams them straight from / to
    the Dataset's arrays (no to_array() block, nd_b200.filters.Filter._apply_joint_streamed); the gather / filter /
    scatter path of the reference (nd/filters.py:164-185) must give the same bits."""
    from nd_b200.dataset import generate_test_dataset
    from nd_b200.filters import NLMeansFilter
    ds = generate_test_dataset(dims={'y': 70, 'x': 40, 'time': 6}, dtype=dtype)
    ds['mask'] = (('y', 'x'), np.ones((70, 40)))
    kw = dict(dims=('y', 'x', 'time'), r=(2, 2, 1), sigma=1, h=1)
    fast = NLMeansFilter(**kw)
    launches0 = dev.launch_count()
    out_fast = fast.apply(ds)
    assert dev.launch_count() > launches0
    slow = NLMeansFilter(**kw)
    slow._filter_variables = None                                  # force the block path
    out_slow = slow.apply(ds)
    for v in ds.data_vars:
        assert np.array_equal(out_fast[v].values, out_slow[v].values), v
        assert out_fast[v].dims == ds[v].dims and out_fast[v].values.dtype == ds[v].values.dtype
    assert out_fast['mask'].values is not ds['mask'].values         # untouched variables are copies, like deep copy
    assert out_fast['C11'].values.std() < ds['C11'].values.std()
    # a 2-D image (the reference's most common use): streamed along 'y' too; the block path lays the axes out as
    # (1, y, x), the streamed one as (y, x, 1) -- the same filter, another summation order
    ds2 = ds.isel(time=0)
    kw2 = dict(dims=('y', 'x'), r=(2, 3), sigma=1, h=1)
    fast2, slow2 = NLMeansFilter(**kw2), NLMeansFilter(**kw2)
    slow2._filter_variables = None
    o_fast, o_slow = fast2.apply(ds2), slow2.apply(ds2)
    for v in ('C11', 'C22', 'C12__re', 'C12__im'):
        assert o_fast[v].dims == ('y', 'x') and scaled_err(o_fast[v].values[..., None], o_slow[v].values[..., None]) < (
            1e-12 if dtype == np.float64 else 1e-6), v
    import torch
    if torch.cuda.device_count() >= 2:
        out_two = NLMeansFilter(**kw).apply(ds, njobs=2)
        for v in ds.data_vars:
            assert np.array_equal(out_two[v].values, out_fast[v].values), v


def _nccl_rank(rank, world, port, shape, r, f, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from nd_b200 import device
        from nd_b200.shard import DistributedShard, ShardPlan
        a = sar_like(shape, seed=31, dtype=np.float32)                  # every rank derives the same global cube
        sp = ShardPlan(shape[0], world, r[0] + f[0])
        lo, hi = sp.ranges[rank]
        slab = torch.from_numpy(np.ascontiguousarray(a[lo:hi])).cuda()
        plan = device.Plan(slab.shape, r, f, 0.3, 0.6)
        shard = DistributedShard(plan, axis=0, rank=rank, world=world)
        out = torch.empty_like(slab)
        shard.stage(slab)
        nmsg = shard.exchange()
        shard.run()
        shard.unstage(out)
        torch.cuda.synchronize()
        q.put((rank, lo, hi, nmsg, out.cpu().numpy()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_nccl_ranks_seam_rows_match_the_oracle(dev, c_oracle):
    """One process per GPU, halo rows over NCCL send/recv (the path bench.py scales on): every rank's rows -- the
    seam rows in particular -- against the oracle run on the whole cube, and bitwise against the unsharded GPU run."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import socket
    import torch.multiprocessing as mp
    shape, r, f = (44, 40, 9, 4), (3, 3, 1), (1, 1, 1)
    a = sar_like(shape, seed=31, dtype=np.float32)
    ref = c_oracle.nlmeans(a, r, f, 0.3, 0.6, threads=8)
    whole, _ = run_plan(dev, a, r, f, 0.3, 0.6)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_rank, args=(k, 2, port, shape, r, f, q)) for k in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, lo, hi, nmsg, out in got:
        assert nmsg == 2                                                 # one send + one recv per interior edge
        assert np.array_equal(out, whole[lo:hi]), rank
        assert scaled_err(out, ref[lo:hi]) < TOL32
        seam = slice(hi - 4, hi) if rank == 0 else slice(lo, lo + 4)
        assert scaled_err(out[seam.start - lo:seam.stop - lo], ref[seam]) < TOL32


def test_apply_on_a_real_xarray_dataset(dev):
    """north_star: `NLMeansFilter(...).apply(ds)` takes and returns an xarray.Dataset with the same dims, coords and
    attrs (reference nd/filters.py:105-191).  xarray is absent from this image; the test runs wherever it exists."""
    xr = pytest.importorskip("xarray")
    from nd_b200.filters import NLMeansFilter
    rng = np.random.default_rng(3)
    coords = {"y": np.linspace(50.0, 51.0, 24), "x": np.linspace(10.0, 11.0, 30), "time": np.arange(6)}
    ds = xr.Dataset({"C11": (("y", "x", "time"), rng.gamma(4.0, 0.25, (24, 30, 6))),
                     "C22": (("y", "x", "time"), rng.gamma(4.0, 0.25, (24, 30, 6))),
                     "C12": (("y", "x", "time"), rng.normal(size=(24, 30, 6)) + 1j * rng.normal(size=(24, 30, 6))),
                     "mask": (("y", "x"), np.ones((24, 30)))}, coords=coords, attrs={"crs": "EPSG:4326"})
    out = NLMeansFilter(dims=("y", "x", "time"), r=(2, 2, 1), sigma=0.3, h=0.6).apply(ds)
    assert isinstance(out, xr.Dataset) and dict(out.attrs) == dict(ds.attrs)
    # the input is reassembled, the result keeps the split real / imaginary parts (quirk of nd/filters.py:186-190)
    assert np.iscomplexobj(ds["C12"].values)
    assert set(out.data_vars) == {"C11", "C22", "C12__re", "C12__im", "mask"} and out["C11"].dims == ds["C11"].dims
    for c in coords:
        assert np.array_equal(out[c].values, ds[c].values)
    assert np.array_equal(out["mask"].values, ds["mask"].values)
    assert out["C11"].values.std() < ds["C11"].values.std()
    t_first = NLMeansFilter(dims=("time", "y", "x"), r=(1, 2, 2), sigma=0.3, h=0.6).apply(ds.transpose("time", "y", "x"))
    np.testing.assert_allclose(t_first["C11"].transpose("y", "x", "time").values, out["C11"].values, rtol=1e-5, atol=1e-7)

