"""
CPU tests of the oracles themselves (run with -m "not gpu"):
the C restatement and the NumPy restatement against the committed golden vectors (outputs of the
reference's own compiled kernel, tests/golden/make_golden.py), against the live compiled
reference when `oracle/_ref` is present, and against closed-form known answers.
"""
import itertools

import numpy as np
import pytest

from helpers import sar_like, scaled_err
from oracle import nlm_numpy, ref


def _cases(meta):
    return sorted(meta)


def test_golden_file_has_all_cases(golden):
    z, meta = golden
    assert len(meta) >= 9
    for name in meta:
        for suffix in ("__in", "__out_compiled", "__out_as_written", "__out_patched"):
            assert name + suffix in z.files


def test_c_port_matches_golden_bit_exact(golden, c_oracle):
    """The C restatement reproduces the compiled reference BIT FOR BIT (both semantics, f32 and f64)."""
    z, meta = golden
    for name, m in meta.items():
        a = z[name + "__in"]
        got_c = c_oracle.nlmeans(a, m["r"], m["f"], m["sigma"], m["h"], m["n_eff"], "reference_compiled")
        assert np.array_equal(got_c, z[name + "__out_compiled"]), name
        got_w = c_oracle.nlmeans(a, m["r"], m["f"], m["sigma"], m["h"], m["n_eff"], "as_written")
        assert np.array_equal(got_w, z[name + "__out_patched"]), name
        # the pad+augment+crop construction around the UNMODIFIED kernel agrees to the last bits
        assert scaled_err(got_w, z[name + "__out_as_written"]) < (1e-14 if a.dtype == np.float64 else 2e-7), name


def test_numpy_restatement_matches_golden(golden):
    z, meta = golden
    for name, m in meta.items():
        a = z[name + "__in"]
        tol = 1e-12 if a.dtype == np.float64 else 3e-6
        for sem, key in (("as_written", "__out_as_written"), ("reference_compiled", "__out_compiled")):
            got = nlm_numpy.nlmeans(a, m["r"], m["f"], m["sigma"], m["h"], m["n_eff"], semantics=sem)
            assert scaled_err(got, z[name + key]) < tol, (name, sem)


def test_semantics_differ_only_when_f_positive(golden):
    """with any f_i > 0 the compiled reference ignores the data; with f == 0 both coincide."""
    z, meta = golden
    for name, m in meta.items():
        same = np.array_equal(z[name + "__out_compiled"], z[name + "__out_as_written"])
        if not any(m["f"]):
            assert same, name
    assert not np.array_equal(z["3d_f1_f32__out_compiled"], z["3d_f1_f32__out_as_written"])


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("r,f", [((1, 2, 1), (1, 1, 1)), ((0, 2, 2), (0, 1, 1)), ((2, 1, 0), (2, 2, 0)),
                                 ((1, 1, 1), (0, 0, 0)), ((0, 0, 3), (0, 0, 1))])
def test_c_port_matches_live_reference(c_oracle, dtype, r, f):
    a = sar_like((7, 9, 8, 3), seed=11, dtype=dtype)
    for n_eff in (-1, 4.0):
        assert np.array_equal(c_oracle.nlmeans(a, r, f, 0.3, 1.5, n_eff, "reference_compiled"),
                              ref.reference_compiled(a, r, f, 0.3, 1.5, n_eff))
        assert np.array_equal(c_oracle.nlmeans(a, r, f, 0.3, 1.5, n_eff, "as_written"),
                              ref.as_written_patched(a, r, f, 0.3, 1.5, n_eff))


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_strided_input_same_as_contiguous(c_oracle):
    """The reference indexes through strides (variable-major Dataset views): layout must not matter."""
    a = sar_like((6, 8, 5, 4), seed=3, dtype=np.float32)
    vmajor = np.moveaxis(np.ascontiguousarray(np.moveaxis(a, -1, 0)), 0, -1)
    assert not vmajor.flags.c_contiguous
    o1 = c_oracle.nlmeans(a, (1, 1, 1), (1, 1, 1), 0.3, 0.6)
    o2 = c_oracle.nlmeans(vmajor, (1, 1, 1), (1, 1, 1), 0.3, 0.6)
    assert np.array_equal(o1, o2)
    assert np.array_equal(o1, ref.as_written_patched(vmajor, (1, 1, 1), (1, 1, 1), 0.3, 0.6))


# ---- known answers ------------------------------------------------------------
def test_constant_cube_is_identity(c_oracle):
    a = np.full((5, 6, 4, 2), 3.25, dtype=np.float32)
    out = c_oracle.nlmeans(a, (1, 2, 1), (1, 1, 1), 0.1, 0.2)
    assert np.array_equal(out, a)


def test_huge_sigma_is_reflect_box_mean(c_oracle):
    a = sar_like((6, 7, 5, 3), seed=5, dtype=np.float64)
    r = (1, 2, 1)
    out = c_oracle.nlmeans(a, r, (1, 1, 1), 1e3, 1.0)
    P = np.pad(a, [(k, k) for k in r] + [(0, 0)], mode="reflect")
    box = np.zeros_like(a)
    for t in itertools.product(*[range(2 * k + 1) for k in r]):
        box += P[t[0]:t[0] + 6, t[1]:t[1] + 7, t[2]:t[2] + 5]
    box /= np.prod([2 * k + 1 for k in r])
    assert np.allclose(out, box, rtol=1e-13, atol=0)


def test_step_image_edge_preserved(c_oracle):
    a = np.zeros((1, 12, 12, 1), dtype=np.float64)
    a[:, :, 6:] = 10.0
    out = c_oracle.nlmeans(a, (0, 2, 2), (0, 1, 1), 0.01, 0.05)
    assert np.allclose(out, a, atol=1e-12)


def test_neff_closed_form_when_all_weights_one(c_oracle):
    a = sar_like((5, 6, 5, 2), seed=9, dtype=np.float64)
    r, n = (1, 1, 1), 5.0
    K = 26
    ws = (K + np.sqrt(n * K * K - n * n * K + n * K)) / (n - 1)
    out = c_oracle.nlmeans(a, r, (0, 0, 0), 1e3, 1.0, n_eff=n)
    P = np.pad(a, [(1, 1)] * 3 + [(0, 0)], mode="reflect")
    s = np.zeros_like(a)
    for t in itertools.product(range(3), repeat=3):
        if t != (1, 1, 1):
            s += P[t[0]:t[0] + 5, t[1]:t[1] + 6, t[2]:t[2] + 5]
    assert np.allclose(out, (s + ws * a) / (K + ws), rtol=1e-13)


def test_neff_no_solution_raises(c_oracle):
    a = sar_like((5, 6, 5, 2), seed=9, dtype=np.float32)
    with pytest.raises(ValueError, match="No solution"):
        c_oracle.nlmeans(a, (1, 1, 1), (0, 0, 0), 0.01, 0.01, n_eff=20.0)
    with pytest.raises(ValueError, match="No solution"):
        nlm_numpy.nlmeans(a, (1, 1, 1), (0, 0, 0), 0.01, 0.01, n_eff=20.0)


def test_nan_footprint(c_oracle):
    """One NaN in variable 0, r=(0,2,2): f=0 -> the 5x5 window, all variables; as-written f=1 ->
    (2r+2f+1)^2 = 49 voxels, all variables; compiled f=1 -> 25 voxels in variable 0 only."""
    a = sar_like((1, 15, 15, 3), seed=2, dtype=np.float64)
    a[0, 7, 7, 0] = np.nan
    o = c_oracle.nlmeans(a, (0, 2, 2), (0, 0, 0), 0.3, 0.6)
    assert np.isnan(o).all(-1).sum() == 25 and np.isnan(o).any(-1).sum() == 25
    o = c_oracle.nlmeans(a, (0, 2, 2), (0, 1, 1), 0.3, 0.6, semantics="as_written")
    assert np.isnan(o).all(-1).sum() == 49
    o = c_oracle.nlmeans(a, (0, 2, 2), (0, 1, 1), 0.3, 0.6, semantics="reference_compiled")
    assert np.isnan(o[..., 0]).sum() == 25 and not np.isnan(o[..., 1:]).any()


def test_dtype_generic_and_int_rejected(c_oracle):
    a = sar_like((4, 5, 4, 2), seed=1, dtype=np.float64)
    assert c_oracle.nlmeans(a, (1, 1, 0), (1, 1, 0), 0.3, 0.6).dtype == np.float64
    assert c_oracle.nlmeans(a.astype(np.float32), (1, 1, 0), (1, 1, 0), 0.3, 0.6).dtype == np.float32
    with pytest.raises(TypeError):
        c_oracle.nlmeans((a * 10).astype(np.int32), (1, 1, 0), (1, 1, 0), 0.3, 0.6)
