"""
TEST INFRASTRUCTURE ONLY -- not product code.

CPU oracles for the non-local-means hot path of jnhansen/nd
(`nd/_filters.pyx::_pixelwise_nlmeans_3d`).  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference`
legs may import anything from this package; `nd_b200` (the product) never does.

Three oracles, in decreasing order of authority:

1. `oracle.ref`       -- the reference's OWN Cython kernel, compiled unmodified from
                         `/root/reference/nd/_filters.pyx` into `oracle/_ref/` by
                         `oracle/build_ref.py` (git-ignored, travels to the GPU box).
                         Pins both semantics (see `oracle.ref` docstring):
                         `reference_compiled` by a direct call and `as_written` through the
                         pad+augment+crop construction.
2. `oracle.c_port`    -- a plain-C restatement (`oracle/nlm_oracle.c`), pinned against (1)
                         by `tests/test_oracle.py` and by the committed golden vectors in
                         `tests/golden/`.  Fast enough for the CPU-baseline timing.
3. `oracle.nlm_numpy` -- an independent float64 NumPy restatement in the box-sum form the
                         CUDA kernels use.

Parity status: PINNED -- every oracle is checked against outputs of the unmodified
reference kernel run in this container (`tests/golden/make_golden.py`).

Oracles of the rows built after the hot path:

4. `oracle.ndimage_numpy`  -- NumPy restatement of the scipy.ndimage arithmetic below the sibling
                              filters (row N2).  PINNED: bit-exact against scipy itself on the CPU
                              (`tests/test_sibling_filters.py`); the GPU tests then use scipy directly.
5. `oracle.omnibus_oracle` -- NumPy restatement of `nd/_change.pyx` (row N4).  PINNED: `oracle.ref_change` wraps the
                              reference's own `_change.pyx`, compiled unmodified into `oracle/_ref/_change` with
                              `oracle/gsl_shim` standing in for its one GSL call (chi-square CDF; checked against
                              published table values); golden vectors `tests/golden/change_golden.npz`.
"""
