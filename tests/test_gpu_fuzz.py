"""
Seeded sweep over random small configurations (shapes around the tile edges, 1..8 variables, radii 0..4, patch
radii 0..2 set like `NLMeansFilter.__init__` does -- f_i = f where r_i > 0 --, both dtypes, both semantics, with and
without n_eff): whatever kernel the plan picks (tiled float32 / float64 / half-group / small 2-D CTAs / box mean /
generic) must agree with the C oracle.  Complements the hand-picked cases of test_gpu_parity.py.
"""
import numpy as np
import pytest

from helpers import sar_like, scaled_err

pytestmark = pytest.mark.gpu


def _configs(n, seed):
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        ndim_f = int(rng.integers(1, 4))                       # number of filtered axes
        axes = sorted(rng.choice(3, size=ndim_f, replace=False).tolist())
        r = [0, 0, 0]
        for a in axes:
            r[a] = int(rng.integers(1, 5))
        fval = int(rng.integers(0, 3))
        f = [fval if r[a] > 0 else 0 for a in range(3)]
        shape = [int(rng.integers(1, 7)) if r[a] == 0 else int(rng.integers(r[a] + f[a] + 1, r[a] + f[a] + 36)) for a in range(3)]
        V = int(rng.integers(1, 9))
        K = np.prod([2 * x + 1 for x in r]) - 1
        patch = np.prod([2 * x + 1 for x in f])
        if np.prod(shape) * K * patch * V > 4e7:              # keep the oracle to a fraction of a second
            continue
        dtype = np.float64 if rng.random() < 0.35 else np.float32
        sem = "reference_compiled" if rng.random() < 0.25 else "as_written"
        n_eff = float(rng.integers(3, 9)) if rng.random() < 0.2 else -1.0
        out.append((tuple(shape) + (V,), tuple(r), tuple(f), dtype, sem, n_eff))
    return out


@pytest.mark.parametrize("cfg", _configs(48, seed=2026) + _configs(48, seed=7), ids=lambda c: "%s-r%s-f%d-%s-%s%s" % (
    "x".join(map(str, c[0])), "".join(map(str, c[1])), max(c[2]), np.dtype(c[3]).name, c[4][:2], "-neff" if c[5] >= 0 else ""))
def test_random_configuration_matches_oracle(cfg):
    import torch
    from nd_b200 import device
    from oracle import c_port
    shape, r, f, dtype, sem, n_eff = cfg
    a = sar_like(shape, seed=sum(shape) + 7 * sum(r), dtype=dtype)
    h = 1.5 if n_eff >= 0 else 0.6                              # wide weights so that find_weight has a solution
    try:
        ref = c_port.nlmeans(a, r, f, 0.3, h, n_eff, sem, threads=8)
    except ValueError:                                          # 'No solution' in the oracle: the GPU must say so too
        with pytest.raises(ValueError, match="No solution"):
            device.Plan(a.shape, r, f, 0.3, h, n_eff, semantics=sem, dtype=dtype).apply(torch.from_numpy(a).cuda())
        return
    plan = device.Plan(a.shape, r, f, 0.3, h, n_eff, semantics=sem, dtype=dtype)
    out = plan.apply(torch.from_numpy(a).cuda()).cpu().numpy()
    tol = 1e-12 if dtype == np.float64 else (1e-5 if "boxmean" in plan.kernel_name else 1e-4)
    assert out.dtype == dtype and np.isfinite(out).all()
    assert scaled_err(out, ref) < tol, (plan.kernel_name, scaled_err(out, ref))
