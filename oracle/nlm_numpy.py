"""
TEST INFRASTRUCTURE ONLY.  Independent float64 NumPy restatement of the reference NLM
filter in the box-sum form the CUDA kernels use.

Follows nd/_filters.pyx:351-420:
  * reflect without edge repeat (`_idx`, nd/_filters.pyx:34-40) == np.pad(mode='reflect');
    because it is applied to the SUM p+d / q+d (:378-384) the whole filter equals
    "reflect-pad by r+f, then filter the interior".
  * d^2 = sum over patch and variables of squared differences / (V * prod(2f+1))  (:337, :372-388)
  * w = exp(-max(d^2 - 2 sigma^2, 0) / h^2)                                        (:391)
  * self weight = max w (1 if all 0) or find_weight(S, Q, n_eff)                   (:406-413, :297-314)
  * out = (sum w a_q + w_self a_p) / (sum w + w_self)                              (:415-420)
`semantics='reference_compiled'` reproduces the LP64 binary: if any f_i > 0, d^2 == 0.
"""
import itertools

import numpy as np


def _boxsum(s, f):
    for ax in range(3):
        k = f[ax]
        if k == 0:
            continue
        n = s.shape[ax] - 2 * k
        acc = np.zeros([n if a == ax else s.shape[a] for a in range(3)], dtype=s.dtype)
        for d in range(2 * k + 1):
            sl = [slice(None)] * 3
            sl[ax] = slice(d, d + n)
            acc += s[tuple(sl)]
        s = acc
    return s


def nlmeans(arr, r, f, sigma, h, n_eff=-1, semantics="as_written", dtype=np.float64):
    a = np.asarray(arr, dtype=dtype)
    r = [int(x) for x in r]
    f = [int(x) for x in f]
    N, V = a.shape[:3], a.shape[3]
    if semantics == "reference_compiled" and any(f):
        zero_dist = True
    else:
        zero_dist = False
    pad = [r[i] + f[i] for i in range(3)]
    P = np.pad(a, [(p, p) for p in pad] + [(0, 0)], mode="reflect")
    norm = V * np.prod([2 * k + 1 for k in f])

    def win(off, halo):
        return P[tuple(slice(pad[i] + off[i] - halo[i], pad[i] + off[i] + N[i] + halo[i])
                       for i in range(3))]

    A = win((0, 0, 0), f)
    S = np.zeros(N, dtype=dtype)
    Q = np.zeros(N, dtype=dtype)
    M = np.zeros(N, dtype=dtype)
    acc = np.zeros(N + (V,), dtype=dtype)
    for t in itertools.product(*[range(-k, k + 1) for k in r]):
        if t == (0, 0, 0):
            continue
        if zero_dist:
            d2 = np.zeros(N, dtype=dtype)
        else:
            B = win(t, f)
            d2 = _boxsum(((A - B) ** 2).sum(-1), f) / norm
        x = d2 - 2 * sigma ** 2
        x = np.where(0 > x, 0, x)          # the reference's `max` (NaN propagates)
        w = np.exp(-x / h ** 2)
        S += w
        Q += w * w
        M = np.where(w > M, w, M)
        acc += w[..., None] * win(t, (0, 0, 0))
    if n_eff < 0:
        ws = np.where(M == 0, 1.0, M)
    else:
        with np.errstate(invalid="ignore", divide="ignore"):
            if np.any(n_eff - 1 > S ** 2 / Q):
                raise ValueError("No solution")
            ws = (S + np.sqrt(n_eff * S * S - n_eff * n_eff * Q + n_eff * Q)) / (n_eff - 1)
    return (acc + ws[..., None] * a) / (S + ws)[..., None]
