#!/usr/bin/env python
"""
Generates tests/golden/slc_cfg1.npz: BASELINE configs[0] on the reference's OWN data.

Input : /root/reference/data/slc.data/{C11,C12_real,C12_imag,C22}.img (ENVI BSQ, big-endian float32, 206 x 500;
        data/slc.data/C11.hdr:3-10), stacked as (206, 500, 4) float32 in the order the reference's
        reader yields the variables.
Filter: NLMeansFilter(dims=('y','x'), r=(3,3), f=1) with sigma = 0.01, h = 0.05 (the data are ~1e-3 with bright scatterers up to 1.4; with the
        defaults sigma = h = 1 every weight saturates to 1 -- SURVEY.md 8(d); h is large enough that no voxel's
        largest weight falls below 1e-9, so the float32 reference itself stays in its own domain).
Output: the reference's own compiled kernel (oracle/_ref, built from /root/reference/nd/_filters.pyx) on the array
        laid out as nd/filters.py:447-463 lays it out ((1, y, x, V), r = (0,3,3), f = (0,1,1)):
        `out_compiled` (unmodified binary), `out_as_written` (pad+augment+crop construction through the unmodified
        binary).
Run in the build container (needs /root/reference); bench.py --workload cfg1 and tests read the .npz.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref  # noqa: E402

SRC = "/root/reference/data/slc.data"
NAMES = ["C11", "C12_real", "C12_imag", "C22"]
R, F, SIGMA, H = (0, 3, 3), (0, 1, 1), 0.01, 0.05


def main():
    a = np.stack([np.fromfile(os.path.join(SRC, n + ".img"), dtype=">f4").reshape(206, 500) for n in NAMES],
                 axis=-1).astype(np.float32)
    arr = np.ascontiguousarray(a[None])                       # (1, y, x, V)
    out_c = ref.reference_compiled(arr, R, F, SIGMA, H)
    out_w = ref.as_written(arr, R, F, SIGMA, H)
    out_p = ref.as_written_patched(arr, R, F, SIGMA, H)
    assert np.array_equal(out_w, out_p), "F3 construction and the patched copy must agree bitwise for float32"
    np.savez_compressed(os.path.join(HERE, "slc_cfg1.npz"), input=a, out_compiled=out_c[0], out_as_written=out_w[0],
                        r=np.array(R), f=np.array(F), sigma=SIGMA, h=H, names=np.array(NAMES))
    print("wrote slc_cfg1.npz", a.shape, "weights non-trivial:", float(np.abs(out_w - out_c).max()))


if __name__ == "__main__":
    main()
