#!/usr/bin/env python
"""Development parity sweep (GPU): libndnlm vs the oracles on small seeded cubes.
Usage: dev_parity.py            -> runs every case, each in its own subprocess (a CUDA fault cannot poison the rest)
       dev_parity.py --case I   -> runs one case in-process"""
import json, os, subprocess, sys, time
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CASES = [
    # name, shape, r, f, dtype, n_eff, kernel, env
    ("3d_f1_small",      (20, 45, 14, 4), (2, 3, 1), (1, 1, 1), "f4", -1, "tiled", {}),
    ("3d_f1_small_ldg",  (20, 45, 14, 4), (2, 3, 1), (1, 1, 1), "f4", -1, "tiled", {"NDNLM_LOADER": "ldg"}),
    ("3d_f1_cfg3like",   (30, 70, 16, 4), (5, 5, 2), (1, 1, 1), "f4", -1, "tiled", {}),
    ("3d_f1_generic",    (12, 20, 9, 4),  (2, 3, 1), (1, 1, 1), "f4", -1, "generic", {}),
    ("2d_f1",            (1, 40, 70, 4),  (0, 3, 3), (0, 1, 1), "f4", -1, "tiled", {}),
    ("2d_batch_t",       (20, 33, 5, 4),  (2, 2, 0), (1, 1, 0), "f4", -1, "tiled", {}),
    ("3d_f0",            (10, 20, 15, 4), (1, 2, 1), (0, 0, 0), "f4", -1, "tiled", {}),
    ("1d_f1_V2",         (5, 6, 50, 2),   (0, 0, 4), (0, 0, 1), "f4", -1, "tiled", {}),
    ("3d_f1_V3",         (14, 37, 8, 3),  (1, 2, 2), (1, 1, 1), "f4", -1, "tiled", {}),
    ("3d_f1_V1",         (14, 37, 8, 1),  (1, 2, 2), (1, 1, 1), "f4", -1, "tiled", {}),
    ("3d_f1_f64",        (10, 16, 7, 4),  (1, 2, 1), (1, 1, 1), "f8", -1, "auto", {}),
    ("3d_f1_neff",       (10, 16, 7, 4),  (2, 2, 1), (1, 1, 1), "f4", 6.0, "auto", {}),
    ("3d_f2_generic",    (12, 18, 9, 4),  (2, 2, 1), (2, 2, 2), "f4", -1, "auto", {}),
    ("3d_compiled_sem",  (12, 18, 9, 4),  (2, 2, 1), (1, 1, 1), "f4", -1, "auto", {"ND_NLM_SEMANTICS": "reference_compiled"}),
    ("3d_f1_tall",       (64, 100, 32, 4), (3, 3, 2), (1, 1, 1), "f4", -1, "tiled", {}),
]


def make_data(shape, dtype, seed=0):
    rng = np.random.default_rng(seed)
    a = rng.gamma(4.0, 0.25, size=shape)
    a[..., 1::2] = rng.normal(0, 0.3, size=a[..., 1::2].shape)
    a *= (1.0 + (np.arange(shape[1]) // 16 % 3)[None, :, None, None] * 0.5)
    return a.astype(dtype)


def scaled_err(out, ref):
    errs = []
    for v in range(ref.shape[-1]):
        d = np.abs(out[..., v].astype(np.float64) - ref[..., v].astype(np.float64))
        errs.append(float(np.nanmax(d) / max(np.nanmax(np.abs(ref[..., v])), 1e-30)))
    return max(errs)


def run_case(i):
    import torch
    from nd_b200 import device
    from oracle import ref as oref, nlm_numpy
    name, shape, r, f, dt, n_eff, kernel, env = CASES[i]
    os.environ.update(env)
    sem = os.environ.get("ND_NLM_SEMANTICS", "as_written")
    a = make_data(shape, np.dtype(dt))
    sigma, h = 0.3, 0.6
    plan = device.Plan(shape, r, f, sigma, h, n_eff, dtype=a.dtype, kernel=kernel)
    t = torch.from_numpy(a).cuda()
    t0 = time.time()
    out = plan.apply(t)
    torch.cuda.synchronize()
    ms = (time.time() - t0) * 1e3
    o = out.cpu().numpy()
    if sem == "reference_compiled":
        ref = oref.reference_compiled(a, r, f, sigma, h, n_eff)
    else:
        ref = oref.as_written(a, r, f, sigma, h, n_eff)
    ref64 = nlm_numpy.nlmeans(a, r, f, sigma, h, n_eff, semantics=sem)
    res = {"case": name, "kernel": plan.kernel_name, "grid": plan.info.grid, "smem": plan.info.smem_bytes,
           "tile": list(plan.info.tile), "roles": list(plan.info.role_axis),
           "err_vs_reference": scaled_err(o, ref), "err_vs_f64": scaled_err(o, ref64),
           "ref_vs_f64": scaled_err(ref, ref64), "nan": int(np.isnan(o).sum()), "ms": round(ms, 2)}
    print("RESULT " + json.dumps(res), flush=True)


if __name__ == "__main__":
    if "--case" in sys.argv:
        run_case(int(sys.argv[sys.argv.index("--case") + 1]))
    elif "--inproc" in sys.argv:      # all cases in this process (fast; a CUDA fault kills the sweep)
        for i in range(len(CASES)):
            env_keys = list(CASES[i][7])
            try:
                run_case(i)
            except Exception as e:
                print("FAILED case %d %s: %r" % (i, CASES[i][0], e), flush=True)
            for k in env_keys:
                os.environ.pop(k, None)
    else:
        for i in range(len(CASES)):
            p = subprocess.run([sys.executable, __file__, "--case", str(i)], capture_output=True, text=True, timeout=300)
            lines = [l for l in p.stdout.splitlines() if l.startswith("RESULT")]
            if p.returncode == 0 and lines:
                print(lines[0])
            else:
                print("FAILED case %d %s rc=%d\n%s\n%s" % (i, CASES[i][0], p.returncode, p.stdout[-1500:], p.stderr[-3000:]))
