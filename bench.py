#!/usr/bin/env python
"""
bench.py -- NLMeansFilter throughput on B200, one JSON line (driver contract).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg3|cfg2|cfg1]
  N > 1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

Workload (default cfg3 = BASELINE.json configs[2], the headline single-GPU configuration):
  synthetic complex-SAR-like cube 4096 x 4096 x 32, 4 float32 variables, NLMeansFilter(dims=('y','x','time'),
  r=(5,5,2), f=1, sigma=0.25, h=0.5), semantics 'as_written'.  With N GPUs the cube is 4096*N rows tall
  (weak scaling): rank k owns rows [4096 k, 4096 (k+1)), halo rows r_y+f_y = 6 are exchanged with NCCL
  send/recv before every apply; there is no other collective on the data path.

A "step" = one pass of the hot path over the (per-rank) cube: stage (reflect-pad) -> halo exchange ->
nlm kernel -> unstage, inputs resident in HBM.  `value` = voxels of all ranks / max-over-ranks time.
`e2e` = the same through the reference-facing entry point `_pixelwise_nlmeans_3d` on pinned HOST
arrays (H2D + kernels + D2H inside the timed region).  `roofline` is for the dominant kernel
(nlm_tiled): algorithmic FP32 flops / CUDA-event time of its launches.
`cpu_baseline` times the reference's own Cython kernel on the host cores on a bounded crop.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (ny_per_gpu, nx, nt, V, r(y,x,t), f, sigma, h, text)
    "cfg3": (4096, 4096, 32, 4, (5, 5, 2), 1, 0.25, 0.5,
             "synthetic SAR-like cube 4096x4096x32 per GPU, 4 float32 variables, r=(5,5,2), f=1 (BASELINE configs[2])"),
    "cfg2": (1024, 1024, 24, 4, (5, 5, 1), 1, 0.25, 0.5,
             "synthetic stand-in 1024x1024x24 for the absent s1_kalimantan file, 4 float32 variables, r=(5,5,1), f=1 (BASELINE configs[1])"),
    "cfg1": (206, 500, 1, 4, (3, 3, 0), 1, 0.25, 0.5,
             "synthetic 206x500 image (shape of data/slc.nc), 4 float32 variables, r=(3,3), f=1 (BASELINE configs[0])"),
}
METRIC = "NLMeansFilter Mvoxel/s"
UNIT = "Mvoxel/s"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": float(p.get("hbm_gbs", 6650.0)), "sm_max_mhz": float(p.get("sm_max_mhz", 1965.0)),
                "source": "MEASURED_PEAKS.json"}
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0, "source": "fallback (B200_PROFILING.md)"}


# ------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            pass
        try:
            rows = [l.strip().split(", ") for l in open(self.path) if l.strip()]
            os.unlink(self.path)
            sm = sorted(float(r[1]) for r in rows if len(r) >= 8)
            if sm:
                out["sm_mhz"] = sm[len(sm) // 2]
                out["sm_max_mhz"] = float(rows[0][2])
                out["power_w_max"] = max(float(r[3]) for r in rows if len(r) >= 8)
                out["samples"] = len(sm)
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                for i, nm in enumerate(names):
                    if any(r[4 + i].strip().lower() == "active" for r in rows if len(r) >= 8):
                        out["reasons"].append(nm)
        except Exception:
            pass
        return out


# ------------------------------------------------------------------------------------------------
# CPU baseline: the reference's own Cython kernel on the host cores, bounded crop
# ------------------------------------------------------------------------------------------------
def _cpu_crop_job(args):
    """Worker: filter one independent crop with the reference kernel."""
    crop, r, f, sigma, h, which = args
    from oracle import ref
    fn = ref.as_written_patched if which == "as_written" else ref.reference_compiled
    return float(fn(crop, r, f, sigma, h, -1).sum())


def cpu_reference_rate(crop, r, f, sigma, h, which, procs):
    """voxels/s of the reference kernel with `procs` processes, each filtering its own copy-sized crop
    (no overlapping halos between workers: the most favourable way to use all host cores; the
    reference's own njobs mechanism, nd/utils.py:343-401, re-reads r+f buffer rows per worker)."""
    import multiprocessing as mp
    jobs = [(crop, r, f, sigma, h, which)] * procs
    t0 = time.perf_counter()
    if procs == 1:
        sums = [_cpu_crop_job(jobs[0])]
    else:
        with mp.get_context("fork").Pool(procs) as pool:
            sums = pool.map(_cpu_crop_job, jobs)
    dt = time.perf_counter() - t0
    return procs * crop.shape[0] * crop.shape[1] * crop.shape[2] / dt, dt, sums


def make_cpu_crop(wl, rows, cols):
    """A (rows, cols, nt, V) float32 crop with the statistics of the synthetic cube (NumPy generator; the
    per-voxel cost of the reference does not depend on the values)."""
    import numpy as np
    ny, nx, nt, V, r, f, sigma, h, _ = WORKLOADS[wl]
    rng = np.random.default_rng(42)
    a = rng.gamma(4.0, 0.25, size=(rows, cols, nt, V)).astype(np.float32)
    a[..., 1:3] = rng.normal(0, 0.3, size=a[..., 1:3].shape).astype(np.float32)
    return a


def cpu_baseline(wl, target_seconds=12.0):
    import numpy as np
    from oracle import build_ref, ref
    ny, nx, nt, V, r, f, sigma, h, _ = WORKLOADS[wl]
    fv = tuple(f if x > 0 else 0 for x in r)
    cores = os.cpu_count() or 1
    if not build_ref.built():
        # no compiled reference on this box: fall back to the C port of the oracle (kind "port")
        from oracle import c_port
        crop = make_cpu_crop(wl, 32, 32)
        t0 = time.perf_counter()
        c_port.nlmeans(crop, r, fv, sigma, h, -1, "as_written", threads=cores)
        dt = time.perf_counter() - t0
        return {"value": crop[..., 0].size / dt / 1e6, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": "32x32x%d crop, C port of the oracle with OpenMP" % nt}
    # calibrate on a small crop, then size the per-core crop for ~target_seconds
    halo = r[0] + fv[0]
    cal = make_cpu_crop(wl, max(halo + 2, 8), 16)
    rate1, _, _ = cpu_reference_rate(cal, r, fv, sigma, h, "as_written", 1)
    cols = 32 if nx >= 32 else nx
    rows = int(max(halo + 1, min(ny, rate1 * target_seconds / (cols * nt))))
    crop = make_cpu_crop(wl, rows, cols)
    rate, dt, _ = cpu_reference_rate(crop, r, fv, sigma, h, "as_written", cores)
    rate_c, dt_c, _ = cpu_reference_rate(crop, r, fv, sigma, h, "reference_compiled", cores)
    return {"value": rate / 1e6, "unit": UNIT, "cores": cores, "kind": "reference",
            "sample": "%d independent %dx%dx%d crops (V=%d, same r/f/sigma/h as the workload), one per host core; "
                      "reference .pyx with the three SIZE_TYPE casts = as-written semantics (%.1f s)"
                      % (cores, rows, cols, nt, V, dt),
            "single_core_value": rate1 / 1e6,
            "unmodified_binary_value": rate_c / 1e6,
            "unmodified_binary_note": "unmodified reference binary on the same crop: because of the unsigned-f bug "
                                      "(SURVEY F1) it computes a box mean, not patch distances (%.1f s)" % dt_c}


# ------------------------------------------------------------------------------------------------
# reference arm
# ------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np
    from oracle import build_ref
    wl = args.workload
    ny, nx, nt, V, r, f, sigma, h, text = WORKLOADS[wl]
    fv = tuple(f if x > 0 else 0 for x in r)
    cores = os.cpu_count() or 1
    kind = "reference" if build_ref.built() else "port"
    halo = r[0] + fv[0]
    if kind == "reference":
        cal = make_cpu_crop(wl, max(halo + 2, 8), 16)
        rate1, _, _ = cpu_reference_rate(cal, r, fv, sigma, h, "as_written", 1)
    else:
        rate1 = 1000.0
    per_step_s = max(2.0, min(8.0, 150.0 / max(args.steps + args.warmup, 1)))
    cols = 32 if nx >= 32 else nx
    rows = int(max(halo + 1, min(ny, rate1 * per_step_s / (cols * nt))))
    crop = make_cpu_crop(wl, rows, cols)
    times = []
    for it in range(args.warmup + args.steps):
        if kind == "reference":
            _, dt, _ = cpu_reference_rate(crop, r, fv, sigma, h, "as_written", cores)
        else:
            from oracle import c_port
            t0 = time.perf_counter()
            c_port.nlmeans(crop, r, fv, sigma, h, -1, "as_written", threads=cores)
            dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    total = sum(times)
    vox = rows * cols * nt * (cores if kind == "reference" else 1)
    value = vox * len(times) / total / 1e6
    sample = ("%s %dx%dx%d crops of the workload (V=%d, same r/f/sigma/h), %s"
              % ("%d independent" % cores if kind == "reference" else "one", rows, cols, nt, V,
                 "one per host core, reference nd/_filters.pyx (as-written: three SIZE_TYPE casts)" if kind == "reference"
                 else "C port of the oracle with OpenMP on %d threads" % cores))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl, "description": text, "semantics": "as_written"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def bind_to_gpu_cpus(gpu_index):
    """One process per GPU: run on the CPUs that are local to this rank's GPU (NVML's affinity mask), so that the
    pinned host chunk of the e2e leg is allocated on the NUMA node its PCIe link hangs off.  Returns the CPU count
    or None when NVML / the affinity call is unavailable (nothing changes then)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from nd_b200 import device
    from nd_b200._filters import _pixelwise_nlmeans_3d
    from nd_b200.shard import DistributedShard

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; there is no CPU fallback")
    # stdout carries exactly ONE JSON line: whatever libraries print on fd 1 meanwhile (NCCL announces its version
    # there) is sent to stderr, and the line is written to the real stdout at the end.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    numa = bind_to_gpu_cpus(local_rank) if world > 1 else None
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    wl = args.workload
    ny, nx, nt, V, r, f, sigma, h, text = WORKLOADS[wl]
    if args.rows:
        ny = args.rows
    fv = tuple(f if x > 0 else 0 for x in r)
    shape = (ny, nx, nt, V)
    pk = peaks()

    # per-rank slab of the global cube (rows [rank*ny, (rank+1)*ny)), generated on the device
    cube = device.synth_cube(ny, nx, nt, V, y_offset=rank * ny, seed=42, device=dev)
    out = torch.empty_like(cube)
    plan = device.Plan(shape, r, fv, sigma, h, -1, semantics="as_written")
    shard = DistributedShard(plan, axis=0, rank=rank, world=world)
    info = plan.describe()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(ev=None):
        shard.stage(cube)
        shard.exchange()
        if ev is not None:
            ev[0].record()
        shard.run()
        if ev is not None:
            ev[1].record()
        shard.unstage(out)

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = device.launch_count()
    barrier()
    e0.record()
    for i in range(args.steps):
        step(kev[i])
    e1.record()
    barrier()
    launches = device.launch_count() - launches0
    ms_total = e0.elapsed_time(e1)
    kernel_ms = sum(a.elapsed_time(b) for a, b in kev) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    flag = int(shard.flag.item())

    t = torch.tensor([ms_total, kernel_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, kernel_ms = float(t[0]), float(t[1])
    ms_per_step = ms_total / args.steps
    vox_rank = ny * nx * nt
    value = world * vox_rank / (ms_per_step * 1e-3) / 1e6

    # ---- e2e: the reference-facing entry point on pinned HOST arrays, every step H2D + kernels + D2H ----
    e2e = None
    if not args.no_e2e:
        del shard, out
        e2e_rows = ny
        nbytes = vox_rank * V * 4
        try:
            import psutil
            avail = psutil.virtual_memory().available
            if 2.2 * nbytes * max(world, 1) > 0.6 * avail:
                e2e_rows = max(64, int(ny * 0.6 * avail / (2.2 * nbytes * world)))
        except Exception:
            pass
        # N > 1: like the reference's own njobs mechanism (xr_split, nd/utils.py:305-310) every rank's host
        # chunk carries a buffer of r+f rows of its neighbours; only the interior rows count as work.
        halo = r[0] + fv[0]
        lo_buf = halo if rank > 0 else 0
        hi_buf = halo if rank < world - 1 else 0
        tot_rows = e2e_rows + lo_buf + hi_buf
        h_in = torch.empty((tot_rows, nx, nt, V), dtype=torch.float32, pin_memory=True)
        h_out = torch.empty((tot_rows, nx, nt, V), dtype=torch.float32, pin_memory=True)
        del cube
        torch.cuda.empty_cache()
        src = device.synth_cube(tot_rows, nx, nt, V, y_offset=rank * ny - lo_buf, seed=42, device=dev)
        h_in.copy_(src)
        del src
        a_in, a_out = h_in.numpy(), h_out.numpy()
        r3 = np.array(r, dtype=np.uint32)
        f3 = np.array(fv, dtype=np.uint32)
        e2e_steps = max(1, min(args.steps, 3))
        _pixelwise_nlmeans_3d(a_in, a_out, r3, f3, sigma, h, -1, semantics="as_written")      # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            _pixelwise_nlmeans_3d(a_in, a_out, r3, f3, sigma, h, -1, semantics="as_written")
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt[0])
        ebytes = tot_rows * nx * nt * V * 4
        e2e = {"value": world * e2e_rows * nx * nt * e2e_steps / dt / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": ebytes, "d2h_bytes_per_step": ebytes, "rows_per_gpu": e2e_rows,
               "steps": e2e_steps, "api": "nd_b200._filters._pixelwise_nlmeans_3d(host arr, host output, r, f, sigma, h, n_eff)",
               "cpus_bound_to_gpu": numa,
               "note": "rank-local host chunk incl. r+f buffer rows of its neighbours (the reference's xr_split rule); "
                       "slab-pipelined H2D / kernels / D2H on three streams; only interior rows are counted",
               "checksum": float(np.float64(a_out[lo_buf:lo_buf + e2e_rows:max(1, e2e_rows // 64)].sum()))}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel ----
    fp32_nominal = 148 * 128 * 2 * pk["sm_max_mhz"] * 1e6 / 1e12           # TFLOP/s, SMs*lanes*2*clock
    try:
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        fp32_nominal = sms * 128 * 2 * pk["sm_max_mhz"] * 1e6 / 1e12
    except Exception:
        pass
    achieved = plan.flops_per_voxel * vox_rank / (kernel_ms * 1e-3) / 1e12
    fp32_measured = device.measure_fp32_peak(0.5)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            if tj.get("workload") == wl and tj.get("rows") == ny:
                traffic = tj.get("dram_bytes_per_launch")
        except Exception:
            pass
    alg_bytes = 8.0 * V * vox_rank
    roofline = {"bound": "fp32_fma", "kernel": plan.kernel_name, "achieved": achieved, "peak": fp32_nominal,
                "unit": "TFLOP/s", "frac": achieved / fp32_nominal, "traffic": traffic,
                "peak_source": "SMs*128*2*sm_max_mhz (%s): nominal FP32 FMA peak, the binding roofline of this path "
                               "(SURVEY.md 8(d)); no tensor cores" % pk["source"],
                "kernel_ms": kernel_ms, "flops_per_voxel": plan.flops_per_voxel,
                "fp32_fma_peak_measured_same_run": fp32_measured, "frac_of_measured_fma_peak": achieved / fp32_measured,
                "hbm": {"algorithmic_bytes_per_launch": alg_bytes, "achieved_gbs": alg_bytes / (kernel_ms * 1e-3) / 1e9,
                        "peak_gbs": pk["hbm_gbs"], "frac": alg_bytes / (kernel_ms * 1e-3) / 1e9 / pk["hbm_gbs"]}}

    cpu = None
    if world == 1 and not args.no_cpu:
        try:
            cpu = cpu_baseline(wl)
        except Exception as e:                                   # never lose the GPU line to the CPU leg
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "unavailable", "sample": repr(e)}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl, "description": text, "rows_per_gpu": ny, "global_shape": [ny * world, nx, nt, V],
                       "r": list(r), "f": list(fv), "sigma": sigma, "h": h, "semantics": "as_written",
                       "sharding": "y-sharded, halo r_y+f_y=%d rows, NCCL send/recv" % (r[0] + fv[0]) if world > 1 else "single GPU",
                       "l2": "inputs (%.1f GB per GPU) are larger than L2, no flush needed" % (alg_bytes / 2e9),
                       "plan": info},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "no_solution_flag": flag}
    sys.stdout.flush()
    os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--rows", type=int, default=0, help="override rows per GPU (development only)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
