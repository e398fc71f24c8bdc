"""
Device-side driver of the NLM hot path: a thin object layer over the C ABI (include/ndnlm.h).
PyTorch is used only for device buffers and streams; every kernel lives in libndnlm.so.
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib


def default_semantics():
    """'as_written' unless overridden by ND_NLM_SEMANTICS."""
    return os.environ.get("ND_NLM_SEMANTICS", "as_written")


class Plan:
    """One (shape, r, f, sigma, h, n_eff, semantics, dtype) configuration (ndnlm_plan_create)."""

    def __init__(self, shape, r, f, sigma, h, n_eff=-1, semantics=None, dtype=np.float32, kernel="auto", roles=None):
        dtype = np.dtype(dtype)
        if dtype == np.float32:
            code = _lib.F32
        elif dtype == np.float64:
            code = _lib.F64
        else:
            raise TypeError("No matching signature found (dtype %s; only float32/float64)" % dtype)
        semantics = semantics or default_semantics()
        if semantics not in _lib.SEMANTICS:
            raise ValueError("semantics must be one of %s" % sorted(_lib.SEMANTICS))
        self.shape = tuple(int(s) for s in shape)
        if len(self.shape) != 4:
            raise ValueError("shape must be (N0, N1, N2, V)")
        self.r = tuple(int(x) for x in r)
        self.f = tuple(int(x) for x in f)
        if len(self.r) != 3 or len(self.f) != 3:
            raise ValueError("Buffer has wrong number of dimensions (expected 3 radii)")
        self.dtype = dtype
        self.torch_dtype = torch.float32 if dtype == np.float32 else torch.float64
        self.semantics = semantics
        L = _lib.lib()
        self._L = L
        self._h = ctypes.c_void_p()
        rc = -1
        # roles = (axis of W, axis of R, axis of X) taken from the plan of the WHOLE array when this plan is one of
        # its shards (ndnlm_plan_create_roles); None = chosen from the shape
        role_arg = (ctypes.c_int32 * 3)(*[int(a) for a in roles]) if roles is not None else None
        if dtype == np.float64 and kernel == "auto" and os.environ.get("ND_NLM_FLOAT64_COMPUTE", "").lower() in ("float32", "fp32", "f32"):
            # opt-in: float64 data through the fp32 tiled kernel (north_star's fp32 compute, ~1e-6 from the float64
            # result, ~100x the generic float64 kernel); configurations without a tiled instantiation stay generic
            rc = L.ndnlm_plan_create_roles(ctypes.byref(self._h), _lib.i64(self.shape), _lib.u32(self.r), _lib.u32(self.f),
                                           float(sigma), float(h), float(n_eff), _lib.SEMANTICS[semantics], code,
                                           _lib.KERNELS["tiled"], role_arg)
        if rc != 0:
            _lib.check(L.ndnlm_plan_create_roles(ctypes.byref(self._h), _lib.i64(self.shape), _lib.u32(self.r),
                                                 _lib.u32(self.f), float(sigma), float(h), float(n_eff),
                                                 _lib.SEMANTICS[semantics], code, _lib.KERNELS[kernel], role_arg))
        info = _lib.Info()
        _lib.check(L.ndnlm_plan_info(self._h, ctypes.byref(info)))
        self.info = info
        self.kernel_name = info.kernel_name.decode()
        self.is_tiled = info.kernel == _lib.KERNEL_TILED
        self.roles = tuple(int(a) for a in info.role_axis)
        # the `kernel=` request that reproduces this plan's choice for another shape (the shards of one array)
        self.kernel_request = ("generic" if not self.is_tiled else
                               "tiled64" if (dtype == np.float64 and int(info.elem_bytes) == 8) else "tiled")
        self.is_tiled = info.kernel == _lib.KERNEL_TILED
        self.padded_bytes = int(info.padded_bytes)
        self.out_bytes = int(info.out_bytes)
        self.flops_per_voxel = float(info.flops_per_voxel)
        self.voxels = int(info.voxels)
        self.n_offsets = int(info.n_offsets)
        self.scratch_bytes = int(L.ndnlm_scratch_bytes(self._h))
        self._scratch = None

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None and self._h.value:
                self._L.ndnlm_plan_destroy(self._h)
                self._h = ctypes.c_void_p()
        except Exception:
            pass

    # ---- buffers -------------------------------------------------------------------------
    def new_padded(self, device):
        return torch.empty(self.padded_bytes, dtype=torch.uint8, device=device)

    def new_internal_out(self, device):
        return torch.empty(self.out_bytes, dtype=torch.uint8, device=device)

    def describe(self):
        i = self.info
        return {
            "kernel": self.kernel_name, "role_axis_WRX": list(i.role_axis), "n_WRX": list(i.n), "pad_WRX": list(i.pad),
            "tile_WRX": list(i.tile), "box_WRX": list(i.box), "warps_WRX": list(i.warps), "threads": i.threads,
            "grid": i.grid, "smem_bytes": i.smem_bytes, "n_offsets": self.n_offsets,
            "flops_per_voxel": self.flops_per_voxel, "padded_bytes": self.padded_bytes, "out_bytes": self.out_bytes,
        }

    # ---- the four steps ------------------------------------------------------------------
    @staticmethod
    def _stream(t=None):
        """Current torch stream of the device that owns tensor `t` (or of the current device)."""
        dev = t.device if t is not None else None
        return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def _check_arr(self, t):
        if not t.is_cuda:
            raise ValueError("expected a CUDA tensor")
        if tuple(t.shape) != self.shape:
            raise ValueError("array shape %s does not match the plan %s" % (tuple(t.shape), self.shape))
        if t.dtype != self.torch_dtype:
            raise TypeError("array dtype %s does not match the plan %s" % (t.dtype, self.torch_dtype))

    def stage(self, arr, padded, shard_axis=-1, lo_edge="reflect", hi_edge="reflect"):
        self._check_arr(arr)
        edge = {"reflect": _lib.EDGE_REFLECT, "halo": _lib.EDGE_HALO, "source": _lib.EDGE_SOURCE}
        _lib.check(self._L.ndnlm_stage(self._h, ctypes.c_void_p(arr.data_ptr()), _lib.i64(arr.stride()),
                                       ctypes.c_void_p(padded.data_ptr()), int(shard_axis), edge[lo_edge], edge[hi_edge],
                                       self._stream(padded)))

    def halo_bytes(self, axis):
        return int(self._L.ndnlm_halo_bytes(self._h, int(axis)))

    def halo_pack(self, padded, axis, side, msg):
        _lib.check(self._L.ndnlm_halo_pack(self._h, ctypes.c_void_p(padded.data_ptr()), int(axis), int(side),
                                           ctypes.c_void_p(msg.data_ptr()), self._stream(padded)))

    def halo_unpack(self, padded, axis, side, msg):
        _lib.check(self._L.ndnlm_halo_unpack(self._h, ctypes.c_void_p(padded.data_ptr()), int(axis), int(side),
                                             ctypes.c_void_p(msg.data_ptr()), self._stream(padded)))

    def run(self, padded, internal_out, err_flag, scratch=None):
        if scratch is None and self.scratch_bytes:
            # kept with the plan: stream-ordered reuse on the current stream (callers running one plan on several
            # streams at once pass their own scratch)
            if self._scratch is None or self._scratch.device != padded.device:
                self._scratch = torch.empty(self.scratch_bytes, dtype=torch.uint8, device=padded.device)
            scratch = self._scratch
        _lib.check(self._L.ndnlm_run_scratch(self._h, ctypes.c_void_p(padded.data_ptr()),
                                             ctypes.c_void_p(internal_out.data_ptr()), ctypes.c_void_p(err_flag.data_ptr()),
                                             ctypes.c_void_p(scratch.data_ptr() if scratch is not None else None),
                                             self._stream(padded)))

    def unstage(self, internal_out, output):
        self._check_arr(output)
        _lib.check(self._L.ndnlm_unstage(self._h, ctypes.c_void_p(internal_out.data_ptr()),
                                         ctypes.c_void_p(output.data_ptr()), _lib.i64(output.stride()), self._stream(internal_out)))

    def output_is_native(self, output):
        """True when `output` already has the layout (and alignment) of the internal output buffer, so that the
        kernels can write it directly (ndnlm_output_is_native) -- no second copy of the cube, no unstage pass."""
        return (tuple(output.shape) == self.shape and output.dtype == self.torch_dtype
                and output.data_ptr() % (4 * output.element_size()) == 0
                and bool(self._L.ndnlm_output_is_native(self._h, _lib.i64(output.stride()))))

    def run_into(self, padded, output, err_flag, internal_out=None, scratch=None):
        """run + unstage: straight into `output` when its layout allows it, through `internal_out` (allocated when
        None) otherwise.  Returns True when the direct route was taken."""
        self._check_arr(output)
        if self.output_is_native(output):
            self.run(padded, output, err_flag, scratch)
            return True
        if internal_out is None:
            internal_out = self.new_internal_out(padded.device)
        self.run(padded, internal_out, err_flag, scratch)
        self.unstage(internal_out, output)
        return False

    def apply(self, arr, output=None, workspace=None):
        """ndnlm_apply: stage + run + unstage on the current stream; raises ValueError('No solution')."""
        self._check_arr(arr)
        if output is None:
            output = torch.empty_like(arr)
        self._check_arr(output)
        if workspace is None:
            workspace = torch.empty(int(self._L.ndnlm_workspace_bytes(self._h)), dtype=torch.uint8, device=arr.device)
        _lib.check(self._L.ndnlm_apply(self._h, ctypes.c_void_p(arr.data_ptr()), _lib.i64(arr.stride()),
                                       ctypes.c_void_p(output.data_ptr()), _lib.i64(output.stride()),
                                       ctypes.c_void_p(workspace.data_ptr()), self._stream(arr)))
        return output


def synth_cube(ny_local, nx, nt, V=4, y_offset=0, seed=42, device="cuda"):
    """Synthetic SAR-like float32 cube (ny_local, nx, nt, V), keyed by GLOBAL index (ndnlm_synth_cube)."""
    out = torch.empty((ny_local, nx, nt, V), dtype=torch.float32, device=device)
    with torch.cuda.device(out.device):
        _lib.check(_lib.lib().ndnlm_synth_cube(ctypes.c_void_p(out.data_ptr()), ny_local, nx, nt, V, y_offset, seed,
                                               ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out


def synth_cube_into(out, y_offset=0, seed=42):
    """Fill the contiguous float32 device tensor `out` (rows, nx, nt, V) with rows [y_offset, y_offset + rows) of the
    synthetic cube (the same generator as `synth_cube`, keyed by GLOBAL index: slabs, shards and crops of the
    same cube agree wherever they overlap)."""
    if not (out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and out.dim() == 4):
        raise ValueError("expected a contiguous float32 CUDA tensor (rows, nx, nt, V)")
    ny, nx, nt, V = (int(x) for x in out.shape)
    with torch.cuda.device(out.device):
        _lib.check(_lib.lib().ndnlm_synth_cube(ctypes.c_void_p(out.data_ptr()), ny, nx, nt, V, int(y_offset), int(seed),
                                               ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out


def measure_fp32_peak(seconds=0.5):
    """Measured FP32 FMA peak of the current device in TFLOP/s (ndnlm_measure_fp32_peak)."""
    v = ctypes.c_double(0.0)
    _lib.check(_lib.lib().ndnlm_measure_fp32_peak(ctypes.byref(v), float(seconds), int(torch.cuda.current_device()),
                                                   ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return float(v.value)


def launch_count():
    return int(_lib.lib().ndnlm_launch_count())
