/*
 * ndnlm.h -- C ABI of libndnlm.so: the B200 (sm_100a) non-local-means hot path.
 *
 * Drop-in boundary.  This library replaces the native kernel of jnhansen/nd,
 *
 *     cpdef void _pixelwise_nlmeans_3d(floating[:,:,:,:] arr, floating[:,:,:,:] output,
 *                                      unsigned int[:] r, unsigned int[:] f,
 *                                      double sigma, double h, double n_eff=-1)
 *                                                         (reference nd/_filters.pyx:320-325)
 *
 * which `NLMeansFilter._filter` calls as
 *     _pixelwise_nlmeans_3d(values, _out, r, f, self.sigma, self.h, self.n_eff)
 *                                                         (reference nd/filters.py:462-463).
 *
 * Everything here is plain C: pointers, sizes, no torch / C++ types.  All array pointers
 * are DEVICE pointers unless a parameter says "host"; the library binds the device that owns the
 * buffers itself (it carries its own CUDA runtime, independent of the caller's "current device"),
 * and `stream` must be a stream of that device.
 * The caller owns every buffer; the library allocates nothing that outlives a call except
 * the plan object.  Functions return 0 on success or a negative NDNLM_E* code and never
 * throw; `ndnlm_last_error()` gives the message for the calling thread.
 *
 * Array convention (the reference's): `arr`/`output` are 4-D (N0, N1, N2, V) with arbitrary
 * ELEMENT strides; `r`, `f` are uint32[3] search / patch radii per axis.
 */
#ifndef NDNLM_H
#define NDNLM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error codes (mapped by the Python shim onto the reference's exception types) ---- */
#define NDNLM_OK              0
#define NDNLM_EINVAL         -1   /* bad argument -> ValueError                                    */
#define NDNLM_EDTYPE         -2   /* non float32/float64 data -> TypeError (ref: "No matching signature") */
#define NDNLM_ECUDA          -3   /* CUDA runtime / driver failure -> RuntimeError                 */
#define NDNLM_ENOSOLUTION    -4   /* find_weight: ValueError('No solution') (ref nd/_filters.pyx:310-311) */
#define NDNLM_WUNDERFLOW      1   /* WARNING (result valid): some voxels saw only weights below the fp32 range and were
                                     left unfiltered -> RuntimeWarning (the float64 reference keeps such weights)        */
#define NDNLM_ERADIUS        -5   /* r_i + f_i > N_i - 1: single reflection undefined (ref `_idx` :34-40 is UB) */

/* ---- dtypes of the caller's arrays ---- */
#define NDNLM_F32 0
#define NDNLM_F64 1

/* ---- semantics ---- */
#define NDNLM_AS_WRITTEN          0  /* patch distances as the .pyx text / docs describe             */
#define NDNLM_REFERENCE_COMPILED  1  /* bug-for-bug the LP64 binary: any f_i>0 => d^2 == 0 (nd/_filters.pyx:323,373-375) */

/* ---- kernel selection ---- */
#define NDNLM_KERNEL_AUTO     0  /* tiled kernel in the data type when the configuration has an instantiation, else generic */
#define NDNLM_KERNEL_GENERIC  1  /* reference-faithful one-thread-per-voxel kernel (fp64 weights), any r/f/V      */
#define NDNLM_KERNEL_TILED    2  /* TMA-tiled fp32 kernel; error if not instantiated for this configuration.      */
                                 /* With float64 data: staged as float32, result widened back (opt-in fp32 compute) */
#define NDNLM_KERNEL_TILED_F64 3 /* float64 data on the float64 instantiations of the tiled kernel (what AUTO picks */
                                 /* for float64 data when one fits); error if there is none                         */

/* ---- how the two ends of user axis `shard_axis` are padded by ndnlm_stage ---- */
#define NDNLM_EDGE_REFLECT 0     /* global edge: reflect locally (reference `_idx`, nd/_filters.pyx:34-40) */
#define NDNLM_EDGE_HALO    1     /* interior shard edge: halo rows are filled by the caller (neighbour exchange) */
#define NDNLM_EDGE_SOURCE  2     /* slab of a larger array: `arr` has r+f real rows beyond this edge (the `buffer`
                                    rows of the reference's xr_split, nd/utils.py:288-312); they are read in place */

typedef struct ndnlm_plan ndnlm_plan_t;

typedef struct ndnlm_info {
    int32_t  kernel;            /* NDNLM_KERNEL_GENERIC or NDNLM_KERNEL_TILED (what will run)          */
    int32_t  role_axis[3];      /* user axis (0..2) playing role W (slowest), R (register column), X (lanes) */
    int32_t  n[3];              /* interior extent per role                                              */
    int32_t  pad[3];            /* r+f per role                                                          */
    int32_t  padded[3];         /* n + 2*pad per role                                                    */
    int32_t  vp;                /* variables after padding to a multiple of 4 (tiled) or V (generic)     */
    int32_t  tile[3];           /* valid voxels per CTA tile per role (tiled kernel)                     */
    int32_t  box[3];            /* shared-memory box per role (tiled kernel)                             */
    int32_t  warps[3];          /* warp grid per role (tiled kernel)                                     */
    int32_t  threads;           /* threads per CTA                                                       */
    int32_t  grid;              /* CTAs per launch                                                       */
    int32_t  smem_bytes;        /* dynamic shared memory per CTA                                         */
    int32_t  elem_bytes;        /* bytes per element of the staged (internal) buffers: 4 or 8            */
    int64_t  n_offsets;         /* K = prod(2 r_i + 1) - 1                                               */
    int64_t  voxels;            /* N0*N1*N2                                                              */
    double   flops_per_voxel;   /* algorithmic F of SURVEY.md 8(d): K*(5V + 2 n_a + 6 [+2]) + 3V + 3     */
    size_t   padded_bytes;      /* size of the staged, reflect-padded input cube                         */
    size_t   out_bytes;         /* size of the internal output buffer                                    */
    char     kernel_name[96];   /* human-readable instantiation name                                     */
} ndnlm_info_t;

/*
 * Create a plan for one (shape, r, f, sigma, h, n_eff, semantics, dtype) configuration.
 * Replaces the argument handling at the top of _pixelwise_nlmeans_3d (nd/_filters.pyx:326-341).
 *   shape     (N0, N1, N2, V)
 *   dtype     NDNLM_F32 | NDNLM_F64 -- dtype of the caller's arrays.  float64 data is computed by the
 *             generic kernel in float64 (the reference's fused `floating`, nd/_filters.pyx:320-321).
 *   kernel    NDNLM_KERNEL_*
 * Errors: NDNLM_EINVAL, NDNLM_EDTYPE, NDNLM_ERADIUS.
 */
int ndnlm_plan_create(ndnlm_plan_t** plan, const int64_t shape[4],
                      const uint32_t r[3], const uint32_t f[3],
                      double sigma, double h, double n_eff,
                      int semantics, int dtype, int kernel);
/*
 * The same with the role assignment given by the caller: role_axis[k] = user axis (0..2) that plays role
 * W, R, X (k = 0, 1, 2), as reported by ndnlm_plan_info of another plan; NULL = choose from the shape.
 * ndnlm_plan_create chooses the X / R roles from the extents of axes 1 and 2, so the shards of an array cut along
 * axis 1 or 2 could otherwise end up with different staged layouts than their neighbours (their halo messages
 * would be mis-read); a sharding driver creates the plan of the WHOLE array first and hands its roles to every shard.
 */
int ndnlm_plan_create_roles(ndnlm_plan_t** plan, const int64_t shape[4],
                            const uint32_t r[3], const uint32_t f[3],
                            double sigma, double h, double n_eff,
                            int semantics, int dtype, int kernel, const int32_t* role_axis);
void ndnlm_plan_destroy(ndnlm_plan_t* plan);
int  ndnlm_plan_info(const ndnlm_plan_t* plan, ndnlm_info_t* info);

/*
 * Stage: strided caller array -> internal reflect-padded cube (`padded`, info.padded_bytes).
 * Replaces the per-access `_idx(p+d, N, m)` reflection (nd/_filters.pyx:378-384) by materialising
 * np.pad(mode='reflect') by r+f once.  `shard_axis` (user axis 0..2, or -1) with edge modes lets a
 * y-shard skip reflection on interior edges: with NDNLM_EDGE_HALO those pad rows are left
 * untouched and must be written by the caller via ndnlm_halo_* before ndnlm_run; with
 * NDNLM_EDGE_SOURCE the r+f rows beyond that edge are read from `arr` itself (indices -(r+f)..-1
 * resp. N..N+r+f-1 along `shard_axis` must be valid memory: `arr` points into a larger array).
 */
int ndnlm_stage(const ndnlm_plan_t* plan, const void* arr, const int64_t arr_strides[4],
                void* padded, int shard_axis, int lo_edge, int hi_edge, void* stream);

/*
 * Halo rows of the padded cube along user axis `axis`:
 *   ndnlm_halo_bytes           size of one packed halo message (pad rows x all other padded extents)
 *   ndnlm_halo_pack(side)      copy MY first (side=0) / last (side=1) `pad` INTERIOR rows into `msg`
 *                              (what my lower / upper neighbour needs)
 *   ndnlm_halo_unpack(side)    write a received `msg` into MY lower (side=0) / upper (side=1) pad rows
 * The exchange itself (NCCL send/recv or peer copies over NVLink) is the caller's.
 */
size_t ndnlm_halo_bytes(const ndnlm_plan_t* plan, int axis);
int ndnlm_halo_pack(const ndnlm_plan_t* plan, const void* padded, int axis, int side, void* msg, void* stream);
int ndnlm_halo_unpack(const ndnlm_plan_t* plan, void* padded, int axis, int side, const void* msg, void* stream);

/*
 * Run the filter: padded cube -> internal output buffer (`out_internal`, info.out_bytes).
 * Replaces the voxel / search-window / patch loops of nd/_filters.pyx:351-420.
 * `err_flag` is a device int32 (zeroed by the caller), a bit mask the kernels OR into:
 *   NDNLM_FLAG_NOSOLUTION  find_weight has no solution at some voxel (nd/_filters.pyx:310-311) -> ValueError('No solution')
 *   NDNLM_FLAG_UNDERFLOW   fp32 tiled kernel only: at some voxel every neighbour weight flushed to zero (< 2^-126);
 *                          that voxel is returned unfiltered -> a warning, never a silent NaN
 */
#define NDNLM_FLAG_NOSOLUTION 1
#define NDNLM_FLAG_UNDERFLOW  2
int ndnlm_run(const ndnlm_plan_t* plan, const void* padded, void* out_internal,
              int32_t* err_flag, void* stream);

/* Scratch memory some kernels need between their passes (today: the intermediate of the separable box-mean fast
 * path of NDNLM_REFERENCE_COMPILED; 0 for everything else).  ndnlm_run allocates it stream-ordered on every call
 * (cudaMallocAsync, slow for multi-GB cubes); ndnlm_run_scratch takes a caller-owned buffer of
 * ndnlm_scratch_bytes(plan) bytes instead (may be NULL when that is 0). */
size_t ndnlm_scratch_bytes(const ndnlm_plan_t* plan);
int ndnlm_run_scratch(const ndnlm_plan_t* plan, const void* padded, void* out_internal,
                      int32_t* err_flag, void* scratch, void* stream);

/* Unstage: internal output buffer -> strided caller array (the in-place write into `output`). */
int ndnlm_unstage(const ndnlm_plan_t* plan, const void* out_internal,
                  void* output, const int64_t out_strides[4], void* stream);

/* 1 when a caller array with these element strides already HAS the layout of the internal output buffer (four
 * variables of the compute type, contiguous behind the axes in the plan's staged order -- e.g. a C-ordered
 * (y, x, time, 4) float32 cube): ndnlm_run may then be given the caller's `output` (16- / 32-byte aligned) as
 * `out_internal` and ndnlm_unstage skipped -- the in-place write of nd/_filters.pyx:420 without a second copy of the
 * cube.  ndnlm_apply does this itself.  0 otherwise. */
int ndnlm_output_is_native(const ndnlm_plan_t* plan, const int64_t out_strides[4]);

/*
 * One-call form of the reference entry point on device arrays: stage + run + unstage on `stream`,
 * then synchronises the stream and returns NDNLM_ENOSOLUTION / NDNLM_WUNDERFLOW according to the flag.
 * `workspace` must hold ndnlm_workspace_bytes(plan) bytes.
 */
size_t ndnlm_workspace_bytes(const ndnlm_plan_t* plan);
int ndnlm_apply(const ndnlm_plan_t* plan, const void* arr, const int64_t arr_strides[4],
                void* output, const int64_t out_strides[4], void* workspace, void* stream);

/*
 * Synthetic "complex-SAR-like" cube: float32 (ny_local, nx, nt, V) C-contiguous rows
 * [y_offset, y_offset+ny_local) of a global cube, every value a pure function of
 * (seed, global y, x, t, v) so any sharding sees the same cube.  V = 4: C11, C12__re, C12__im, C22;
 * V = 6 adds C33 and a second cross term.
 */
int ndnlm_synth_cube(float* out, int64_t ny_local, int64_t nx, int64_t nt, int32_t V,
                     int64_t y_offset, uint64_t seed, void* stream);

/*
 * FP32 FMA-chain microbenchmark on the current device (runs for about `seconds`): the measured
 * CUDA-core peak in TFLOP/s, reported by bench.py beside the nominal SMs*128*2*clock roofline.
 */
int ndnlm_measure_fp32_peak(double* tflops, double seconds, int device, void* stream);

/* Number of kernel launches issued by this library since load (bench.py's gpu_launches). */
int64_t ndnlm_launch_count(void);

const char* ndnlm_last_error(void);
const char* ndnlm_version(void);

#ifdef __cplusplus
}
#endif
#endif /* NDNLM_H */
