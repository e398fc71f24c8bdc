// Explicit instantiations of the tiled kernel, group 6 (instances_g6.inc: double-duty halo warps); see nlm_tiled_launch.cuh.
#include "nlm_tiled_launch.cuh"

#define TILED_INST(NV4, FW, FX, FR, L, NW, CH, NEFF)                                                 \
    template cudaError_t launch_tiled<float, NV4, FW, FX, FR, L, NW, CH, NEFF>(                       \
        const CUtensorMap&, const ndnlm::DevParams&, const void*, void*, int*, int, size_t, cudaStream_t);
#define TILED_INSTH(NV4, FW, FX, FR, L, NW, CH, NEFF)                                                \
    template cudaError_t launch_tiled<float, NV4, FW, FX, FR, L, NW, CH, NEFF, true>(                 \
        const CUtensorMap&, const ndnlm::DevParams&, const void*, void*, int*, int, size_t, cudaStream_t);
#define TILED_INST64(NV4, FW, FX, FR, L, NW, CH, NEFF)                                               \
    template cudaError_t launch_tiled<double, NV4, FW, FX, FR, L, NW, CH, NEFF>(                      \
        const CUtensorMap&, const ndnlm::DevParams&, const void*, void*, int*, int, size_t, cudaStream_t);
#define TILED_INSTD(NV4, FW, FX, FR, L, NW, CH, NEFF)                                                \
    template cudaError_t launch_tiled<float, NV4, FW, FX, FR, L, NW, CH, NEFF, false, true>(          \
        const CUtensorMap&, const ndnlm::DevParams&, const void*, void*, int*, int, size_t, cudaStream_t);
#define TILED_INSTD64(NV4, FW, FX, FR, L, NW, CH, NEFF)                                              \
    template cudaError_t launch_tiled<double, NV4, FW, FX, FR, L, NW, CH, NEFF, false, true>(         \
        const CUtensorMap&, const ndnlm::DevParams&, const void*, void*, int*, int, size_t, cudaStream_t);
#include "instances_g6.inc"
