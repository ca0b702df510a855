// kmc_stencil_nn4.cu -- lattice-stencil step kernels for 4 neighbour slots per site (Hematite Fe:Fe,
// PyCD examples/Hematite/InputFiles/sys_config.yml): the benchmark shape.
#include "kmc_stencil.cuh"
#include "kmc_stencil_launch.h"

namespace pycd {

bool stencil_launch_nn4(pycd_ctx *ctx, int nwc, int cpl, unsigned grid, size_t smem, const SysDev &S,
                        const StencilDev &T, const EnsDev &E, const AdvanceArgs &A)
{
    if (nwc == 1 && cpl == 1) launch_warp_step<1, 1, 4>(ctx, grid, smem, S, T, E, A);
    else if (nwc == 2 && cpl == 1) launch_warp_step<2, 1, 4>(ctx, grid, smem, S, T, E, A);
    else if (nwc == 1 && cpl == 2) launch_warp_step<1, 2, 4>(ctx, grid, smem, S, T, E, A);
    else if (nwc == 2 && cpl == 2) launch_warp_step<2, 2, 4>(ctx, grid, smem, S, T, E, A);
    else return false;
    return true;
}

}  // namespace pycd

#ifdef PYCD_TRACE
extern "C" int pycd_debug_trace(long long *out) {
    return pycd::guarded([&] {
        PYCD_CUDA(cudaMemcpyFromSymbol(out, pycd::g_st_trace, sizeof(long long) * 256 * 16 * 16));
    });
}
#endif
