// kmc_stencil_nn8.cu -- lattice-stencil step kernels for 8 neighbour slots per site
// (BVO V:V electrons in any supercell >= 3x3xn, PyCD examples/BVO/InputFiles/sys_config.yml:20).
#include "kmc_stencil.cuh"
#include "kmc_stencil_launch.h"

namespace pycd {

bool stencil_launch_nn8(pycd_ctx *ctx, int nwc, int cpl, unsigned grid, size_t smem, const SysDev &S,
                         const StencilDev &T, const EnsDev &E, const AdvanceArgs &A)
{
    if (nwc == 1 && cpl == 1) launch_warp_step<1, 1, 8>(ctx, grid, smem, S, T, E, A);
    else if (nwc == 2 && cpl == 1) launch_warp_step<2, 1, 8>(ctx, grid, smem, S, T, E, A);
    else return false;
    return true;
}

}  // namespace pycd
