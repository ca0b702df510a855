// kmc_stencil_nn12.cu -- lattice-stencil step kernels for 12 neighbour slots per site
// (BVO O:O holes, two site classes, PyCD/core.py:1961-1970).
#include "kmc_stencil.cuh"
#include "kmc_stencil_launch.h"

namespace pycd {

bool stencil_launch_nn12(pycd_ctx *ctx, int nwc, int cpl, unsigned grid, size_t smem, const SysDev &S,
                         const StencilDev &T, const EnsDev &E, const AdvanceArgs &A)
{
    if (nwc == 1 && cpl == 1) launch_warp_step<1, 1, 12>(ctx, grid, smem, S, T, E, A);
    else if (nwc == 2 && cpl == 1) launch_warp_step<2, 1, 12>(ctx, grid, smem, S, T, E, A);
    else return false;
    return true;
}

}  // namespace pycd
