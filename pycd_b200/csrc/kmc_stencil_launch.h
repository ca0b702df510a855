// kmc_stencil_launch.h -- host entry points of the lattice-stencil step kernels.  The kernels are
// instantiated per neighbour-slot count in their own translation units (kmc_stencil_nn4.cu, _nn8.cu,
// _nn12.cu) so that they compile in parallel.
#pragma once

#include "kmc_types.cuh"

namespace pycd {

// Launches kmc_step_warp_kernel<nwc, cpl, NN> (one CTA of nwc warps per trajectory, cpl carriers per
// lane); returns false when the shape is not instantiated.
bool stencil_launch_nn4(pycd_ctx *ctx, int nwc, int cpl, unsigned grid, size_t smem, const SysDev &S,
                        const StencilDev &T, const EnsDev &E, const AdvanceArgs &A);
bool stencil_launch_nn8(pycd_ctx *ctx, int nwc, int cpl, unsigned grid, size_t smem, const SysDev &S,
                        const StencilDev &T, const EnsDev &E, const AdvanceArgs &A);
bool stencil_launch_nn12(pycd_ctx *ctx, int nwc, int cpl, unsigned grid, size_t smem, const SysDev &S,
                         const StencilDev &T, const EnsDev &E, const AdvanceArgs &A);

}  // namespace pycd
