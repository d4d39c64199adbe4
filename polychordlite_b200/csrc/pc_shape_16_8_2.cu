// Instantiates the run kernels and the templated probes for G=16 lanes per point, DPL=8 dimensions per lane, likelihood kind 2.
#include "pc_run_kernel.cuh"
#include "pc_shapes.h"
namespace pc {
ShapeFns shape_fns_16_8_2() {
    return ShapeFns{(const void*)pc_run_kernel<16, 8, 2, 0>, (const void*)pc_slice_chains_kernel<16, 8, 2>,
                    (const void*)pc_calculate_points_kernel<16, 8, 2>,
                    nullptr, nullptr,
                    16, 8, 2};
}
}  // namespace pc
