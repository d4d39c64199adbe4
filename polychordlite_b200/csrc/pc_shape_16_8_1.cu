// Instantiates the run kernels and the templated probes for G=16 lanes per point, DPL=8 dimensions per lane, likelihood kind 1.
#include "pc_run_kernel.cuh"
#include "pc_shapes.h"
namespace pc {
ShapeFns shape_fns_16_8_1() {
    return ShapeFns{(const void*)pc_run_kernel<16, 8, 1, 0>, (const void*)pc_slice_chains_kernel<16, 8, 1>,
                    (const void*)pc_calculate_points_kernel<16, 8, 1>,
                    (const void*)pc_run_kernel<16, 8, 1, 1>, (const void*)pc_slice_chains_dense_kernel<16, 8, 1>,
                    16, 8, 1};
}
}  // namespace pc
