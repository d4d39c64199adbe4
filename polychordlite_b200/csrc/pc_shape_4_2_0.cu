// Instantiates the run kernels and the templated probes for G=4 lanes per point, DPL=2 dimensions per lane, likelihood kind 0.
#include "pc_run_kernel.cuh"
#include "pc_shapes.h"
namespace pc {
ShapeFns shape_fns_4_2_0() {
    return ShapeFns{(const void*)pc_run_kernel<4, 2, 0, 0>, (const void*)pc_slice_chains_kernel<4, 2, 0>,
                    (const void*)pc_calculate_points_kernel<4, 2, 0>,
                    (const void*)pc_run_kernel<4, 2, 0, 1>, (const void*)pc_slice_chains_dense_kernel<4, 2, 0>,
                    4, 2, 0};
}
}  // namespace pc
