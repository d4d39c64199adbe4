// Instantiates the run kernels and the templated probes for G=4 lanes per point, DPL=5 dimensions per lane, likelihood kind 2.
#include "pc_run_kernel.cuh"
#include "pc_shapes.h"
namespace pc {
ShapeFns shape_fns_4_5_2() {
    return ShapeFns{(const void*)pc_run_kernel<4, 5, 2, 0>, (const void*)pc_slice_chains_kernel<4, 5, 2>,
                    (const void*)pc_calculate_points_kernel<4, 5, 2>,
                    nullptr, nullptr,
                    4, 5, 2};
}
}  // namespace pc
