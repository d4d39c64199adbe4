// pc_shapes.h -- the (G, DPL, KIND) instantiations of the templated kernels, one translation unit each.
// G = lanes per point group, DPL = dimensions per lane (a warp evaluates 32/G trial points at once),
// KIND = likelihood (0 Gaussian, 1 Rastrigin, 2 correlated Gaussian): compiled in, so the slice loop carries
// only its own likelihood and stays small in the instruction cache.
#pragma once
namespace pc {
struct ShapeFns {
    const void* run;     // pc_run_kernel<G, DPL, KIND, 0>
    const void* slice;   // pc_slice_chains_kernel<G, DPL, KIND>
    const void* calc;    // pc_calculate_points_kernel<G, DPL, KIND>
    const void* run_dense;    // pc_run_kernel<G, DPL, KIND, 1>: the dense chain phase (pc_dense.cuh); null for KIND 2
    const void* slice_dense;  // pc_slice_chains_dense_kernel<G, DPL, KIND>; null for KIND 2
    int G, DPL, KIND;
};
ShapeFns shape_fns_4_2_0();
ShapeFns shape_fns_4_2_1();
ShapeFns shape_fns_4_2_2();
ShapeFns shape_fns_4_4_0();
ShapeFns shape_fns_4_4_1();
ShapeFns shape_fns_4_4_2();
ShapeFns shape_fns_4_5_0();
ShapeFns shape_fns_4_5_1();
ShapeFns shape_fns_4_5_2();
ShapeFns shape_fns_4_8_0();
ShapeFns shape_fns_4_8_1();
ShapeFns shape_fns_4_8_2();
ShapeFns shape_fns_8_8_0();
ShapeFns shape_fns_8_8_1();
ShapeFns shape_fns_8_8_2();
ShapeFns shape_fns_16_8_0();
ShapeFns shape_fns_16_8_1();
ShapeFns shape_fns_16_8_2();
}  // namespace pc
