// pc_shapes.h -- the (G, DPL) instantiations of the templated kernels, one translation unit each.
// G = lanes per point group, DPL = dimensions per lane; a warp evaluates 32/G trial points at once.
#pragma once
namespace pc {
struct ShapeFns {
    const void* run;     // pc_run_kernel<G, DPL>
    const void* slice;   // pc_slice_chains_kernel<G, DPL>
    const void* calc;    // pc_calculate_points_kernel<G, DPL>
    int G, DPL;
};
ShapeFns shape_fns_4_2();
ShapeFns shape_fns_4_4();
ShapeFns shape_fns_4_5();
ShapeFns shape_fns_4_8();
ShapeFns shape_fns_8_8();
ShapeFns shape_fns_16_8();
}  // namespace pc
