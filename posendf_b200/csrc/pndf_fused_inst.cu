// pndf_fused_inst.cu -- one translation unit per (softplus DFNet, softplus encoder) combination of the fused kernel
// (compiled four times with -DPNDF_DSOFT=0|1 -DPNDF_ESOFT=0|1, in parallel, by __graft_entry__.build()): the kernel is
// ~40 k SASS instructions per instance, twelve instances in one unit take minutes to compile.
#include "pndf_kernel.cuh"

#ifndef PNDF_DSOFT
#error "compile with -DPNDF_DSOFT=0|1 -DPNDF_ESOFT=0|1"
#endif
#define PNDF_CAT3(a, b, c) a##b##c
#define PNDF_ENTRY(d, e) PNDF_CAT3(pndf_fused_entry_, d, e)

namespace pndf {
using FusedFn = void (*)(const KParams);
// kernel of MODE 0 (forward) / 1 (forward + reverse) / 2 (tangent) for this unit's activation combination
// small_tile: the 8-pose-tile variant (MODE 0 / 1 only; the training launches always use 32-pose tiles)
FusedFn PNDF_ENTRY(PNDF_DSOFT, PNDF_ESOFT)(int mode, int small_tile) {
    constexpr bool D = PNDF_DSOFT != 0, E = PNDF_ESOFT != 0;
    if (small_tile) return mode == 1 ? pndf_fused_kernel<1, D, E, true> : pndf_fused_kernel<0, D, E, true>;
    if (mode == 1) return pndf_fused_kernel<1, D, E>;
    if (mode == 2) return pndf_fused_kernel<2, D, E>;
    return pndf_fused_kernel<0, D, E>;
}
}  // namespace pndf
