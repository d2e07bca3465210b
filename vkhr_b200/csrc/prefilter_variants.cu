// prefilter_variants.cu -- instantiations of the tiled prefilter kernel (prefilter.cuh), one tap-offset pair
// (VKHR_PF_NO0, VKHR_PF_PO0) per object file, both tile depths; group 0 also carries the row-wise instantiation.
// Compiled once per group by vkhr_b200/build.py (-DVKHR_PF_G=g ...): the 19 instantiations used to take 1 m 45 s
// in one translation unit.  The host's dispatch table (vkhr_b200.cu) is assembled from pf_variants_<g>().
#define VKHR_PF_TILED_ONLY
#include "prefilter.cuh"

#if !defined(VKHR_PF_G) || !defined(VKHR_PF_NO0) || !defined(VKHR_PF_PO0)
#error "build with -DVKHR_PF_G=<group> -DVKHR_PF_NO0=<n> -DVKHR_PF_PO0=<p> (vkhr_b200/build.py)"
#endif

#define VKHR_PF_CAT_(a, b) a##b
#define VKHR_PF_CAT(a, b) VKHR_PF_CAT_(a, b)

namespace vkhr_b200 {

// Appends this group's variants to `out`; returns how many.
int VKHR_PF_CAT(pf_variants_, VKHR_PF_G)(PfVariant* out) {
    int n = 0;
#if VKHR_PF_G == 0
    out[n++] = {kPfRowWise, kPfRowWise, kPfTZ, k_prefilter_tiled<kPfRowWise, kPfRowWise, kPfTZ>};
#endif
    out[n++] = {VKHR_PF_NO0, VKHR_PF_PO0, kPfTZ, k_prefilter_tiled<VKHR_PF_NO0, VKHR_PF_PO0, kPfTZ>};
    out[n++] = {VKHR_PF_NO0, VKHR_PF_PO0, kPfTZDeep, k_prefilter_tiled<VKHR_PF_NO0, VKHR_PF_PO0, kPfTZDeep>};
    return n;
}

}  // namespace vkhr_b200
