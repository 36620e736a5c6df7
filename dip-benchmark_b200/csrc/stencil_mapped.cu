// The tiled stencil kernels with a per-byte map fused in front of and / or behind the stencil
// (dip_apply_chain): `pre` is applied to the source words as a thread unpacks them from the shared-memory tile, `post`
// to the 16 result bytes in registers right before the store, so a chain such as
// threshold -> erode -> invert is one pass over HBM.  Same kernels as stencil.cu, instantiated with MAPPED = true
// in a translation unit of their own (they compile in parallel with the plain ones).
#include "stencil_tiled.cuh"

namespace dip {

namespace {

template <int OP>
int launch_mapped_c(dip_ctx* ctx, const dip_image* src, dip_image* dst, const RowGeom& g, const ChainMaps& maps) {
    switch (src->v.channels) {
        case 1: return launch_tiled<OP, 1, 15, 4, true>(ctx, src, dst, g, &maps);
        case 3: return launch_tiled<OP, 3, 15, 4, true>(ctx, src, dst, g, &maps);
        case 4: return launch_tiled<OP, 4, 15, 4, true>(ctx, src, dst, g, &maps);
    }
    return fail(ctx, DIP_ERR_ARG, "bad channel count");
}

}  // namespace

int launch_tiled_mapped(dip_ctx* ctx, int op, const dip_image* src, dip_image* dst, const RowGeom& g, const ChainMaps& maps) {
    switch (op) {
#define DIP_T(OP) case OP: return launch_mapped_c<OP>(ctx, src, dst, g, maps);
        // the two-dimensional kernels only: the one-dimensional passes are not instantiated with maps
        // (stencil_takes_maps; the planner gives their neighbours' maps a point pass of their own)
        DIP_T(kErodeSq3) DIP_T(kErodeCross3) DIP_T(kDilateSq3) DIP_T(kDilateCross3) DIP_T(kConv3) DIP_T(kConv3Sep) DIP_T(kGauss3)
        DIP_T(kConv5) DIP_T(kConv5Sep)
#undef DIP_T
    }
    return fail(ctx, DIP_ERR_UNSUPPORTED, "tiled op %d takes no per-byte maps", op);
}

}  // namespace dip
