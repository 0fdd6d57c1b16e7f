// Single-launch forward (gta_attn_fwd4.cuh) with the head layout as a compile-time parameter: the staging warps run
// straight-line rep code with half a row in flight.  bf16 inputs; layouts of runs/clevrtr/GTA/gta (se3 32 | so2 32) and
// BASELINE config 1 (se3 16 | so2 16): 3.5 % faster than the run-time-layout code at the CLEVR decoder shape.  The d_h = 96
// layouts (runs/msn/GTA/gta_so3, gta) were measured too and are NOT instantiated: with the 34-entry Wigner table or six
// SE(3) chunks the straight-line code spills in the 88-register staging budget and is 4 % slower than the run-time code.
// Anything else uses the run-time-layout instantiations of gta_attn_fwd4.cu.
#include "gta_attn_fwd4.cuh"

namespace gta {

template <typename LY>
static int launch_ct(const GtaAttnParams& p, const AttnArgs& a, const Fused4Args& f, cudaStream_t st) {
    if (p.out_dtype == GTA_DTYPE_BF16) return launch4_one<__nv_bfloat16, __nv_bfloat16, LY::D, LY>(a, f, p, st);
    return launch4_one<__nv_bfloat16, float, LY::D, LY>(a, f, p, st);
}

int launch_attn_fwd_v3_ct(const GtaAttnParams& p, const AttnArgs& a, const Fused4Args& f, cudaStream_t st, bool* handled) {
    *handled = true;
    if (p.triv == 0 && p.se3 == 32 && p.so3 == 0 && p.so2 == 32) return launch_ct<HeadLayout<0, 32, 0, 32>>(p, a, f, st);
    if (p.triv == 0 && p.se3 == 16 && p.so3 == 0 && p.so2 == 16) return launch_ct<HeadLayout<0, 16, 0, 16>>(p, a, f, st);
    *handled = false;
    return GTA_OK;
}

}  // namespace gta
