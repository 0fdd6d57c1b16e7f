// Instantiations of the v4 pipeline (gta_attn_fwd5.cuh), head dim 96: runs/msn/GTA/gta_so3 (se3 48 | so3 24 | so2 24),
// runs/msn/GTA/gta (se3 48 | so2 48) and the all-trivial layout the generic path hands over.
#include "gta_attn_fwd5.cuh"

namespace gta {

int launch_attn_fwd_v4_d96(const GtaAttnParams& p, cudaStream_t st, bool* handled) {
    *handled = true;
    if (p.triv == 0 && p.se3 == 48 && p.so3 == 24 && p.so2 == 24) return launch5_layout<HeadLayout<0, 48, 24, 24>>(p, st);
    if (p.triv == 0 && p.se3 == 48 && p.so3 == 0 && p.so2 == 48) return launch5_layout<HeadLayout<0, 48, 0, 48>>(p, st);
    if (p.triv == 96 && p.se3 == 0 && p.so3 == 0 && p.so2 == 0) return launch5_layout<HeadLayout<96, 0, 0, 0>>(p, st);
    *handled = false;
    return GTA_OK;
}

}  // namespace gta
