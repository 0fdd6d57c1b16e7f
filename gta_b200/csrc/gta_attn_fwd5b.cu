// Instantiations of the v4 pipeline (gta_attn_fwd5.cuh), head dims 64 and 32: runs/clevrtr/GTA/gta (se3 32 | so2 32),
// BASELINE config 1 (se3 16 | so2 16 and se3 16 | so3 8 | so2 8) and the all-trivial layouts of the generic path.
#include "gta_attn_fwd5.cuh"

namespace gta {

int launch_attn_fwd_v4_d96(const GtaAttnParams& p, cudaStream_t st, bool* handled);

int launch_attn_fwd_v4(const GtaAttnParams& p, cudaStream_t st, bool* handled) {
    *handled = true;
    if (p.D == 96) return launch_attn_fwd_v4_d96(p, st, handled);
    if (p.D == 64) {
        if (p.triv == 0 && p.se3 == 32 && p.so3 == 0 && p.so2 == 32) return launch5_layout<HeadLayout<0, 32, 0, 32>>(p, st);
        if (p.triv == 64 && p.se3 == 0 && p.so3 == 0 && p.so2 == 0) return launch5_layout<HeadLayout<64, 0, 0, 0>>(p, st);
    }
    if (p.D == 32) {
        if (p.triv == 0 && p.se3 == 16 && p.so3 == 0 && p.so2 == 16) return launch5_layout<HeadLayout<0, 16, 0, 16>>(p, st);
        if (p.triv == 0 && p.se3 == 16 && p.so3 == 8 && p.so2 == 8) return launch5_layout<HeadLayout<0, 16, 8, 8>>(p, st);
    }
    *handled = false;
    return GTA_OK;
}

}  // namespace gta
