// Pieces shared by the attention kernels: UMMA descriptors of the operand tile images and the kernel arguments.
#pragma once
#include <type_traits>

#include "common.cuh"
#include "reps.cuh"

namespace gta {

constexpr uint32_t kTmemCols = 512;

// Head layout [triv | se3 | so3 | so2] as a compile-time parameter (straight-line rep code for the shipped configs).
template <int TRIV, int SE3, int SO3, int SO2>
struct HeadLayout {
    static constexpr int kTriv = TRIV, kSe3 = SE3, kSo3 = SO3, kSo2 = SO2;
    static constexpr int D = TRIV + SE3 + SO3 + SO2;
    static constexpr int c1 = TRIV / 8, c2 = c1 + SE3 / 8, c3 = c2 + SO3 / 8, c4 = c3 + SO2 / 8;   // chunk boundaries
    static_assert(TRIV % 8 == 0 && SE3 % 8 == 0 && SO3 % 8 == 0 && SO2 % 8 == 0, "blocks are whole 16-byte chunks");
};

template <int I, int N, typename F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}



// K-step kk (16 elements of the head dim) of a K-major operand tile (Q' or K'), 64B swizzle:
// 8-row groups are 512 B apart (SBO); inside a 64-byte row the step advances the start address by 32 B.
__device__ __forceinline__ uint64_t desc_kmajor_sw64(uint32_t tile_addr, int kk) {
    return make_smem_desc(tile_addr + (kk >> 1) * 8192u + (kk & 1) * 32u, 16u, 512u, kLayoutSW64);
}
// K-step kk (16 keys) of the MN-major V' tile: N (head dim) is contiguous in 32-element atoms 8192 B apart
// (LBO); the 8-key groups along K are 512 B apart (SBO); 16 keys = 1024 B.
__device__ __forceinline__ uint64_t desc_mnmajor_sw64(uint32_t tile_addr, int kk) {
    return make_smem_desc(tile_addr + kk * 1024u, 8192u, 512u, kLayoutSW64);
}
// K-step kk (16 keys) of the K-major P tile (128B swizzle, 64-key column blocks 16 KB apart).
__device__ __forceinline__ uint64_t desc_p_sw128(uint32_t p_addr, int kk) {
    return make_smem_desc(p_addr + (kk >> 2) * 16384u + (kk & 3) * 32u, 16u, 1024u, kLayoutSW128);
}

// Low-word increments (16-byte units) of the K-step descriptors above, for the lean issue path.
__device__ __forceinline__ constexpr uint32_t kstep_kmajor_sw64(int kk) { return (kk >> 1) * 512u + (kk & 1) * 2u; }
__device__ __forceinline__ constexpr uint32_t kstep_mnmajor_sw64(int kk) { return kk * 64u; }
__device__ __forceinline__ constexpr uint32_t kstep_p_sw128(int kk) { return (kk >> 2) * 1024u + (kk & 3) * 2u; }

struct AttnArgs {
    const void* q;
    int64_t q_sb, q_sh, q_st;
    void* out;
    float* lse;
    const uint8_t* ws_k;
    const uint8_t* ws_v;
    int B, H, Tq, Tk, Nq, tpvq, ntiles_k, C;
    HeadDims hd;
    const float* se3_q;
    const float* so3_q;
    const float* so2_q;
    const float* tc_ptr;
    float scale, scale_log2;
    int v_transform;
    long long* dbg;
};


inline AttnArgs make_attn_args(const GtaAttnParams& p) {
    AttnArgs a;
    a.q = p.q; a.q_sb = p.q_stride_b; a.q_sh = p.q_stride_h; a.q_st = p.q_stride_t;
    a.out = p.out; a.lse = p.lse;
    a.ntiles_k = num_kv_tiles(p.Tk);
    const size_t half = static_cast<size_t>(p.B) * p.H * a.ntiles_k * kv_tile_bytes(p.D);
    a.ws_k = static_cast<const uint8_t*>(p.workspace);
    a.ws_v = a.ws_k + half;
    a.B = p.B; a.H = p.H; a.Tq = p.Tq; a.Tk = p.Tk; a.Nq = p.Nq; a.tpvq = p.Tq / p.Nq;
    a.C = p.so2 >> 1;
    a.hd = HeadDims{p.triv, p.se3, p.so3, p.so2};
    a.se3_q = p.reps.se3_q; a.so3_q = p.reps.so3_q; a.so2_q = p.reps.so2_q; a.tc_ptr = p.trans_coeff;
    a.scale = p.scale;
    a.scale_log2 = p.scale * 1.4426950408889634f;
    a.v_transform = p.v_transform;
    a.dbg = p.debug_clocks;
    return a;
}

}  // namespace gta
