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
    int so2_stage;      // attn_fwd3_kernel: row stride (floats) of the per-warp SO(2) staging rows in shared memory, 0 = off
};


// "Optimistic" softmax of one 128-column score tile: the exponentials are taken against the row's CURRENT reference maximum
// m_used (lazy rescaling lets a tile exceed it by up to 2^threshold), so they can start as soon as the first 32 score columns
// are in registers -- the tcgen05.ld of the next quarter is in flight while a quarter is exponentiated, and the tile maximum
// is computed alongside on the ALU pipe instead of in front of the first MUFU op.  Only after the last quarter is the maximum
// checked: if no row of the warp needs its reference advanced, P (bf16, packed in place over the registers the scores came in)
// is stored over S and the row sum is added to l_run -- bit-identical to the load-all / max / exp sequence of the caller.
// Otherwise nothing has been written (S is intact in tensor memory, sreg is clobbered, l_run untouched) and the caller runs
// that sequence for the tile.  Requires a finite m_used (not the first tile of an item) and a full tile (no masked columns).
// MEASURED SLOWER and therefore compiled out by default (GTA_OPTIMISTIC=1 enables it; profiles/r02_optimistic_softmax_ab.txt):
// attention kernel 0.430 vs 0.402 ms at MSN B=64, parity green.  ptxas already streams the four tcgen05.ld of the plain
// sequence through the scoreboard (tcgen05.wait::ld emits no instruction), and starting the exponentials earlier only makes
// the two warpgroups overlap their MUFU phases more: the loop is bound by the 16 MUFU lanes, not by what precedes them.
template <int kPolyNum, int kPolyDen>
__device__ __forceinline__ bool softmax_tile_optimistic(const uint32_t s_addr, uint32_t* sreg, const float m_used,
                                                        const float cs, const uint64_t cs2, const float threshold,
                                                        float& l_run) {
    float* s = reinterpret_cast<float*>(sreg);
    const float neg = -m_used * cs;
    const uint64_t neg2 = pack_f32x2(neg, neg);
    uint64_t lsum2 = pack_f32x2(0.f, 0.f);
    float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;     // maxima of x = (s - m_used) * cs
    tmem_ld32(s_addr, sreg);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        tmem_ld_wait32(sreg + q * 32);
        if (q < 3) tmem_ld32(s_addr + (q + 1) * 32, sreg + (q + 1) * 32);
        // same in-place packing as the caller's sequence: pair il of half h reads s[h*64 + 2il], s[h*64 + 2il + 1] and
        // lands in sreg[h*64 + il] (a register of this or the previous quarter, never one a load in flight is writing).
        // The maximum is taken over the scaled differences x (a score register dies with its FFMA2, as in the caller's
        // sequence: taking it over s kept both alive and spilled inside the loop at 184 registers).
        const int h = q >> 1;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int il = (q & 1) * 16 + k;
            const uint64_t x2 = ffma2(pack_f32x2(s[h * 64 + 2 * il], s[h * 64 + 2 * il + 1]), cs2, neg2);
            float x0, x1;
            unpack_f32x2(x2, x0, x1);
            if ((k & 3) == 0) mx0 = fmax3(mx0, x0, x1);
            else if ((k & 3) == 1) mx1 = fmax3(mx1, x0, x1);
            else if ((k & 3) == 2) mx2 = fmax3(mx2, x0, x1);
            else mx3 = fmax3(mx3, x0, x1);
            float p0, p1;
            if ((il % kPolyDen) < kPolyNum) {
                poly_exp2x2(x2, p0, p1);
            } else {
                p0 = fast_exp2(x0); p1 = fast_exp2(x1);
            }
            lsum2 = fadd2(lsum2, pack_f32x2(p0, p1));
            sreg[h * 64 + il] = pack_bf16x2(p0, p1);
        }
    }
    // (s - m_used) * cs > threshold for some column <=> the caller's test on the tile maximum (up to the rounding of the FMA)
    const bool grow = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) > threshold;
    if (__any_sync(0xffffffffu, grow)) return false;
    tmem_st32(s_addr, sreg);
    tmem_st32(s_addr + 32, sreg + 64);
    float ls0, ls1;
    unpack_f32x2(lsum2, ls0, ls1);
    l_run += ls0 + ls1;
    return true;
}

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
    a.so2_stage = 0;
    return a;
}

}  // namespace gta
