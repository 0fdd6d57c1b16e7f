// K'/V' staging, device side: one 128-key tile of one (batch, head) rotated by 128 threads and written as bf16 UMMA operand
// tile images (see gta_rotate_kv.cu for the layout and the reference semantics).  Shared by the stand-alone staging kernel
// and by the staging warps of the fused single-launch attention kernel.
#pragma once
#include "common.cuh"
#include "reps.cuh"

namespace gta {

struct RotArgs {
    const void* k; const void* v;
    int64_t k_sb, k_sh, k_st, v_sb, v_sh, v_st;
    uint8_t* ws_k; uint8_t* ws_v;
    int H, Tk, D, Nk, ntiles, tpv;
    HeadDims hd;
    const float* se3_k; const float* so3_k; const float* so2_k;
    const float* tc_ptr;
    int C;            // so2 pairs per token
    int v_transform;
    size_t lo_offset; // split-precision mode: byte offset from a hi tile image to its lo (residual) image, else 0
};

// x -> bf16 hi part and bf16 residual (x - hi): hi + lo carries ~16 mantissa bits (fp32-accurate mode)
__device__ __forceinline__ void split_chunk_bf16(const float* x, uint4& hi, uint4& lo) {
    hi = pack_chunk_bf16(x);
    float r[8];
    r[0] = x[0] - bf16_lo(hi.x); r[1] = x[1] - bf16_hi(hi.x); r[2] = x[2] - bf16_lo(hi.y); r[3] = x[3] - bf16_hi(hi.y);
    r[4] = x[4] - bf16_lo(hi.z); r[5] = x[5] - bf16_hi(hi.z); r[6] = x[6] - bf16_lo(hi.w); r[7] = x[7] - bf16_hi(hi.w);
    lo = pack_chunk_bf16(r);
}

// One warp owns 32 consecutive keys of the tile and walks the head row block type by block type (se3 chunks, then
// so3, then so2) so that every instruction is type-uniform while lanes still cover contiguous 16-byte chunks.
// K and V of the same (key, chunk) share the rep data, and kUnroll (key, chunk) items are loaded before the first
// one is consumed: ~2*kUnroll 16-byte loads in flight per thread keep HBM busy at low occupancy cost.
#ifndef GTA_ROT_UNROLL
#define GTA_ROT_UNROLL 3
#endif
constexpr int kRotUnroll = GTA_ROT_UNROLL;

// kQSide = false: K' = rho_k K, V' = rho_k V (forward and backward staging).
// kQSide = true : the same walk applied to the QUERY side of the backward: Q' = rho_q^{-T} Q and dO' = rho_q^{-T} dO share
// their rep data exactly as K and V do (the "k" slot carries q, the "v" slot dout, v_transform gates the dO' rotation);
// only the SE(3) block differs (transposed matrix).
// `warp` in [0,4), `lane`: the 128 cooperating threads (a whole CTA of rotate_kv_kernel, or the four staging warps of the
// fused attention kernel, gta_attn_fwd4.cu).
template <typename T, bool kQSide>
__device__ __forceinline__ void rotate_tile(const RotArgs& a, const int tile, const int h, const int b, const int warp,
                                            const int lane, const float tc) {
    const size_t tile_bytes = static_cast<size_t>(128) * a.D * 2;
    const size_t blob = (static_cast<size_t>(b) * a.H + h) * a.ntiles + tile;
    const int seg_n[4] = {a.hd.triv >> 3, a.hd.se3 >> 3, a.hd.so3 >> 3, a.hd.so2 >> 3};
    const T* ksrc = reinterpret_cast<const T*>(a.k) + static_cast<int64_t>(b) * a.k_sb + static_cast<int64_t>(h) * a.k_sh;
    const T* vsrc = reinterpret_cast<const T*>(a.v) + static_cast<int64_t>(b) * a.v_sb + static_cast<int64_t>(h) * a.v_sh;
    uint8_t* kdst = a.ws_k + blob * tile_bytes;
    uint8_t* vdst = a.ws_v + blob * tile_bytes;

    // view of a row without a division per item: tile rows are consecutive tokens, and with at least 128 tokens per view
    // (every model shape) a tile crosses at most one view boundary
    const int t0 = tile * 128;
    const int view0 = t0 / a.tpv, rem0 = t0 - view0 * a.tpv;
    const bool one_boundary = a.tpv >= 128;

    int cbase = 0;
#pragma unroll 1
    for (int seg = 0; seg < 4; ++seg) {
        const int n_t = seg_n[seg];
        if (n_t == 0) continue;
        // this warp's items of the segment: (row r, chunk cbase + c), item index it = r * n_t + c, it in [0, 32 n_t); lane l
        // holds items l, l + 32, ...: (r, c) advance by (32 / n_t, 32 % n_t) with a carry instead of a division per item
        const int q32 = 32 / n_t, m32 = 32 - q32 * n_t;
        int r_it = lane / n_t, c_it = lane - r_it * n_t;
#pragma unroll 1
        for (int base = 0; base < n_t; base += kRotUnroll) {
            RawChunk<T> rk[kRotUnroll], rv[kRotUnroll];
            int row[kRotUnroll], ch[kRotUnroll];
#pragma unroll
            for (int u = 0; u < kRotUnroll; ++u) {
                ch[u] = cbase + c_it;
                row[u] = (base + u < n_t) ? warp * 32 + r_it : -1;
                c_it += m32; r_it += q32;
                if (c_it >= n_t) { c_it -= n_t; ++r_it; }
                const int t = t0 + row[u];
                zero_raw(rk[u]); zero_raw(rv[u]);
                if (row[u] >= 0 && t < a.Tk) {
                    load_raw(ksrc + t * a.k_st + ch[u] * 8, rk[u]);
                    load_raw(vsrc + t * a.v_st + ch[u] * 8, rv[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < kRotUnroll; ++u) {
                if (row[u] < 0) continue;
                const int t = t0 + row[u];
                float xk[8], xv[8];
                raw_to_f32(rk[u], xk);
                raw_to_f32(rv[u], xv);
                if (seg > 0 && t < a.Tk) {
                    const int vrow = one_boundary ? view0 + (rem0 + row[u] >= a.tpv ? 1 : 0) : t / a.tpv;
                    const size_t view = static_cast<size_t>(b) * a.Nk + vrow;
                    const float* se3 = a.se3_k + view * 16;
                    const float* so3 = a.so3_k + view * 34;
                    const float* so2 = a.so2_k + (static_cast<size_t>(b) * a.Tk + t) * a.C * 2;
                    if (seg == 1) {
                        float M[16];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            float4 q4 = __ldg(reinterpret_cast<const float4*>(se3) + i);
                            M[4 * i] = q4.x; M[4 * i + 1] = q4.y; M[4 * i + 2] = q4.z; M[4 * i + 3] = q4.w;
                        }
                        if (kQSide) {
                            se3_apply_T(xk, M, tc);
                            if (a.v_transform) se3_apply_T(xv, M, tc);
                        } else {
                            se3_apply(xk, M, tc);
                            if (a.v_transform) se3_apply(xv, M, tc);
                        }
                    } else if (seg == 2) {
                        float W[34];
#pragma unroll
                        for (int i = 0; i < 17; ++i) {
                            float2 q2 = __ldg(reinterpret_cast<const float2*>(so3) + i);
                            W[2 * i] = q2.x; W[2 * i + 1] = q2.y;
                        }
                        so3_apply<false>(xk, W);
                        if (a.v_transform) so3_apply<false>(xv, W);
                    } else {
                        const So2Chunk sc = load_so2_chunk(so2, ch[u], a.hd);
                        const float cs[8] = {sc.a.x, sc.a.y, sc.a.z, sc.a.w, sc.b.x, sc.b.y, sc.b.z, sc.b.w};
                        so2_apply<false>(xk, cs);
                        if (a.v_transform) so2_apply<false>(xv, cs);
                    }
                }
                const uint32_t off = tile_sw64_offset(row[u], ch[u]);
                if (a.lo_offset) {
                    uint4 hi, lo;
                    split_chunk_bf16(xk, hi, lo);
                    *reinterpret_cast<uint4*>(kdst + off) = hi;
                    *reinterpret_cast<uint4*>(kdst + a.lo_offset + off) = lo;
                    split_chunk_bf16(xv, hi, lo);
                    *reinterpret_cast<uint4*>(vdst + off) = hi;
                    *reinterpret_cast<uint4*>(vdst + a.lo_offset + off) = lo;
                } else {
                    *reinterpret_cast<uint4*>(kdst + off) = pack_chunk_bf16(xk);
                    *reinterpret_cast<uint4*>(vdst + off) = pack_chunk_bf16(xv);
                }
            }
        }
        cbase += n_t;
    }
}

// K/V side arguments of a forward call (workspace = [K' | V'] or, split precision, [K'hi | K'lo | V'hi | V'lo]).
inline RotArgs make_rot_args_kv(const GtaAttnParams& p) {
    RotArgs a;
    a.k = p.k; a.v = p.v;
    a.k_sb = p.k_stride_b; a.k_sh = p.k_stride_h; a.k_st = p.k_stride_t;
    a.v_sb = p.v_stride_b; a.v_sh = p.v_stride_h; a.v_st = p.v_stride_t;
    a.ntiles = num_kv_tiles(p.Tk);
    const size_t half = static_cast<size_t>(p.B) * p.H * a.ntiles * kv_tile_bytes(p.D);
    const bool hp = attn_is_split_precision(p);
    a.ws_k = static_cast<uint8_t*>(p.workspace);
    a.ws_v = a.ws_k + (hp ? 2 * half : half);
    a.lo_offset = hp ? half : 0;
    a.H = p.H; a.Tk = p.Tk; a.D = p.D; a.Nk = p.Nk; a.tpv = p.Tk / p.Nk;
    a.hd = HeadDims{p.triv, p.se3, p.so3, p.so2};
    a.se3_k = p.reps.se3_k; a.so3_k = p.reps.so3_k; a.so2_k = p.reps.so2_k;
    a.tc_ptr = p.trans_coeff; a.C = p.so2 >> 1; a.v_transform = p.v_transform;
    return a;
}

}  // namespace gta
