// Block-diagonal representation application on 8-element chunks, in fp32 registers.
//
// A head row of D elements is [triv | se3 | so3 | so2] (reference: source/utils/gta.py:112-122); all
// block boundaries are multiples of 8 elements for every shipped config (se3 % 4, so3 % 8, so2 = 4*nfreq),
// and the host validates it, so one 16-byte bf16 chunk never straddles two block types:
//   se3 chunk = two 4-vectors, each multiplied by the view's 4x4 (gta.py:160-167, :255-257)
//   so3 chunk = [3 | 5] multiplied by Wigner D_1, D_2 of the view (gta.py:182-201, :259-268)
//   so2 chunk = four (x,y) pairs, each rotated by the token's angle (gta.py:203-219, :269-271)
#pragma once
#include <type_traits>

#include "ptx.cuh"

namespace gta {

enum RepMode : int {
    kModeQ = 0,    // rho_q^{-T}: (E_q*msk)^T, D_q, R(th_q)
    kModeKV = 1,   // rho_k:      inv(E_k)*msk, D_k, R(th_k)
    kModeOut = 2,  // rho_q^{-1}: E_q*msk, D_q^T, R(th_q)^T
    kModeKVT = 3,  // rho_k^T:    (inv(E_k)*msk)^T, D_k^T, R(th_k)^T   (backward: dK = rho_k^T dK', dV = rho_k^T dV')
};

struct HeadDims {
    int triv, se3, so3, so2;  // element counts; sum = D
};

// y = (M * scale_mask(tc)) x for two 4-vectors; M row-major 4x4 (unscaled), scale_mask multiplies the
// translation column (rows 0..2 of column 3) by tc and zeroes row 3 except M33 (gta.py:40-44).
__device__ __forceinline__ void se3_apply(float* x, const float* __restrict__ M, float tc) {
    const float m03 = M[3] * tc, m13 = M[7] * tc, m23 = M[11] * tc, m33 = M[15];
#pragma unroll
    for (int v = 0; v < 2; ++v) {
        float a = x[4 * v], b = x[4 * v + 1], c = x[4 * v + 2], d = x[4 * v + 3];
        x[4 * v + 0] = fmaf(M[0], a, fmaf(M[1], b, fmaf(M[2], c, m03 * d)));
        x[4 * v + 1] = fmaf(M[4], a, fmaf(M[5], b, fmaf(M[6], c, m13 * d)));
        x[4 * v + 2] = fmaf(M[8], a, fmaf(M[9], b, fmaf(M[10], c, m23 * d)));
        x[4 * v + 3] = m33 * d;
    }
}
// y = (M * scale_mask(tc))^T x
__device__ __forceinline__ void se3_apply_T(float* x, const float* __restrict__ M, float tc) {
    const float m33 = M[15];
#pragma unroll
    for (int v = 0; v < 2; ++v) {
        float a = x[4 * v], b = x[4 * v + 1], c = x[4 * v + 2], d = x[4 * v + 3];
        x[4 * v + 0] = fmaf(M[0], a, fmaf(M[4], b, M[8] * c));
        x[4 * v + 1] = fmaf(M[1], a, fmaf(M[5], b, M[9] * c));
        x[4 * v + 2] = fmaf(M[2], a, fmaf(M[6], b, M[10] * c));
        x[4 * v + 3] = fmaf(tc, fmaf(M[3], a, fmaf(M[7], b, M[11] * c)), m33 * d);
    }
}
// W = D1 (9, row-major) | D2 (25, row-major); x = [3 | 5].
template <bool kTranspose>
__device__ __forceinline__ void so3_apply(float* x, const float* __restrict__ W) {
    float y[8];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 3; ++j) s = fmaf(kTranspose ? W[j * 3 + i] : W[i * 3 + j], x[j], s);
        y[i] = s;
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 5; ++j) s = fmaf(kTranspose ? W[9 + j * 5 + i] : W[9 + i * 5 + j], x[3 + j], s);
        y[3 + i] = s;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = y[i];
}
// cs = (cos, sin) x 4 pairs; R = [[c,-s],[s,c]]; inverse uses R^T.
template <bool kInverse>
__device__ __forceinline__ void so2_apply(float* x, const float* __restrict__ cs) {
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        float c = cs[2 * p], s = kInverse ? -cs[2 * p + 1] : cs[2 * p + 1];
        float a = x[2 * p], b = x[2 * p + 1];
        x[2 * p] = fmaf(c, a, -s * b);
        x[2 * p + 1] = fmaf(s, a, c * b);
    }
}

// Applies the rep of `mode` to chunk c (8 elements) of a head row.
//   se3m: 16 floats of the token's view (E_q for kModeQ/kModeOut, inv(E_k) for kModeKV), unscaled
//   so3m: 34 floats of the token's view;  so2cs: the token's [C][2] (cos,sin) table
template <int kMode>
__device__ __forceinline__ void apply_rep_chunk(float* x, int c, const HeadDims& hd, const float* __restrict__ se3m,
                                                const float* __restrict__ so3m, const float* __restrict__ so2cs,
                                                float tc) {
    const int e = c * 8;
    if (e < hd.triv) return;
    if (e < hd.triv + hd.se3) {
        float M[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float4 r = __ldg(reinterpret_cast<const float4*>(se3m) + i);
            M[4 * i] = r.x; M[4 * i + 1] = r.y; M[4 * i + 2] = r.z; M[4 * i + 3] = r.w;
        }
        if (kMode == kModeQ || kMode == kModeKVT) se3_apply_T(x, M, tc); else se3_apply(x, M, tc);
        return;
    }
    if (e < hd.triv + hd.se3 + hd.so3) {
        float W[34];
#pragma unroll
        for (int i = 0; i < 17; ++i) {
            float2 r = __ldg(reinterpret_cast<const float2*>(so3m) + i);
            W[2 * i] = r.x; W[2 * i + 1] = r.y;
        }
        if (kMode == kModeOut || kMode == kModeKVT) so3_apply<true>(x, W); else so3_apply<false>(x, W);
        return;
    }
    {
        const int pc = (e - hd.triv - hd.se3 - hd.so3) >> 1;  // first pair index of this chunk
        float cs[8];
        float4 r0 = __ldg(reinterpret_cast<const float4*>(so2cs + 2 * pc));
        float4 r1 = __ldg(reinterpret_cast<const float4*>(so2cs + 2 * pc) + 1);
        cs[0] = r0.x; cs[1] = r0.y; cs[2] = r0.z; cs[3] = r0.w;
        cs[4] = r1.x; cs[5] = r1.y; cs[6] = r1.z; cs[7] = r1.w;
        if (kMode == kModeOut || kMode == kModeKVT) so2_apply<true>(x, cs); else so2_apply<false>(x, cs);
    }
}

// Same as apply_rep_chunk for TWO operands that share their rep data (chunk c of K and of V of one token, or of Q and dO):
// the view matrix / token angles are loaded once.  `rot_b` = false leaves xb untouched (v_transform = False).
template <int kMode>
__device__ __forceinline__ void apply_rep_chunk_pair(float* xa, float* xb, bool rot_b, int c, const HeadDims& hd,
                                                     const float* __restrict__ se3m, const float* __restrict__ so3m,
                                                     const float* __restrict__ so2cs, float tc) {
    const int e = c * 8;
    if (e < hd.triv) return;
    constexpr bool kT = (kMode == kModeQ || kMode == kModeKVT);        // se3: transposed matrix
    constexpr bool kInv = (kMode == kModeOut || kMode == kModeKVT);    // so3 / so2: transposed (= inverse)
    if (e < hd.triv + hd.se3) {
        float M[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float4 r = __ldg(reinterpret_cast<const float4*>(se3m) + i);
            M[4 * i] = r.x; M[4 * i + 1] = r.y; M[4 * i + 2] = r.z; M[4 * i + 3] = r.w;
        }
        if (kT) { se3_apply_T(xa, M, tc); if (rot_b) se3_apply_T(xb, M, tc); }
        else { se3_apply(xa, M, tc); if (rot_b) se3_apply(xb, M, tc); }
        return;
    }
    if (e < hd.triv + hd.se3 + hd.so3) {
        float W[34];
#pragma unroll
        for (int i = 0; i < 17; ++i) {
            float2 r = __ldg(reinterpret_cast<const float2*>(so3m) + i);
            W[2 * i] = r.x; W[2 * i + 1] = r.y;
        }
        so3_apply<kInv>(xa, W);
        if (rot_b) so3_apply<kInv>(xb, W);
        return;
    }
    {
        const int pc = (e - hd.triv - hd.se3 - hd.so3) >> 1;
        float cs[8];
        float4 r0 = __ldg(reinterpret_cast<const float4*>(so2cs + 2 * pc));
        float4 r1 = __ldg(reinterpret_cast<const float4*>(so2cs + 2 * pc) + 1);
        cs[0] = r0.x; cs[1] = r0.y; cs[2] = r0.z; cs[3] = r0.w;
        cs[4] = r1.x; cs[5] = r1.y; cs[6] = r1.z; cs[7] = r1.w;
        so2_apply<kInv>(xa, cs);
        if (rot_b) so2_apply<kInv>(xb, cs);
    }
}

// 8 consecutive elements of a row -> fp32 registers (bf16 or fp32 source; 16-byte / 32-byte aligned).
template <typename T>
__device__ __forceinline__ void load_chunk(const T* __restrict__ p, float* x);
template <>
__device__ __forceinline__ void load_chunk<__nv_bfloat16>(const __nv_bfloat16* __restrict__ p, float* x) {
    uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    x[0] = bf16_lo(v.x); x[1] = bf16_hi(v.x); x[2] = bf16_lo(v.y); x[3] = bf16_hi(v.y);
    x[4] = bf16_lo(v.z); x[5] = bf16_hi(v.z); x[6] = bf16_lo(v.w); x[7] = bf16_hi(v.w);
}
template <>
__device__ __forceinline__ void load_chunk<float>(const float* __restrict__ p, float* x) {
    float4 a = __ldg(reinterpret_cast<const float4*>(p));
    float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}
__device__ __forceinline__ uint4 pack_chunk_bf16(const float* x) {
    uint4 v;
    v.x = pack_bf16x2(x[0], x[1]); v.y = pack_bf16x2(x[2], x[3]);
    v.z = pack_bf16x2(x[4], x[5]); v.w = pack_bf16x2(x[6], x[7]);
    return v;
}
template <typename T>
__device__ __forceinline__ void store_chunk(T* p, const float* x);
template <>
__device__ __forceinline__ void store_chunk<__nv_bfloat16>(__nv_bfloat16* p, const float* x) {
    *reinterpret_cast<uint4*>(p) = pack_chunk_bf16(x);
}
template <>
__device__ __forceinline__ void store_chunk<float>(float* p, const float* x) {
    reinterpret_cast<float4*>(p)[0] = make_float4(x[0], x[1], x[2], x[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(x[4], x[5], x[6], x[7]);
}


// ------------------------------------------------------------------------------------------------
// Latency-friendly variants: the per-view matrices are loaded ONCE per row into registers and the per-chunk
// loads (raw data + SO(2) table) are issued for a whole group of chunks before any of them is consumed, so a row
// costs ~2 global-load round trips instead of one per chunk.
struct ViewReps {
    float M[16];   // se3 4x4 of the row's view (unscaled)
    float W[34];   // so3 D1 | D2 of the row's view
};
__device__ __forceinline__ void load_view_reps(ViewReps& vr, const HeadDims& hd, const float* __restrict__ se3m,
                                               const float* __restrict__ so3m) {
    if (hd.se3) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float4 r = __ldg(reinterpret_cast<const float4*>(se3m) + i);
            vr.M[4 * i] = r.x; vr.M[4 * i + 1] = r.y; vr.M[4 * i + 2] = r.z; vr.M[4 * i + 3] = r.w;
        }
    }
    if (hd.so3) {
#pragma unroll
        for (int i = 0; i < 17; ++i) {
            float2 r = __ldg(reinterpret_cast<const float2*>(so3m) + i);
            vr.W[2 * i] = r.x; vr.W[2 * i + 1] = r.y;
        }
    }
}
// (cos,sin) x 4 pairs of chunk c of a token (valid only when chunk c lies in the so2 block).
struct So2Chunk { float4 a, b; };
__device__ __forceinline__ So2Chunk load_so2_chunk(const float* __restrict__ so2cs, int c, const HeadDims& hd) {
    So2Chunk r;
    r.a = make_float4(1.f, 0.f, 1.f, 0.f); r.b = r.a;
    const int off = c * 8 - (hd.triv + hd.se3 + hd.so3);      // float index into the token's [C][2] table
    if (hd.so2 && off >= 0) {
        r.a = __ldg(reinterpret_cast<const float4*>(so2cs + off));
        r.b = __ldg(reinterpret_cast<const float4*>(so2cs + off) + 1);
    }
    return r;
}
template <int kMode>
__device__ __forceinline__ void apply_rep_chunk_pre(float* x, int c, const HeadDims& hd, const ViewReps& vr,
                                                    const So2Chunk& sc, float tc) {
    const int e = c * 8;
    if (e < hd.triv) return;
    if (e < hd.triv + hd.se3) {
        if (kMode == kModeQ || kMode == kModeKVT) se3_apply_T(x, vr.M, tc); else se3_apply(x, vr.M, tc);
        return;
    }
    if (e < hd.triv + hd.se3 + hd.so3) {
        if (kMode == kModeOut || kMode == kModeKVT) so3_apply<true>(x, vr.W); else so3_apply<false>(x, vr.W);
        return;
    }
    const float cs[8] = {sc.a.x, sc.a.y, sc.a.z, sc.a.w, sc.b.x, sc.b.y, sc.b.z, sc.b.w};
    if (kMode == kModeOut || kMode == kModeKVT) so2_apply<true>(x, cs); else so2_apply<false>(x, cs);
}

// Raw (unconverted) chunk loads so that several can be in flight before the first conversion.
template <typename T> struct RawChunk;
template <> struct RawChunk<__nv_bfloat16> { uint4 v; };
template <> struct RawChunk<float> { float4 a, b; };
__device__ __forceinline__ void load_raw(const __nv_bfloat16* __restrict__ p, RawChunk<__nv_bfloat16>& r) {
    r.v = __ldg(reinterpret_cast<const uint4*>(p));
}
__device__ __forceinline__ void load_raw(const float* __restrict__ p, RawChunk<float>& r) {
    r.a = __ldg(reinterpret_cast<const float4*>(p));
    r.b = __ldg(reinterpret_cast<const float4*>(p) + 1);
}
// The same loads for rows that are read once (raw q rows of the Q stager): allocate in L1 for the thread's own neighbouring
// chunks of the line, but as the first candidates for eviction, so that they do not push out the rep tables other warps reuse.
__device__ __forceinline__ void load_raw_stream(const __nv_bfloat16* __restrict__ p, RawChunk<__nv_bfloat16>& r) {
    asm volatile("ld.global.nc.L1::evict_first.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r.v.x), "=r"(r.v.y), "=r"(r.v.z), "=r"(r.v.w) : "l"(p));
}
__device__ __forceinline__ void load_raw_stream(const float* __restrict__ p, RawChunk<float>& r) {
    asm volatile("ld.global.nc.L1::evict_first.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w) : "l"(p));
    asm volatile("ld.global.nc.L1::evict_first.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w) : "l"(p + 4));
}
__device__ __forceinline__ void zero_raw(RawChunk<__nv_bfloat16>& r) { r.v = make_uint4(0, 0, 0, 0); }
__device__ __forceinline__ void zero_raw(RawChunk<float>& r) { r.a = make_float4(0, 0, 0, 0); r.b = r.a; }
__device__ __forceinline__ void raw_to_f32(const RawChunk<__nv_bfloat16>& r, float* x) {
    x[0] = bf16_lo(r.v.x); x[1] = bf16_hi(r.v.x); x[2] = bf16_lo(r.v.y); x[3] = bf16_hi(r.v.y);
    x[4] = bf16_lo(r.v.z); x[5] = bf16_hi(r.v.z); x[6] = bf16_lo(r.v.w); x[7] = bf16_hi(r.v.w);
}
__device__ __forceinline__ void raw_to_f32(const RawChunk<float>& r, float* x) {
    x[0] = r.a.x; x[1] = r.a.y; x[2] = r.a.z; x[3] = r.a.w; x[4] = r.b.x; x[5] = r.b.y; x[6] = r.b.z; x[7] = r.b.w;
}


// ------------------------------------------------------------------------------------------------
// One whole head row (thread = token), walked block type by block type in batches of kB 16-byte chunks.  Everything a
// batch needs (raw chunks of operand a and, with kPair, of operand b that shares its rep data; the view matrix of the
// block type or the token's angles) is requested before the first use, so a row costs one memory round trip per batch
// instead of one per chunk — the staging warps of the fused attention kernel have no other warps to hide latency behind.
//   store(c, xa, xb): called once per chunk c in [0, D/8) with the transformed values (xb only meaningful with kPair).
// `valid` = false writes zeros (rows past the end of the sequence).  `rot_b` = false leaves operand b untransformed.
// (Measured alternatives that were slower in the fused kernel, see DESIGN.md: per-chunk rep loads, 32-byte paired
// accesses with a generic segment walker, lanes along the row with per-item rep loads.)
template <typename TIn, int kMode, bool kPair, int kB, typename Store>
__device__ __forceinline__ void stage_row(const TIn* __restrict__ arow, const TIn* __restrict__ brow, const bool valid,
                                          const bool rot_b, const HeadDims& hd, const float* __restrict__ se3m,
                                          const float* __restrict__ so3m, const float* __restrict__ so2cs, const float tc,
                                          Store&& store) {
    constexpr bool kT = (kMode == kModeQ || kMode == kModeKVT);        // se3: transposed matrix
    constexpr bool kInv = (kMode == kModeOut || kMode == kModeKVT);    // so3 / so2: transposed (= inverse)
    const int c1 = hd.triv >> 3, c2 = c1 + (hd.se3 >> 3), c3 = c2 + (hd.so3 >> 3), c4 = c3 + (hd.so2 >> 3);
    RawChunk<TIn> ra[kB], rb[kPair ? kB : 1];
    auto load_batch = [&](int c0, int cend) {
#pragma unroll
        for (int i = 0; i < kB; ++i) {
            zero_raw(ra[i]);
            if (kPair) zero_raw(rb[i]);
            if (valid && c0 + i < cend) {
                load_raw(arow + (c0 + i) * 8, ra[i]);
                if (kPair) load_raw(brow + (c0 + i) * 8, rb[i]);
            }
        }
    };
    // ---- trivial block: pass-through
#pragma unroll 1
    for (int c0 = 0; c0 < c1; c0 += kB) {
        load_batch(c0, c1);
#pragma unroll
        for (int i = 0; i < kB; ++i) {
            if (c0 + i < c1) {
                float xa[8], xb[8];
                raw_to_f32(ra[i], xa);
                if (kPair) raw_to_f32(rb[i], xb);
                store(c0 + i, xa, xb);
            }
        }
    }
    // ---- SE(3): two 4-vectors per chunk, one 4x4 per view
    if (c2 > c1) {
        float M[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 r = __ldg(reinterpret_cast<const float4*>(se3m) + i);
            M[4 * i] = r.x; M[4 * i + 1] = r.y; M[4 * i + 2] = r.z; M[4 * i + 3] = r.w;
        }
#pragma unroll 1
        for (int c0 = c1; c0 < c2; c0 += kB) {
            load_batch(c0, c2);
#pragma unroll
            for (int i = 0; i < kB; ++i) {
                if (c0 + i < c2) {
                    float xa[8], xb[8];
                    raw_to_f32(ra[i], xa);
                    if (kPair) raw_to_f32(rb[i], xb);
                    if (valid) {
                        if (kT) se3_apply_T(xa, M, tc); else se3_apply(xa, M, tc);
                        if (kPair && rot_b) { if (kT) se3_apply_T(xb, M, tc); else se3_apply(xb, M, tc); }
                    }
                    store(c0 + i, xa, xb);
                }
            }
        }
    }
    // ---- SO(3): [3 | 5] per chunk, Wigner D_1 | D_2 per view
    if (c3 > c2) {
        float W[34];
#pragma unroll
        for (int i = 0; i < 17; ++i) {
            const float2 r = __ldg(reinterpret_cast<const float2*>(so3m) + i);
            W[2 * i] = r.x; W[2 * i + 1] = r.y;
        }
#pragma unroll 1
        for (int c0 = c2; c0 < c3; c0 += kB) {
            load_batch(c0, c3);
#pragma unroll
            for (int i = 0; i < kB; ++i) {
                if (c0 + i < c3) {
                    float xa[8], xb[8];
                    raw_to_f32(ra[i], xa);
                    if (kPair) raw_to_f32(rb[i], xb);
                    if (valid) {
                        so3_apply<kInv>(xa, W);
                        if (kPair && rot_b) so3_apply<kInv>(xb, W);
                    }
                    store(c0 + i, xa, xb);
                }
            }
        }
    }
    // ---- SO(2): four pairs per chunk, the token's (cos, sin) table
#pragma unroll 1
    for (int c0 = c3; c0 < c4; c0 += kB) {
        So2Chunk cs[kB];
#pragma unroll
        for (int i = 0; i < kB; ++i) {
            cs[i].a = make_float4(1.f, 0.f, 1.f, 0.f); cs[i].b = cs[i].a;
            if (valid && c0 + i < c4) {
                const float4* p4 = reinterpret_cast<const float4*>(so2cs + (c0 + i - c3) * 8);
                cs[i].a = __ldg(p4); cs[i].b = __ldg(p4 + 1);
            }
        }
        load_batch(c0, c4);
#pragma unroll
        for (int i = 0; i < kB; ++i) {
            if (c0 + i < c4) {
                float xa[8], xb[8];
                raw_to_f32(ra[i], xa);
                if (kPair) raw_to_f32(rb[i], xb);
                const float c8[8] = {cs[i].a.x, cs[i].a.y, cs[i].a.z, cs[i].a.w, cs[i].b.x, cs[i].b.y, cs[i].b.z, cs[i].b.w};
                so2_apply<kInv>(xa, c8);
                if (kPair && rot_b) so2_apply<kInv>(xb, c8);
                store(c0 + i, xa, xb);
            }
        }
    }
}


// ------------------------------------------------------------------------------------------------
// stage_row with the head layout known at compile time (LY = HeadLayout<triv, se3, so3, so2>): straight-line code, every
// raw chunk of a batch (the whole row for bf16) requested up front, view matrices loaded once, no per-chunk block-type
// logic.  kNB = number of batches the row is split into (1 for bf16, more for fp32 inputs to bound the raw registers).
template <int I, int N, typename F>
__device__ __forceinline__ void static_for_(F&& f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for_<I + 1, N>(f);
    }
}
template <typename TIn, typename LY, int kMode, bool kPair, int kNB, typename Store>
__device__ __forceinline__ void stage_row_ct(const TIn* __restrict__ arow, const TIn* __restrict__ brow, const bool valid,
                                             const bool rot_b, const float* __restrict__ se3m, const float* __restrict__ so3m,
                                             const float* __restrict__ so2cs, const float tc, Store&& store) {
    constexpr bool kT = (kMode == kModeQ || kMode == kModeKVT);
    constexpr bool kInv = (kMode == kModeOut || kMode == kModeKVT);
    constexpr int NC = LY::D / 8;
    static_assert(NC % kNB == 0, "batches must divide the row");
    constexpr int CB = NC / kNB;
    float M[LY::kSe3 ? 16 : 1], W[LY::kSo3 ? 34 : 1];
    static_for_<0, kNB>([&](auto ib) {
        constexpr int c0 = decltype(ib)::value * CB;
        RawChunk<TIn> ra[CB], rb[kPair ? CB : 1];
#pragma unroll
        for (int i = 0; i < CB; ++i) {
            zero_raw(ra[i]);
            if (kPair) zero_raw(rb[i]);
            if (valid) {
                load_raw(arow + (c0 + i) * 8, ra[i]);
                if (kPair) load_raw(brow + (c0 + i) * 8, rb[i]);
            }
        }
        static_for_<0, CB>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            constexpr int c = c0 + i;
            float xa[8], xb[8];
            raw_to_f32(ra[i], xa);
            if (kPair) raw_to_f32(rb[i], xb);
            if constexpr (c >= LY::c1 && c < LY::c2) {
                if constexpr (c == LY::c1) {                     // first se3 chunk of the row: the view matrix, once
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 r4 = __ldg(reinterpret_cast<const float4*>(se3m) + q);
                        M[4 * q] = r4.x; M[4 * q + 1] = r4.y; M[4 * q + 2] = r4.z; M[4 * q + 3] = r4.w;
                    }
                }
                if (valid) {
                    if (kT) se3_apply_T(xa, M, tc); else se3_apply(xa, M, tc);
                    if (kPair && rot_b) { if (kT) se3_apply_T(xb, M, tc); else se3_apply(xb, M, tc); }
                }
            } else if constexpr (c >= LY::c2 && c < LY::c3) {
                if constexpr (c == LY::c2) {
#pragma unroll
                    for (int q = 0; q < 17; ++q) {
                        const float2 r2 = __ldg(reinterpret_cast<const float2*>(so3m) + q);
                        W[2 * q] = r2.x; W[2 * q + 1] = r2.y;
                    }
                }
                if (valid) {
                    so3_apply<kInv>(xa, W);
                    if (kPair && rot_b) so3_apply<kInv>(xb, W);
                }
            } else if constexpr (c >= LY::c3) {
                const float4* p4 = reinterpret_cast<const float4*>(so2cs + (c - LY::c3) * 8);
                const float4 ca = __ldg(p4), cb = __ldg(p4 + 1);
                const float c8[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
                if (valid) {
                    so2_apply<kInv>(xa, c8);
                    if (kPair && rot_b) so2_apply<kInv>(xb, c8);
                }
            }
            store(c, xa, xb);
        });
    });
}

}  // namespace gta
