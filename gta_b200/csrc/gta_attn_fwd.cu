// Fused GTA attention forward for sm_100a.
//
//   O = rho_q^{-1} softmax((rho_q^{-T} Q)(K')^T * scale) V'      K' = rho_k K, V' = rho_k V (staged tiles)
//
// One CTA per (batch, head, 128-query tile); 6 warps:
//   warps 0-3  softmax/correction/epilogue — thread i owns query row i == TMEM lane i.  Prologue: load the raw
//              (strided) Q row, apply rho_q^{-T} in fp32 registers, write the bf16 UMMA operand tile.
//              Epilogue: O/l, apply rho_q^{-1} in registers, store [B,Tq,H,D].
//   warp 4     UMMA issuer (one lane): S = Q'K'^T (SS, M=128,N=128,K=16 x D/16) into a double-buffered TMEM
//              accumulator, O += P V' (M=128,N=D,K=16 x 8) with P from shared memory (SS) or tensor memory (TS).
//   warp 5     bulk-copy producer: one cp.async.bulk per K'/V' tile image into a 2-stage ring.
// All producer/consumer hand-offs are mbarriers; tcgen05.commit signals MMA completion.
//
// Reference semantics: source/utils/gta.py:92-279 and source/layers.py:202-211.
#include <cmath>

#include "attn_common.cuh"

namespace gta {

constexpr int kThreads = 192;
constexpr int kStages = 2;
constexpr uint32_t kTmemS0 = 0, kTmemS1 = 128, kTmemO = 256;

enum BarIdx {
    kBarQFull = 0,
    kBarKFull = 1,                       // [kStages]
    kBarVFull = kBarKFull + kStages,     // [kStages]
    kBarKEmpty = kBarVFull + kStages,    // [kStages]
    kBarVEmpty = kBarKEmpty + kStages,   // [kStages]
    kBarSFull = kBarVEmpty + kStages,    // [2]
    kBarPFull = kBarSFull + 2,           // [2]
    kBarPVDone = kBarPFull + 2,
    kNumBars
};

template <int D, bool P_TMEM>
struct AttnSmem {
    static constexpr uint32_t kTile = 128u * D * 2u;
    static constexpr uint32_t kQ = 0;
    static constexpr uint32_t kK = kTile;
    static constexpr uint32_t kV = kTile * (1 + kStages);
    static constexpr uint32_t kP = kTile * (1 + 2 * kStages);
    static constexpr uint32_t kBars = kP + (P_TMEM ? 0u : 32768u);
    static constexpr uint32_t kTmemSlot = kBars + kNumBars * 8;
    static constexpr uint32_t kUsed = kTmemSlot + 16;
    // >= 120 KB so that only one CTA (one 512-column TMEM allocation) is resident per SM
    static constexpr uint32_t kBytes = (kUsed + 1024 > 120u * 1024u) ? kUsed + 1024 : 120u * 1024u;
};

template <typename TIn, typename TOut, int D, bool P_TMEM>
__global__ void __launch_bounds__(kThreads, 1) attn_fwd_kernel(const AttnArgs a) {
    using L = AttnSmem<D, P_TMEM>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBars);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kTmemSlot);

    const int qtile = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = a.ntiles_k;

    if (threadIdx.x == 0) {
        mbar_init(&bars[kBarQFull], 128);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&bars[kBarKFull + s], 1);
            mbar_init(&bars[kBarVFull + s], 1);
            mbar_init(&bars[kBarKEmpty + s], 1);
            mbar_init(&bars[kBarVEmpty + s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars[kBarSFull + i], 1);
            mbar_init(&bars[kBarPFull + i], 128);
        }
        mbar_init(&bars[kBarPVDone], 1);
        fence_mbar_init();
    }
    if (warp == 4) {
        tmem_alloc(tmem_slot, kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

    if (warp < 4) {
        // =========================================================== softmax / correction / epilogue
        const int r = threadIdx.x;
        const int t = qtile * 128 + r;
        const bool valid = t < a.Tq;
        const int tt = valid ? t : a.Tq - 1;
        const float tc = a.tc_ptr ? __ldg(a.tc_ptr) : 1.0f;
        const size_t view = static_cast<size_t>(b) * a.Nq + tt / a.tpvq;
        const float* se3 = a.se3_q + view * 16;
        const float* so3 = a.so3_q + view * 34;
        const float* so2 = a.so2_q + (static_cast<size_t>(b) * a.Tq + tt) * a.C * 2;

        {   // ---- Q prologue: raw row -> rho_q^{-T} -> bf16 operand tile
            const TIn* qrow = reinterpret_cast<const TIn*>(a.q) + static_cast<int64_t>(b) * a.q_sb +
                              static_cast<int64_t>(h) * a.q_sh + static_cast<int64_t>(tt) * a.q_st;
#pragma unroll 1
            for (int c = 0; c < D / 8; ++c) {
                float x[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = 0.f;
                if (valid) {
                    load_chunk<TIn>(qrow + c * 8, x);
                    apply_rep_chunk<kModeQ>(x, c, a.hd, se3, so3, so2, tc);
                }
                *reinterpret_cast<uint4*>(smem + L::kQ + tile_sw64_offset(r, c)) = pack_chunk_bf16(x);
            }
            fence_proxy_async_smem();
            mbar_arrive(&bars[kBarQFull]);
        }

        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
        const float cs = a.scale_log2;
        float m_run = -INFINITY, l_run = 0.f;

#pragma unroll 1
        for (int j = 0; j < n; ++j) {
            const int buf = j & 1;
            mbar_wait(&bars[kBarSFull + buf], (j >> 1) & 1);
            tc_fence_after();
            uint32_t sreg[128];
            const uint32_t s_addr = lane_base + (buf ? kTmemS1 : kTmemS0);
            tmem_ld32(s_addr, sreg);
            tmem_ld32(s_addr + 32, sreg + 32);
            tmem_ld32(s_addr + 64, sreg + 64);
            tmem_ld32(s_addr + 96, sreg + 96);
            tmem_ld_wait();
            float* s = reinterpret_cast<float*>(sreg);
            if (j == n - 1) {
                const int nvalid = a.Tk - j * 128;
                if (nvalid < 128) {
#pragma unroll
                    for (int i = 0; i < 128; ++i) if (i >= nvalid) s[i] = -INFINITY;
                }
            }
            float mx0 = s[0], mx1 = s[1], mx2 = s[2], mx3 = s[3];
#pragma unroll
            for (int i = 4; i < 128; i += 4) {
                mx0 = fmaxf(mx0, s[i]); mx1 = fmaxf(mx1, s[i + 1]);
                mx2 = fmaxf(mx2, s[i + 2]); mx3 = fmaxf(mx3, s[i + 3]);
            }
            const float m_new = fmaxf(m_run, fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)));
            const float alpha = fast_exp2((m_run - m_new) * cs);

            if (j > 0) {
                // O of tiles < j must be complete before it is rescaled / before P is overwritten.
                mbar_wait(&bars[kBarPVDone], (j - 1) & 1);
                tc_fence_after();
                if (__any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll
                    for (int cb = 0; cb < D / 32; ++cb) {
                        uint32_t o[32];
                        tmem_ld32(lane_base + kTmemO + cb * 32, o);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                        tmem_st32(lane_base + kTmemO + cb * 32, o);
                    }
                    tmem_st_wait();
                }
            }

            const float neg = -m_new * cs;
            float ls0 = 0.f, ls1 = 0.f;
            if (P_TMEM) {
                // P (bf16, two keys per 32-bit column) overwrites the first 64 columns of this S buffer.
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t pr[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        float p0 = fast_exp2(fmaf(s[half * 64 + 2 * i], cs, neg));
                        float p1 = fast_exp2(fmaf(s[half * 64 + 2 * i + 1], cs, neg));
                        ls0 += p0; ls1 += p1;
                        pr[i] = pack_bf16x2(p0, p1);
                    }
                    tmem_st32(s_addr + half * 32, pr);
                }
                tmem_st_wait();
            } else {
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    float p[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) p[i] = fast_exp2(fmaf(s[c * 8 + i], cs, neg));
                    ls0 += (p[0] + p[2]) + (p[4] + p[6]);
                    ls1 += (p[1] + p[3]) + (p[5] + p[7]);
                    *reinterpret_cast<uint4*>(smem + L::kP + tile_sw128_offset(r, c)) = pack_chunk_bf16(p);
                }
                fence_proxy_async_smem();
            }
            l_run = fmaf(l_run, alpha, ls0 + ls1);
            m_run = m_new;
            tc_fence_before();
            mbar_arrive(&bars[kBarPFull + buf]);
        }

        // ---- epilogue: O / l, rho_q^{-1}, store [B,Tq,H,D]
        mbar_wait(&bars[kBarPVDone], (n - 1) & 1);
        tc_fence_after();
        const float inv_l = 1.0f / l_run;
        TOut* orow = reinterpret_cast<TOut*>(a.out) + ((static_cast<int64_t>(b) * a.Tq + tt) * a.H + h) * D;
#pragma unroll 1
        for (int cb = 0; cb < D / 32; ++cb) {
            uint32_t o[32];
            tmem_ld32(lane_base + kTmemO + cb * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                float x[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = __uint_as_float(o[cc * 8 + i]) * inv_l;
                const int c = cb * 4 + cc;
                if (a.v_transform) apply_rep_chunk<kModeOut>(x, c, a.hd, se3, so3, so2, tc);
                if (valid) store_chunk<TOut>(orow + c * 8, x);
            }
        }
        if (a.lse && valid)
            a.lse[(static_cast<int64_t>(b) * a.H + h) * a.Tq + t] = m_run * a.scale + logf(l_run);
        tc_fence_before();
    } else if (warp == 4) {
        // =========================================================== UMMA issuer
        constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);
        constexpr uint32_t idesc_pv = make_idesc_bf16(128, D, 0, 1);
        const uint32_t q_addr = smem_u32(smem + L::kQ);
        const uint32_t p_addr = smem_u32(smem + L::kP);
        mbar_wait(&bars[kBarQFull], 0);
        tc_fence_after();

        auto issue_qk = [&](int j) {
            const int s = j % kStages;
            mbar_wait(&bars[kBarKFull + s], (j / kStages) & 1);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t k_addr = smem_u32(smem + L::kK + s * L::kTile);
                const uint32_t d_addr = tmem_base + ((j & 1) ? kTmemS1 : kTmemS0);
#pragma unroll
                for (int kk = 0; kk < D / 16; ++kk)
                    umma_ss(d_addr, desc_kmajor_sw64(q_addr, kk), desc_kmajor_sw64(k_addr, kk), idesc_qk, kk > 0);
                umma_commit(&bars[kBarKEmpty + s]);
                umma_commit(&bars[kBarSFull + (j & 1)]);
            }
            __syncwarp();
        };
        auto issue_pv = [&](int j) {
            const int s = j % kStages;
            mbar_wait(&bars[kBarVFull + s], (j / kStages) & 1);
            mbar_wait(&bars[kBarPFull + (j & 1)], (j >> 1) & 1);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t v_addr = smem_u32(smem + L::kV + s * L::kTile);
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    const uint32_t acc = (j > 0 || kk > 0) ? 1u : 0u;
                    if (P_TMEM)
                        umma_ts(tmem_base + kTmemO, tmem_base + ((j & 1) ? kTmemS1 : kTmemS0) + kk * 8,
                                desc_mnmajor_sw64(v_addr, kk), idesc_pv, acc);
                    else
                        umma_ss(tmem_base + kTmemO, desc_p_sw128(p_addr, kk), desc_mnmajor_sw64(v_addr, kk),
                                idesc_pv, acc);
                }
                umma_commit(&bars[kBarVEmpty + s]);
                umma_commit(&bars[kBarPVDone]);
            }
            __syncwarp();
        };
        issue_qk(0);
#pragma unroll 1
        for (int j = 0; j < n; ++j) {
            if (j + 1 < n) issue_qk(j + 1);
            issue_pv(j);
        }
    } else {
        // =========================================================== bulk-copy producer
        const size_t blob0 = (static_cast<size_t>(b) * a.H + h) * n;
#pragma unroll 1
        for (int j = 0; j < n; ++j) {
            const int s = j % kStages;
            if (j >= kStages) mbar_wait(&bars[kBarKEmpty + s], ((j / kStages) - 1) & 1);
            if (lane == 0) {
                mbar_arrive_expect_tx(&bars[kBarKFull + s], L::kTile);
                bulk_g2s(smem + L::kK + s * L::kTile, a.ws_k + (blob0 + j) * L::kTile, L::kTile, &bars[kBarKFull + s]);
            }
            if (j >= kStages) mbar_wait(&bars[kBarVEmpty + s], ((j / kStages) - 1) & 1);
            if (lane == 0) {
                mbar_arrive_expect_tx(&bars[kBarVFull + s], L::kTile);
                bulk_g2s(smem + L::kV + s * L::kTile, a.ws_v + (blob0 + j) * L::kTile, L::kTile, &bars[kBarVFull + s]);
            }
            __syncwarp();
        }
    }

    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

template <typename TIn, typename TOut, int D, bool P_TMEM>
static int launch_one(const AttnArgs& a, dim3 grid, cudaStream_t st) {
    using L = AttnSmem<D, P_TMEM>;
    auto kern = attn_fwd_kernel<TIn, TOut, D, P_TMEM>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(L::kBytes));
    if (e != cudaSuccess) return set_error(GTA_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    kern<<<grid, kThreads, L::kBytes, st>>>(a);
    return check_launch("gta_attn_fwd");
}

template <typename TIn, typename TOut, bool P_TMEM>
static int launch_d(const AttnArgs& a, int D, dim3 grid, cudaStream_t st) {
    switch (D) {
        case 32: return launch_one<TIn, TOut, 32, P_TMEM>(a, grid, st);
        case 64: return launch_one<TIn, TOut, 64, P_TMEM>(a, grid, st);
        case 96: return launch_one<TIn, TOut, 96, P_TMEM>(a, grid, st);
        case 128: return launch_one<TIn, TOut, 128, P_TMEM>(a, grid, st);
    }
    return set_error(GTA_ERR_UNSUPPORTED, "gta_attn_fwd: head dim %d not in {32,64,96,128}", D);
}

template <bool P_TMEM>
static int launch_t(const GtaAttnParams& p, const AttnArgs& a, dim3 grid, cudaStream_t st) {
    const bool ib = p.in_dtype == GTA_DTYPE_BF16, ob = p.out_dtype == GTA_DTYPE_BF16;
    if (ib && ob) return launch_d<__nv_bfloat16, __nv_bfloat16, P_TMEM>(a, p.D, grid, st);
    if (ib && !ob) return launch_d<__nv_bfloat16, float, P_TMEM>(a, p.D, grid, st);
    if (!ib && ob) return launch_d<float, __nv_bfloat16, P_TMEM>(a, p.D, grid, st);
    return launch_d<float, float, P_TMEM>(a, p.D, grid, st);
}

int launch_attn_fwd_v0(const GtaAttnParams& p, cudaStream_t st) {
    const AttnArgs a = make_attn_args(p);
    dim3 grid((p.Tq + 127) / 128, p.H, p.B);
    if (p.flags & 1 /* P operand in tensor memory */) return launch_t<true>(p, a, grid, st);
    return launch_t<false>(p, a, grid, st);
}

// ===================================================================================================
// tcgen05 self-test: one CTA, S = A B^T and O = P V through the same tile images + descriptors.
template <int D, bool P_TMEM>
__global__ void __launch_bounds__(128, 1) umma_probe_kernel(const __nv_bfloat16* __restrict__ A,
                                                           const __nv_bfloat16* __restrict__ Bm,
                                                           const __nv_bfloat16* __restrict__ P,
                                                           const __nv_bfloat16* __restrict__ V, float* __restrict__ outS,
                                                           float* __restrict__ outO) {
    constexpr uint32_t kTile = 128u * D * 2u;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem; uint8_t* sB = smem + kTile; uint8_t* sV = smem + 2 * kTile; uint8_t* sP = smem + 3 * kTile;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 3 * kTile + 32768);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
    const int r = threadIdx.x, warp = r >> 5;
    if (r == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
    if (warp == 0) { tmem_alloc(slot, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = *reinterpret_cast<volatile uint32_t*>(slot);
    const uint32_t lane_base = tb + (static_cast<uint32_t>(warp * 32) << 16);
    for (int c = 0; c < D / 8; ++c) {
        *reinterpret_cast<uint4*>(sA + tile_sw64_offset(r, c)) = *reinterpret_cast<const uint4*>(A + r * D + c * 8);
        *reinterpret_cast<uint4*>(sB + tile_sw64_offset(r, c)) = *reinterpret_cast<const uint4*>(Bm + r * D + c * 8);
        *reinterpret_cast<uint4*>(sV + tile_sw64_offset(r, c)) = *reinterpret_cast<const uint4*>(V + r * D + c * 8);
    }
    if (P_TMEM) {
        uint32_t pr[32];
        for (int half = 0; half < 2; ++half) {
            for (int i = 0; i < 32; ++i) pr[i] = *reinterpret_cast<const uint32_t*>(P + r * 128 + half * 64 + 2 * i);
            tmem_st32(lane_base + 128 + half * 32, pr);
        }
        tmem_st_wait();
    } else {
        for (int c = 0; c < 16; ++c)
            *reinterpret_cast<uint4*>(sP + tile_sw128_offset(r, c)) = *reinterpret_cast<const uint4*>(P + r * 128 + c * 8);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (r == 0) {
        constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);
        constexpr uint32_t idesc_pv = make_idesc_bf16(128, D, 0, 1);
        for (int kk = 0; kk < D / 16; ++kk)
            umma_ss(tb + 0, desc_kmajor_sw64(smem_u32(sA), kk), desc_kmajor_sw64(smem_u32(sB), kk), idesc_qk, kk > 0);
        umma_commit(&bar[0]);
        for (int kk = 0; kk < 8; ++kk) {
            if (P_TMEM) umma_ts(tb + 256, tb + 128 + kk * 8, desc_mnmajor_sw64(smem_u32(sV), kk), idesc_pv, kk > 0);
            else umma_ss(tb + 256, desc_p_sw128(smem_u32(sP), kk), desc_mnmajor_sw64(smem_u32(sV), kk), idesc_pv, kk > 0);
        }
        umma_commit(&bar[1]);
    }
    __syncwarp();
    mbar_wait(&bar[0], 0);
    mbar_wait(&bar[1], 0);
    tc_fence_after();
    for (int cb = 0; cb < 4; ++cb) {
        uint32_t o[32];
        tmem_ld32(lane_base + cb * 32, o);
        tmem_ld_wait();
        for (int i = 0; i < 32; ++i) outS[r * 128 + cb * 32 + i] = __uint_as_float(o[i]);
    }
    for (int cb = 0; cb < D / 32; ++cb) {
        uint32_t o[32];
        tmem_ld32(lane_base + 256 + cb * 32, o);
        tmem_ld_wait();
        for (int i = 0; i < 32; ++i) outO[r * D + cb * 32 + i] = __uint_as_float(o[i]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

template <int D, bool P_TMEM>
static int probe_one(const void* A, const void* Bm, const void* P, const void* V, float* outS, float* outO,
                     cudaStream_t st) {
    auto kern = umma_probe_kernel<D, P_TMEM>;
    const int bytes = 3 * 128 * D * 2 + 32768 + 64 + 1024;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return set_error(GTA_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    kern<<<1, 128, bytes, st>>>(static_cast<const __nv_bfloat16*>(A), static_cast<const __nv_bfloat16*>(Bm),
                                static_cast<const __nv_bfloat16*>(P), static_cast<const __nv_bfloat16*>(V), outS, outO);
    return check_launch("gta_umma_probe");
}

int launch_umma_probe(const void* A, const void* Bm, const void* P, const void* V, int D, int p_in_tmem, float* outS,
                      float* outO, cudaStream_t st) {
#define GTA_PROBE_CASE(DD)                                                              \
    case DD:                                                                            \
        return p_in_tmem ? probe_one<DD, true>(A, Bm, P, V, outS, outO, st)             \
                         : probe_one<DD, false>(A, Bm, P, V, outS, outO, st);
    switch (D) {
        GTA_PROBE_CASE(32)
        GTA_PROBE_CASE(64)
        GTA_PROBE_CASE(96)
        GTA_PROBE_CASE(128)
    }
#undef GTA_PROBE_CASE
    return set_error(GTA_ERR_UNSUPPORTED, "gta_umma_probe: head dim %d not in {32,64,96,128}", D);
}


// ===================================================================================================
// UMMA throughput micro-benchmark (tools/umma_bench.py): one CTA per SM, one thread issues `reps` repetitions of an
// MMA group on garbage operands and measures clock64 from first issue to the completion commit.
//   mode 0: S = Q K^T group, SS, M=128 N=128, K-steps D/16 (both operands K-major, 64B swizzle)
//   mode 1: same with N=64
//   mode 2: O += P V group, TS (A in TMEM), M=128 N=D, 8 K-steps (B MN-major, 64B swizzle)
//   mode 3: O += P V group, SS (A = P tile in smem, 128B swizzle), 8 K-steps
//   mode 4: mode 2 with 4 K-steps (64-key half tile)
template <int D>
__global__ void __launch_bounds__(128, 1) umma_bench_kernel(int mode, int reps, long long* out) {
    constexpr uint32_t kTile = 128u * D * 2u;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 3 * kTile + 32768);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
    const int warp = threadIdx.x >> 5;
    for (uint32_t i = threadIdx.x; i < (3 * kTile + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); fence_mbar_init(); }
    if (warp == 0) { tmem_alloc(slot, 512); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = *reinterpret_cast<volatile uint32_t*>(slot);
    if (threadIdx.x == 0) {
        const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + kTile), v_addr = smem_u32(smem + 2 * kTile);
        const uint32_t p_addr = smem_u32(smem + 3 * kTile);
        constexpr uint32_t id128 = make_idesc_bf16(128, 128, 0, 0), id64 = make_idesc_bf16(128, 64, 0, 0);
        constexpr uint32_t idpv = make_idesc_bf16(128, D, 0, 1);
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            if (mode == 0) {
#pragma unroll
                for (int kk = 0; kk < D / 16; ++kk)
                    umma_ss(tb + (r & 1) * 128, desc_kmajor_sw64(a_addr, kk), desc_kmajor_sw64(b_addr, kk), id128, kk > 0);
            } else if (mode == 1) {
#pragma unroll
                for (int kk = 0; kk < D / 16; ++kk)
                    umma_ss(tb + (r & 3) * 64, desc_kmajor_sw64(a_addr, kk), desc_kmajor_sw64(b_addr + (r & 1) * 4096, kk), id64, kk > 0);
            } else if (mode == 2) {
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)
                    umma_ts(tb + 256 + (r & 1) * 128, tb + kk * 8, desc_mnmajor_sw64(v_addr, kk), idpv, 1u);
            } else if (mode == 3) {
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)
                    umma_ss(tb + 256 + (r & 1) * 128, desc_p_sw128(p_addr, kk), desc_mnmajor_sw64(v_addr, kk), idpv, 1u);
            } else {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    umma_ts(tb + 256 + (r & 1) * 128, tb + kk * 8, desc_mnmajor_sw64(v_addr, (r & 1) * 4 + kk), idpv, 1u);
            }
        }
        const long long t1 = clock64();
        umma_commit(&bar[0]);
        mbar_wait(&bar[0], 0);
        const long long t2 = clock64();
        out[blockIdx.x * 2] = t1 - t0;      // issue time
        out[blockIdx.x * 2 + 1] = t2 - t0;  // until everything completed
    }
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

int launch_umma_bench(int D, int mode, int reps, int grid, long long* out, cudaStream_t st) {
    const int bytes = 3 * 128 * D * 2 + 32768 + 64 + 1024;
#define GTA_UB_CASE(DD)                                                                                          \
    case DD: {                                                                                                   \
        auto kern = umma_bench_kernel<DD>;                                                                       \
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);                          \
        kern<<<grid, 128, bytes, st>>>(mode, reps, out);                                                         \
        return check_launch("gta_umma_bench");                                                                   \
    }
    switch (D) {
        GTA_UB_CASE(64)
        GTA_UB_CASE(96)
        GTA_UB_CASE(128)
    }
#undef GTA_UB_CASE
    return set_error(GTA_ERR_UNSUPPORTED, "gta_umma_bench: D must be 64, 96 or 128");
}


// ===================================================================================================
// Softmax inner-block micro-benchmark (tools/softmax_bench.py): the exp2 / row-sum / bf16-pack phase of one
// 128-column score row per thread, repeated `reps` times, with POLY of every DEN pairs evaluated by poly_exp2x2.
template <int NUM, int DEN>
__global__ void softmax_bench_kernel(const float* __restrict__ in, float* __restrict__ out, int reps, long long* clk) {
    float s[128];
#pragma unroll
    for (int i = 0; i < 128; ++i) s[i] = in[(threadIdx.x * 128 + i) & 1023];
    const float cs = 0.1f, neg = -0.3f;
    const uint64_t cs2 = pack_f32x2(cs, cs), neg2 = pack_f32x2(neg, neg);
    uint32_t acc = 0;
    uint64_t lsum2 = pack_f32x2(0.f, 0.f);
    __syncthreads();
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int i = 0; i < 64; ++i) {
            const uint64_t x2 = ffma2(pack_f32x2(s[2 * i], s[2 * i + 1]), cs2, neg2);
            float p0, p1;
            if ((i % DEN) < NUM) {
                poly_exp2x2(x2, p0, p1);
            } else {
                float x0, x1;
                unpack_f32x2(x2, x0, x1);
                p0 = fast_exp2(x0); p1 = fast_exp2(x1);
            }
            lsum2 = fadd2(lsum2, pack_f32x2(p0, p1));
            acc ^= pack_bf16x2(p0, p1);
            s[2 * i] = p0 - 1.0f;           // feed back so that the loop cannot be hoisted
        }
    }
    const long long t1 = clock64();
    float l0, l1;
    unpack_f32x2(lsum2, l0, l1);
    out[blockIdx.x * blockDim.x + threadIdx.x] = l0 + l1 + __uint_as_float(acc);
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

int launch_softmax_bench(int num, int den, int warps, int reps, int grid, const float* in, float* out, long long* clk,
                         cudaStream_t st) {
#define GTA_SB(N_, D_) if (num == N_ && den == D_) { softmax_bench_kernel<N_, D_><<<grid, warps * 32, 0, st>>>(in, out, reps, clk); return check_launch("gta_softmax_bench"); }
    GTA_SB(0, 4) GTA_SB(1, 4) GTA_SB(1, 3) GTA_SB(1, 2) GTA_SB(2, 3) GTA_SB(1, 1)
#undef GTA_SB
    return set_error(GTA_ERR_UNSUPPORTED, "gta_softmax_bench: unsupported poly fraction");
}

}  // namespace gta
