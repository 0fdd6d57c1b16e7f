// Split-precision ("fp32-accurate") fused GTA attention forward.
//
// fp32 callers of the reference get fp32 arithmetic (source/layers.py:207-211 under trainer.py's fp32 configs); the
// tensor cores have no fp32 mode, so every operand x is carried as bf16 hi + bf16 residual lo = bf16(x - hi) (~16 mantissa
// bits) and every product is evaluated as  hi*hi + hi*lo + lo*hi  with fp32 accumulation in TMEM:
//     S = Q'h K'h^T + Q'h K'l^T + Q'l K'h^T            O += Ph V'h + Ph V'l + Pl V'h
// Measured max-abs error vs the fp32 oracle: ~1e-5 (budget 1e-3), at 3x the tensor work of the bf16 path.
//
// Pipeline: the simple first-generation structure (one 128-query tile per CTA; warps 0-3 softmax + Q prologue + epilogue,
// warp 4 UMMA issuer, warp 5 bulk-copy producer), K'h/K'l double-buffered, V'h/V'l single-buffered, S double-buffered
// in TMEM with Ph|Pl written over the S buffer they were computed from.
// Reference semantics: source/utils/gta.py:92-279 and source/layers.py:202-211.
#include <cmath>

#include "attn_common.cuh"

namespace gta {

constexpr int kThreadsHp = 192;
constexpr uint32_t kHpTmemS0 = 0, kHpTmemS1 = 128, kHpTmemO = 256;

enum HpBar {
    hQFull = 0,
    hKFull = 1,      // [2]
    hKEmpty = 3,     // [2]
    hVFull = 5,
    hVEmpty = 6,
    hSFull = 7,      // [2]
    hPFull = 9,      // [2]
    hPVDone = 11,
    hNumBars = 12
};

template <int D>
struct HpSmem {
    static constexpr uint32_t kTile = 128u * D * 2u;
    static constexpr uint32_t kQh = 0, kQl = kTile;
    static constexpr uint32_t kK = 2 * kTile;             // [2 stages][hi, lo]
    static constexpr uint32_t kV = 6 * kTile;             // [hi, lo]
    static constexpr uint32_t kBars = 8 * kTile;
    static constexpr uint32_t kTmemSlot = kBars + hNumBars * 8;
    static constexpr uint32_t kUsed = kTmemSlot + 16;
    static constexpr uint32_t kBytes = (kUsed + 1024 > 120u * 1024u) ? kUsed + 1024 : 120u * 1024u;
};

__device__ __forceinline__ void split_bf16_chunk(const float* x, uint4& hi, uint4& lo) {
    hi = pack_chunk_bf16(x);
    float r[8];
    r[0] = x[0] - bf16_lo(hi.x); r[1] = x[1] - bf16_hi(hi.x); r[2] = x[2] - bf16_lo(hi.y); r[3] = x[3] - bf16_hi(hi.y);
    r[4] = x[4] - bf16_lo(hi.z); r[5] = x[5] - bf16_hi(hi.z); r[6] = x[6] - bf16_lo(hi.w); r[7] = x[7] - bf16_hi(hi.w);
    lo = pack_chunk_bf16(r);
}

template <typename TIn, typename TOut, int D>
__global__ void __launch_bounds__(kThreadsHp, 1) attn_fwd_hp_kernel(const AttnArgs a, const size_t lo_offset) {
    using L = HpSmem<D>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBars);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kTmemSlot);

    const int qtile = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = a.ntiles_k;

    if (threadIdx.x == 0) {
        mbar_init(&bars[hQFull], 128);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bars[hKFull + s], 1);
            mbar_init(&bars[hKEmpty + s], 1);
            mbar_init(&bars[hSFull + s], 1);
            mbar_init(&bars[hPFull + s], 128);
        }
        mbar_init(&bars[hVFull], 1);
        mbar_init(&bars[hVEmpty], 1);
        mbar_init(&bars[hPVDone], 1);
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

        {   // ---- Q prologue: raw row -> rho_q^{-T} (fp32) -> hi / residual operand tiles
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
                uint4 hi, lo;
                split_bf16_chunk(x, hi, lo);
                const uint32_t off = tile_sw64_offset(r, c);
                *reinterpret_cast<uint4*>(smem + L::kQh + off) = hi;
                *reinterpret_cast<uint4*>(smem + L::kQl + off) = lo;
            }
            fence_proxy_async_smem();
            mbar_arrive(&bars[hQFull]);
        }

        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
        const float cs = a.scale_log2;
        float m_run = -INFINITY, l_run = 0.f;

#pragma unroll 1
        for (int j = 0; j < n; ++j) {
            const int buf = j & 1;
            mbar_wait(&bars[hSFull + buf], (j >> 1) & 1);
            tc_fence_after();
            uint32_t sreg[128];
            const uint32_t s_addr = lane_base + (buf ? kHpTmemS1 : kHpTmemS0);
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
            const float alpha = exp2f((m_run - m_new) * cs);

            if (j > 0) {
                mbar_wait(&bars[hPVDone], (j - 1) & 1);      // O complete up to tile j-1; the P columns are free again
                tc_fence_after();
                if (__any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll 1
                    for (int c8 = 0; c8 < D / 8; ++c8) {
                        uint32_t o8[8];
                        tmem_ld8(lane_base + kHpTmemO + c8 * 8, o8);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 8; ++i) o8[i] = __float_as_uint(__uint_as_float(o8[i]) * alpha);
                        tmem_st8(lane_base + kHpTmemO + c8 * 8, o8);
                    }
                }
            }

            // P = exp2(s*cs - m*cs) in full precision (exp2f, not the approximate MUFU path), split into hi + residual:
            // Ph -> columns 0..63 of this S buffer, Pl -> columns 64..127.
            const float neg = -m_new * cs;
            float lsum = 0.f;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t plo[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float p0 = exp2f(fmaf(s[half * 64 + 2 * i], cs, neg));
                    const float p1 = exp2f(fmaf(s[half * 64 + 2 * i + 1], cs, neg));
                    lsum += p0 + p1;
                    const uint32_t hi = pack_bf16x2(p0, p1);
                    plo[i] = pack_bf16x2(p0 - bf16_lo(hi), p1 - bf16_hi(hi));
                    sreg[half * 64 + i] = hi;                 // in place: pair i is consumed before slot i is reused
                }
                tmem_st32(s_addr + half * 32, sreg + half * 64);
                tmem_st32(s_addr + 64 + half * 32, plo);
            }
            l_run = fmaf(l_run, alpha, lsum);
            m_run = m_new;
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&bars[hPFull + buf]);
        }

        // ---- epilogue: O / l, rho_q^{-1}, store [B,Tq,H,D]
        mbar_wait(&bars[hPVDone], (n - 1) & 1);
        tc_fence_after();
        const float inv_l = 1.0f / l_run;
        TOut* orow = reinterpret_cast<TOut*>(a.out) + ((static_cast<int64_t>(b) * a.Tq + tt) * a.H + h) * D;
#pragma unroll 1
        for (int c = 0; c < D / 8; ++c) {
            uint32_t o8[8];
            tmem_ld8(lane_base + kHpTmemO + c * 8, o8);
            tmem_ld_wait();
            float x[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = __uint_as_float(o8[i]) * inv_l;
            if (a.v_transform) apply_rep_chunk<kModeOut>(x, c, a.hd, se3, so3, so2, tc);
            if (valid) store_chunk<TOut>(orow + c * 8, x);
        }
        if (a.lse && valid)
            a.lse[(static_cast<int64_t>(b) * a.H + h) * a.Tq + t] = m_run * a.scale + logf(l_run);
        tc_fence_before();
    } else if (warp == 4) {
        // =========================================================== UMMA issuer
        constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);
        constexpr uint32_t idesc_pv = make_idesc_bf16(128, D, 0, 1);
        const uint32_t bar0 = smem_u32(bars);
        const uint64_t qhd = desc_kmajor_sw64(smem_u32(smem + L::kQh), 0), qld = desc_kmajor_sw64(smem_u32(smem + L::kQl), 0);
        const uint32_t qh_lo = static_cast<uint32_t>(qhd), ql_lo = static_cast<uint32_t>(qld), q_hi = static_cast<uint32_t>(qhd >> 32);
        mbar_wait(&bars[hQFull], 0);
        tc_fence_after();

        auto issue_qk = [&](int j) {
            const int s = j & 1;
            mbar_wait(&bars[hKFull + s], (j >> 1) & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t khd = desc_kmajor_sw64(smem_u32(smem + L::kK + (2 * s) * L::kTile), 0);
                const uint64_t kld = desc_kmajor_sw64(smem_u32(smem + L::kK + (2 * s + 1) * L::kTile), 0);
                const uint32_t kh_lo = static_cast<uint32_t>(khd), kl_lo = static_cast<uint32_t>(kld), k_hi = static_cast<uint32_t>(khd >> 32);
                const uint32_t d_addr = tmem_base + (s ? kHpTmemS1 : kHpTmemS0);
#pragma unroll
                for (int kk = 0; kk < D / 16; ++kk) {
                    const uint32_t st = kstep_kmajor_sw64(kk);
                    umma_ss_lohi(d_addr, qh_lo + st, q_hi, kh_lo + st, k_hi, idesc_qk, kk > 0);
                    umma_ss_lohi(d_addr, qh_lo + st, q_hi, kl_lo + st, k_hi, idesc_qk, 1u);
                    umma_ss_lohi(d_addr, ql_lo + st, q_hi, kh_lo + st, k_hi, idesc_qk, 1u);
                }
                umma_commit_addr(bar0 + (hKEmpty + s) * 8);
                umma_commit_addr(bar0 + (hSFull + s) * 8);
            }
            __syncwarp();
        };
        auto issue_pv = [&](int j) {
            mbar_wait(&bars[hVFull], j & 1);
            mbar_wait(&bars[hPFull + (j & 1)], (j >> 1) & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t vhd = desc_mnmajor_sw64(smem_u32(smem + L::kV), 0);
                const uint64_t vld = desc_mnmajor_sw64(smem_u32(smem + L::kV + L::kTile), 0);
                const uint32_t vh_lo = static_cast<uint32_t>(vhd), vl_lo = static_cast<uint32_t>(vld), v_hi = static_cast<uint32_t>(vhd >> 32);
                const uint32_t ph = tmem_base + ((j & 1) ? kHpTmemS1 : kHpTmemS0), pl = ph + 64;
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    const uint32_t st = kstep_mnmajor_sw64(kk);
                    umma_ts_lohi(tmem_base + kHpTmemO, ph + kk * 8, vh_lo + st, v_hi, idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
                    umma_ts_lohi(tmem_base + kHpTmemO, ph + kk * 8, vl_lo + st, v_hi, idesc_pv, 1u);
                    umma_ts_lohi(tmem_base + kHpTmemO, pl + kk * 8, vh_lo + st, v_hi, idesc_pv, 1u);
                }
                umma_commit_addr(bar0 + hVEmpty * 8);
                umma_commit_addr(bar0 + hPVDone * 8);
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
            const int s = j & 1;
            if (j >= 2) mbar_wait(&bars[hKEmpty + s], ((j >> 1) - 1) & 1);
            if (lane == 0) {
                mbar_arrive_expect_tx(&bars[hKFull + s], 2 * L::kTile);
                const uint8_t* src = a.ws_k + (blob0 + j) * L::kTile;
                bulk_g2s(smem + L::kK + (2 * s) * L::kTile, src, L::kTile, &bars[hKFull + s]);
                bulk_g2s(smem + L::kK + (2 * s + 1) * L::kTile, src + lo_offset, L::kTile, &bars[hKFull + s]);
            }
            if (j >= 1) mbar_wait(&bars[hVEmpty], (j - 1) & 1);
            if (lane == 0) {
                mbar_arrive_expect_tx(&bars[hVFull], 2 * L::kTile);
                const uint8_t* src = a.ws_v + (blob0 + j) * L::kTile;
                bulk_g2s(smem + L::kV, src, L::kTile, &bars[hVFull]);
                bulk_g2s(smem + L::kV + L::kTile, src + lo_offset, L::kTile, &bars[hVFull]);
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

template <typename TIn, typename TOut, int D>
static int launch_hp_one(const AttnArgs& a, size_t lo_offset, dim3 grid, cudaStream_t st) {
    using L = HpSmem<D>;
    auto kern = attn_fwd_hp_kernel<TIn, TOut, D>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(L::kBytes));
    if (e != cudaSuccess) return set_error(GTA_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    kern<<<grid, kThreadsHp, L::kBytes, st>>>(a, lo_offset);
    return check_launch("gta_attn_fwd (split precision)");
}

template <typename TOut>
static int launch_hp_d(const AttnArgs& a, size_t lo_offset, int D, dim3 grid, cudaStream_t st) {
    switch (D) {
        case 32: return launch_hp_one<float, TOut, 32>(a, lo_offset, grid, st);
        case 64: return launch_hp_one<float, TOut, 64>(a, lo_offset, grid, st);
        case 96: return launch_hp_one<float, TOut, 96>(a, lo_offset, grid, st);
    }
    return set_error(GTA_ERR_UNSUPPORTED,
                     "fp32-accurate path supports head dims 32/64/96 (use GTA_FLAG_FAST_FP32 or bf16 inputs for %d)", D);
}

int launch_attn_fwd_hp(const GtaAttnParams& p, cudaStream_t st) {
    AttnArgs a = make_attn_args(p);
    const size_t half = static_cast<size_t>(p.B) * p.H * a.ntiles_k * kv_tile_bytes(p.D);
    a.ws_k = static_cast<const uint8_t*>(p.workspace);      // [K'hi | K'lo | V'hi | V'lo]
    a.ws_v = a.ws_k + 2 * half;
    dim3 grid((p.Tq + 127) / 128, p.H, p.B);
    if (p.out_dtype == GTA_DTYPE_BF16) return launch_hp_d<__nv_bfloat16>(a, half, p.D, grid, st);
    return launch_hp_d<float>(a, half, p.D, grid, st);
}

}  // namespace gta
