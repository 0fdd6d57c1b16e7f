// Fused GTA attention forward, v1 pipeline: TWO 128-query tiles per CTA.
//
//   warps 0-3   softmax warpgroup A  (query tile A: rows 0..127 of the CTA's 256-row slab)
//   warps 4-7   softmax warpgroup B  (query tile B)
//   warp  8     UMMA issuer          (one lane issues every tcgen05.mma / tcgen05.commit)
//   warp  9     bulk-copy producer   (K'/V' tile images, kStages-deep ring)
//   warps 10-11 idle (keep the third warpgroup aligned for setmaxnreg)
//
// Tensor-core schedule per key tile j (all MMAs execute in issue order on the single tensor pipe):
//     PV_A(j)  QK_A(j+1)  PV_B(j)  QK_B(j+1)
// so while warpgroup A runs the softmax of S_A(j+1) the pipe is busy with tile B and vice versa.
// P (bf16) is written back into the first 64 columns of its own S accumulator in tensor memory and consumed
// from there as the A operand (TS form), so neither S nor P ever touches shared memory.
// The running max is only advanced when it grows by more than 2^8 (lazy rescaling): the O accumulator in TMEM is
// then rescaled by the owning softmax warpgroup; otherwise the hot loop never reads O.  That is safe without an
// extra barrier because the commit that publishes S_X(j) also covers PV_X(j-1) (tcgen05.commit tracks ALL prior
// MMAs of the issuing thread), and PV_X(j) is not issued before P_X(j) is published.
//
// Reference semantics: source/utils/gta.py:92-279 and source/layers.py:202-211.
#include <cmath>

#include "attn_common.cuh"

namespace gta {

constexpr int kThreads2 = 384;
constexpr uint32_t kTmemSA = 0, kTmemSB = 128, kTmemOA = 256, kTmemOB = 384;
constexpr float kRescaleThreshold = 8.0f;   // log2 units

template <int D>
struct Attn2Cfg {
    static constexpr int kStages = (D == 128) ? 2 : 3;
    static constexpr uint32_t kTile = 128u * D * 2u;
    static constexpr uint32_t kQ = 0;                          // [2] tiles
    static constexpr uint32_t kK = 2 * kTile;                  // [kStages]
    static constexpr uint32_t kV = kTile * (2 + kStages);      // [kStages]
    static constexpr uint32_t kBars = kTile * (2 + 2 * kStages);
    enum : int {
        bQFull = 0,                       // [2]  count 128
        bKFull = 2,                       // [kStages]
        bVFull = bKFull + kStages,
        bKEmpty = bVFull + kStages,
        bVEmpty = bKEmpty + kStages,
        bSFull = bVEmpty + kStages,       // [2]  tcgen05.commit
        bPFull = bSFull + 2,              // [2]  count 128
        bOFinal = bPFull + 2,             // [2]  tcgen05.commit, single phase
        bCount = bOFinal + 2
    };
    static constexpr uint32_t kTmemSlot = kBars + bCount * 8;
    static constexpr uint32_t kUsed = kTmemSlot + 16;
    static constexpr uint32_t kBytes = (kUsed + 1024 > 120u * 1024u) ? kUsed + 1024 : 120u * 1024u;
};

template <typename TIn, typename TOut, int D>
__global__ void __launch_bounds__(kThreads2, 1) attn_fwd2_kernel(const AttnArgs a) {
    using L = Attn2Cfg<D>;
    constexpr int NS = L::kStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBars);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kTmemSlot);

    const int qpair = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = a.ntiles_k;
    const bool has_b = (qpair * 256 + 128) < a.Tq;

    if (threadIdx.x == 0) {
        for (int x = 0; x < 2; ++x) {
            mbar_init(&bars[L::bQFull + x], 128);
            mbar_init(&bars[L::bSFull + x], 1);
            mbar_init(&bars[L::bPFull + x], 128);
            mbar_init(&bars[L::bOFinal + x], 1);
        }
        for (int s = 0; s < NS; ++s) {
            mbar_init(&bars[L::bKFull + s], 1);
            mbar_init(&bars[L::bVFull + s], 1);
            mbar_init(&bars[L::bKEmpty + s], 1);
            mbar_init(&bars[L::bVEmpty + s], 1);
        }
        fence_mbar_init();
    }
    if (warp == 8) {
        tmem_alloc(tmem_slot, kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

    if (warp < 8) {
        // =========================================================== softmax warpgroups
        setmaxnreg_inc<232>();
        const int X = warp >> 2;                       // 0 = tile A, 1 = tile B
        if (X == 0 || has_b) {                         // warpgroup-uniform
        const int r = threadIdx.x & 127;
        const int t = qpair * 256 + X * 128 + r;
        const bool valid = t < a.Tq;
        const int tt = valid ? t : a.Tq - 1;
        const float tc = a.tc_ptr ? __ldg(a.tc_ptr) : 1.0f;
        const size_t view = static_cast<size_t>(b) * a.Nq + tt / a.tpvq;
        const float* se3 = a.se3_q + view * 16;
        const float* so3 = a.so3_q + view * 34;
        const float* so2 = a.so2_q + (static_cast<size_t>(b) * a.Tq + tt) * a.C * 2;
        long long* dbg = nullptr;
        if (a.dbg && threadIdx.x == 0)
            dbg = a.dbg + ((static_cast<size_t>(b) * gridDim.y + h) * gridDim.x + qpair) * 8;
        long long wait_acc = 0;
        if (dbg) dbg[0] = clock64();

        {   // ---- Q prologue: raw strided row -> rho_q^{-T} in registers -> bf16 operand tile image.
            // All global loads of a group of chunks (data + reps) are issued before the first use.
            const TIn* qrow = reinterpret_cast<const TIn*>(a.q) + static_cast<int64_t>(b) * a.q_sb +
                              static_cast<int64_t>(h) * a.q_sh + static_cast<int64_t>(tt) * a.q_st;
            uint8_t* sQ = smem + L::kQ + X * L::kTile;
            constexpr int NC = D / 8;
            constexpr int G = (NC % 6 == 0) ? 6 : ((NC % 8 == 0) ? 8 : 4);
            ViewReps vr;
            load_view_reps(vr, a.hd, se3, so3);
#pragma unroll 1
            for (int g = 0; g < NC / G; ++g) {
                RawChunk<TIn> raw[G];
                So2Chunk sc[G];
#pragma unroll
                for (int i = 0; i < G; ++i) {
                    zero_raw(raw[i]);
                    if (valid) load_raw(qrow + (g * G + i) * 8, raw[i]);
                    sc[i] = load_so2_chunk(so2, g * G + i, a.hd);
                }
#pragma unroll
                for (int i = 0; i < G; ++i) {
                    float x[8];
                    raw_to_f32(raw[i], x);
                    apply_rep_chunk_pre<kModeQ>(x, g * G + i, a.hd, vr, sc[i], tc);
                    *reinterpret_cast<uint4*>(sQ + tile_sw64_offset(r, g * G + i)) = pack_chunk_bf16(x);
                }
            }
            fence_proxy_async_smem();
            mbar_arrive(&bars[L::bQFull + X]);
        }
        if (dbg) dbg[1] = clock64();

        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
        const uint32_t s_addr = lane_base + (X ? kTmemSB : kTmemSA);
        const uint32_t o_addr = lane_base + (X ? kTmemOB : kTmemOA);
        const float cs = a.scale_log2;
        const uint64_t cs2 = pack_f32x2(cs, cs);
        float m_used = -INFINITY, l_run = 0.f;

#pragma unroll 1
        for (int j = 0; j < n; ++j) {
            long long tw = 0;
            if (dbg) tw = clock64();
            mbar_wait(&bars[L::bSFull + X], j & 1);
            if (dbg) { const long long now = clock64(); if (j == 0) dbg[2] = now; else wait_acc += now - tw; }
            tc_fence_after();
            uint32_t sreg[128];
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
            float mx0 = fmax3(s[0], s[1], s[2]), mx1 = fmax3(s[3], s[4], s[5]);
            float mx2 = fmax3(s[6], s[7], s[8]), mx3 = fmax3(s[9], s[10], s[11]);
#pragma unroll
            for (int i = 12; i < 124; i += 8) {
                mx0 = fmax3(mx0, s[i], s[i + 1]); mx1 = fmax3(mx1, s[i + 2], s[i + 3]);
                mx2 = fmax3(mx2, s[i + 4], s[i + 5]); mx3 = fmax3(mx3, s[i + 6], s[i + 7]);
            }
            mx0 = fmax3(mx0, s[124], s[125]); mx1 = fmax3(mx1, s[126], s[127]);
            const float m_tile = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));

            // ---- lazy rescale: advance the reference max only when it grew by more than 2^kRescaleThreshold
            const bool grow = (m_tile - m_used) * cs > kRescaleThreshold;      // true on the first tile (m_used = -inf)
            if (__any_sync(0xffffffffu, grow)) {
                const float m_new = grow ? m_tile : m_used;
                const float alpha = grow ? fast_exp2((m_used - m_new) * cs) : 1.0f;
                l_run *= alpha;
                m_used = m_new;
                if (j > 0) {
#pragma unroll 1
                    for (int c8 = 0; c8 < D / 8; ++c8) {      // rare: keep the footprint at 8 registers
                        uint32_t o8[8];
                        tmem_ld8(o_addr + c8 * 8, o8);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 8; ++i) o8[i] = __float_as_uint(__uint_as_float(o8[i]) * alpha);
                        tmem_st8(o_addr + c8 * 8, o8);
                    }
                }
            }

            // ---- P = exp2(s*cs - m_used*cs), row sum in fp32, bf16 pairs back into S's first 64 columns
            const float neg = -m_used * cs;
            const uint64_t neg2 = pack_f32x2(neg, neg);
            uint64_t lsum2 = pack_f32x2(0.f, 0.f);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                // packed in place: P pair i overwrites sreg[half*64 + i] after s[half*64 + 2i], s[.. + 2i+1] were consumed,
                // so the store reuses the register block the load filled (no second 32-register block is needed)
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float x0, x1;
                    unpack_f32x2(ffma2(pack_f32x2(s[half * 64 + 2 * i], s[half * 64 + 2 * i + 1]), cs2, neg2), x0, x1);
                    const float p0 = fast_exp2(x0), p1 = fast_exp2(x1);
                    lsum2 = fadd2(lsum2, pack_f32x2(p0, p1));
                    sreg[half * 64 + i] = pack_bf16x2(p0, p1);
                }
                tmem_st32(s_addr + half * 32, sreg + half * 64);
            }
            float ls0, ls1;
            unpack_f32x2(lsum2, ls0, ls1);
            l_run += ls0 + ls1;
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&bars[L::bPFull + X]);
        }

        // ---- epilogue: O / l, rho_q^{-1} in registers, store [B,Tq,H,D].  The rep data is prefetched into registers
        // BEFORE waiting for the last PV so that its latency overlaps the tail of the MMA pipeline.
        if (dbg) dbg[3] = clock64();
        ViewReps vr;
        So2Chunk sc[D / 8];
        if (a.v_transform) {
            load_view_reps(vr, a.hd, se3, so3);
#pragma unroll
            for (int c = 0; c < D / 8; ++c) sc[c] = load_so2_chunk(so2, c, a.hd);
        }
        mbar_wait(&bars[L::bOFinal + X], 0);
        if (dbg) dbg[4] = clock64();
        tc_fence_after();
        const float inv_l = 1.0f / l_run;
        TOut* orow = reinterpret_cast<TOut*>(a.out) + ((static_cast<int64_t>(b) * a.Tq + tt) * a.H + h) * D;
#pragma unroll
        for (int cb = 0; cb < D / 32; ++cb) {
            uint32_t o[32];
            tmem_ld32(o_addr + cb * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                float x[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = __uint_as_float(o[cc * 8 + i]) * inv_l;
                const int c = cb * 4 + cc;
                if (a.v_transform) apply_rep_chunk_pre<kModeOut>(x, c, a.hd, vr, sc[c], tc);
                if (valid) store_chunk<TOut>(orow + c * 8, x);
            }
        }
        if (a.lse && valid)
            a.lse[(static_cast<int64_t>(b) * a.H + h) * a.Tq + t] = m_used * a.scale + logf(l_run);
        if (dbg) {
            dbg[5] = clock64();
            dbg[6] = wait_acc;
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            dbg[7] = smid;
        }
        tc_fence_before();
        }
    } else {
        setmaxnreg_dec<40>();
        if (warp == 8) {
            // ======================================================= UMMA issuer
            constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);
            constexpr uint32_t idesc_pv = make_idesc_bf16(128, D, 0, 1);
            const uint32_t q_addr = smem_u32(smem + L::kQ);
            const int nx = has_b ? 2 : 1;

            auto issue_qk = [&](int X, int j) {       // S_X(j) = Q_X K'(j)^T ; commit publishes it (and PV_X(j-1))
                const int s = j % NS;
                if (lane == 0) {
                    const uint32_t k_addr = smem_u32(smem + L::kK + s * L::kTile);
                    const uint32_t d_addr = tmem_base + (X ? kTmemSB : kTmemSA);
#pragma unroll
                    for (int kk = 0; kk < D / 16; ++kk)
                        umma_ss(d_addr, desc_kmajor_sw64(q_addr + X * L::kTile, kk), desc_kmajor_sw64(k_addr, kk),
                                idesc_qk, kk > 0);
                    if (X == nx - 1) umma_commit(&bars[L::bKEmpty + s]);
                    umma_commit(&bars[L::bSFull + X]);
                }
                __syncwarp();
            };
            auto issue_pv = [&](int X, int j) {       // O_X += P_X(j) V'(j)
                const int s = j % NS;
                mbar_wait(&bars[L::bPFull + X], j & 1);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t v_addr = smem_u32(smem + L::kV + s * L::kTile);
                    const uint32_t d_addr = tmem_base + (X ? kTmemOB : kTmemOA);
                    const uint32_t p_addr = tmem_base + (X ? kTmemSB : kTmemSA);
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk)
                        umma_ts(d_addr, p_addr + kk * 8, desc_mnmajor_sw64(v_addr, kk), idesc_pv,
                                (j > 0 || kk > 0) ? 1u : 0u);
                    if (X == nx - 1) umma_commit(&bars[L::bVEmpty + s]);
                    if (j == n - 1) umma_commit(&bars[L::bOFinal + X]);
                }
                __syncwarp();
            };

            mbar_wait(&bars[L::bKFull + 0], 0);
            for (int X = 0; X < nx; ++X) {
                mbar_wait(&bars[L::bQFull + X], 0);
                tc_fence_after();
                issue_qk(X, 0);
            }
#pragma unroll 1
            for (int j = 0; j < n; ++j) {
                const int s = j % NS;
                mbar_wait(&bars[L::bVFull + s], (j / NS) & 1);
                if (j + 1 < n) mbar_wait(&bars[L::bKFull + (j + 1) % NS], ((j + 1) / NS) & 1);
                for (int X = 0; X < nx; ++X) {
                    issue_pv(X, j);
                    if (j + 1 < n) issue_qk(X, j + 1);
                }
            }
        } else if (warp == 9) {
            // ======================================================= bulk-copy producer
            const size_t blob0 = (static_cast<size_t>(b) * a.H + h) * n;
#pragma unroll 1
            for (int j = 0; j < n; ++j) {
                const int s = j % NS;
                if (j >= NS) mbar_wait(&bars[L::bKEmpty + s], ((j / NS) - 1) & 1);
                if (lane == 0) {
                    mbar_arrive_expect_tx(&bars[L::bKFull + s], L::kTile);
                    bulk_g2s(smem + L::kK + s * L::kTile, a.ws_k + (blob0 + j) * L::kTile, L::kTile, &bars[L::bKFull + s]);
                }
                if (j >= NS) mbar_wait(&bars[L::bVEmpty + s], ((j / NS) - 1) & 1);
                if (lane == 0) {
                    mbar_arrive_expect_tx(&bars[L::bVFull + s], L::kTile);
                    bulk_g2s(smem + L::kV + s * L::kTile, a.ws_v + (blob0 + j) * L::kTile, L::kTile, &bars[L::bVFull + s]);
                }
                __syncwarp();
            }
        }
    }
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

template <typename TIn, typename TOut, int D>
static int launch2_one(const AttnArgs& a, dim3 grid, cudaStream_t st) {
    using L = Attn2Cfg<D>;
    auto kern = attn_fwd2_kernel<TIn, TOut, D>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(L::kBytes));
    if (e != cudaSuccess) return set_error(GTA_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    kern<<<grid, kThreads2, L::kBytes, st>>>(a);
    return check_launch("gta_attn_fwd");
}

template <typename TIn, typename TOut>
static int launch2_d(const AttnArgs& a, int D, dim3 grid, cudaStream_t st) {
    switch (D) {
        case 32: return launch2_one<TIn, TOut, 32>(a, grid, st);
        case 64: return launch2_one<TIn, TOut, 64>(a, grid, st);
        case 96: return launch2_one<TIn, TOut, 96>(a, grid, st);
        case 128: return launch2_one<TIn, TOut, 128>(a, grid, st);
    }
    return set_error(GTA_ERR_UNSUPPORTED, "gta_attn_fwd: head dim %d not in {32,64,96,128}", D);
}

int launch_attn_fwd(const GtaAttnParams& p, cudaStream_t st) {
    if (!(p.flags & GTA_FLAG_V1_PIPELINE) && p.D <= 96) {
        if (p.flags & GTA_FLAG_V5_PIPELINE) return launch_attn_fwd_v5(p, st);
        if (p.flags & GTA_FLAG_V4_PIPELINE) {
            bool handled = false;
            const int rc = launch_attn_fwd_v4(p, st, &handled);
            if (handled) return rc;
        }
        return launch_attn_fwd_v2(p, st);  // persistent pipeline
    }
    const AttnArgs a = make_attn_args(p);
    dim3 grid((p.Tq + 255) / 256, p.H, p.B);
    const bool ib = p.in_dtype == GTA_DTYPE_BF16, ob = p.out_dtype == GTA_DTYPE_BF16;
    if (ib && ob) return launch2_d<__nv_bfloat16, __nv_bfloat16>(a, p.D, grid, st);
    if (ib && !ob) return launch2_d<__nv_bfloat16, float>(a, p.D, grid, st);
    if (!ib && ob) return launch2_d<float, __nv_bfloat16>(a, p.D, grid, st);
    return launch2_d<float, float>(a, p.D, grid, st);
}

}  // namespace gta
