// Fused GTA attention forward, v3 pipeline (default): persistent CTAs, two 128-query tiles per work item, and the
// score tile split into 64-key HALF tiles with a DOUBLE-BUFFERED S accumulator per query tile, so that
// S_X(jj+1) is already in tensor memory when the softmax warpgroup finishes half tile jj (the v2 profile showed the
// softmax warpgroups and the UMMA issuer waiting on each other ~45 % of the time with one S buffer per query tile).
//
// Work item = (batch b, head h, query pair p): 256 query rows of one (b,h) against all Tk keys.
// CTA c processes items c, c+grid, ...; K'/V' tile images stay 128 keys (one bulk copy each) and are consumed as
// two 64-key halves.
//
//   warps 0-3   softmax warpgroup A   thread i <-> query row i of tile A <-> TMEM lane i
//   warps 4-7   softmax warpgroup B
//   warp  8     UMMA issuer (one lane):  per half tile  PV_A(jj) QK_A(jj+2) PV_B(jj) QK_B(jj+2)
//   warp  9     bulk-copy producer (2-stage K'/V' ring, runs ahead across item boundaries)
//   warps 10-11 Q stager for the NEXT item (rho_q^{-T} in fp32 registers -> bf16 operand tiles, double-buffered)
//
// TMEM (512 columns): S_A[2] 0..127, S_B[2] 128..255 (64 fp32 columns per buffer), O_A 256..351, O_B 384..479.
// P (bf16, 64 keys = 32 columns) overwrites the first half of the S buffer it was computed from and is the A operand
// of the PV product (TS form).  The reference max only advances when a half-tile max exceeds it by > 2^8.
//
// Reference semantics: source/utils/gta.py:92-279 and source/layers.py:202-211.
#include <cmath>

#include "attn_common.cuh"

namespace gta {

constexpr int kThreads4 = 384;
constexpr int k4StagerThreads = 64;
constexpr uint32_t k4TmemS = 0;       // + X*128 + buf*64
constexpr uint32_t k4TmemO = 256;     // + X*128
constexpr float k4RescaleThreshold = 8.0f;
#ifndef GTA_POLY_NUM
#define GTA_POLY_NUM 0
#endif
#ifndef GTA_POLY_DEN
#define GTA_POLY_DEN 4
#endif

template <int D>
struct Attn4Cfg {
    static constexpr int kStages = 2;
    static constexpr uint32_t kTile = 128u * D * 2u;
    static constexpr uint32_t kQ = 0;                          // [2 buffers][2 tiles]
    static constexpr uint32_t kK = 4 * kTile;                  // [kStages]
    static constexpr uint32_t kV = kTile * (4 + kStages);      // [kStages]
    static constexpr uint32_t kBars = kTile * (4 + 2 * kStages);
    enum : int {
        bQFull = 0,                        // [qbuf][X]  count 64 (stager threads)
        bQFree = 4,                        // [qbuf][X]  commit after the item's last QK_X
        bKFull = 8,                        // [kStages]
        bVFull = bKFull + kStages,
        bKEmpty = bVFull + kStages,
        bVEmpty = bKEmpty + kStages,
        bSFull = bVEmpty + kStages,        // [X][sbuf] commit
        bPFull = bSFull + 4,               // [X][sbuf] count 128
        bOFinal = bPFull + 4,              // [X] commit after the item's last PV_X
        bOFree = bOFinal + 2,              // [X] count 128: O_X read out
        bPVDone = bOFree + 2,              // [X] commit after EVERY PV_X (only waited on by the rare rescale path)
        bCount = bPVDone + 2
    };
    static constexpr uint32_t kTmemSlot = kBars + bCount * 8;
    static constexpr uint32_t kUsed = kTmemSlot + 16;
    static constexpr uint32_t kBytes = (kUsed + 1024 > 120u * 1024u) ? kUsed + 1024 : 120u * 1024u;
};

struct Item4 {
    int b, h, p;
    bool has_b;
};
__device__ __forceinline__ Item4 decode_item4(int item, int npairs, int H, int Tq) {
    Item4 c;
    c.p = item % npairs;
    const int bh = item / npairs;
    c.h = bh % H;
    c.b = bh / H;
    c.has_b = (c.p * 256 + 128) < Tq;
    return c;
}

template <typename TIn, typename TOut, int D>
__global__ void __launch_bounds__(kThreads4, 1) attn_fwd4_kernel(const AttnArgs a, const int npairs, const int nitems) {
    using L = Attn4Cfg<D>;
    constexpr int NS = L::kStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBars);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kTmemSlot);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = a.ntiles_k;                 // 128-key tile images
    const int nh = (a.Tk + 63) >> 6;          // 64-key half tiles actually holding keys

    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) {
            mbar_init(&bars[L::bQFull + i], k4StagerThreads);
            mbar_init(&bars[L::bQFree + i], 1);
            mbar_init(&bars[L::bSFull + i], 1);
            mbar_init(&bars[L::bPFull + i], 128);
        }
        for (int x = 0; x < 2; ++x) {
            mbar_init(&bars[L::bOFinal + x], 1);
            mbar_init(&bars[L::bOFree + x], 128);
            mbar_init(&bars[L::bPVDone + x], 1);
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
    const float tc = a.tc_ptr ? __ldg(a.tc_ptr) : 1.0f;

    if (warp < 8) {
        // =========================================================== softmax warpgroups (+ epilogue)
        setmaxnreg_inc<200>();
        const int X = warp >> 2;
        const int r = threadIdx.x & 127;
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
        const uint32_t s_base = lane_base + k4TmemS + X * 128;
        const uint32_t o_addr = lane_base + k4TmemO + X * 128;
        const float cs = a.scale_log2;
        const uint64_t cs2 = pack_f32x2(cs, cs);
        uint32_t gh = 0;      // half tiles processed by this warpgroup: buffer = gh & 1, phase = (gh >> 1) & 1
        uint32_t cnt = 0;     // items processed by this warpgroup (o_final phase)
        long long* dbg = (a.dbg && threadIdx.x == 0) ? a.dbg + static_cast<size_t>(blockIdx.x) * 16 : nullptr;
        long long d_loop = 0, d_epi = 0, d_wait_s = 0, d_wait_o = 0, d_items = 0;
        const long long d_start = dbg ? clock64() : 0;

#pragma unroll 1
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const Item4 ic = decode_item4(item, npairs, a.H, a.Tq);
            if (X == 1 && !ic.has_b) continue;
            float m_used = -INFINITY, l_run = 0.f;
            const long long d_t0 = dbg ? clock64() : 0;

#pragma unroll 1
            for (int jj = 0; jj < nh; ++jj, ++gh) {
                const uint32_t sb = gh & 1;
                const uint32_t s_addr = s_base + sb * 64;
                const long long d_w0 = dbg ? clock64() : 0;
                mbar_wait(&bars[L::bSFull + X * 2 + sb], (gh >> 1) & 1);
                if (dbg) d_wait_s += clock64() - d_w0;
                tc_fence_after();
                uint32_t sreg[64];
                tmem_ld32(s_addr, sreg);
                tmem_ld32(s_addr + 32, sreg + 32);
                tmem_ld_wait();
                float* s = reinterpret_cast<float*>(sreg);
                if (jj == nh - 1) {
                    const int nvalid = a.Tk - jj * 64;
                    if (nvalid < 64) {
#pragma unroll
                        for (int i = 0; i < 64; ++i) if (i >= nvalid) s[i] = -INFINITY;
                    }
                }
                float mx0 = fmax3(s[0], s[1], s[2]), mx1 = fmax3(s[3], s[4], s[5]);
                float mx2 = fmax3(s[6], s[7], s[8]), mx3 = fmax3(s[9], s[10], s[11]);
#pragma unroll
                for (int i = 12; i < 60; i += 8) {
                    mx0 = fmax3(mx0, s[i], s[i + 1]); mx1 = fmax3(mx1, s[i + 2], s[i + 3]);
                    mx2 = fmax3(mx2, s[i + 4], s[i + 5]); mx3 = fmax3(mx3, s[i + 6], s[i + 7]);
                }
                mx0 = fmax3(mx0, s[60], s[61]); mx1 = fmax3(mx1, s[62], s[63]);
                const float m_tile = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));

                const bool grow = (m_tile - m_used) * cs > k4RescaleThreshold;   // always true on the item's first half tile
                if (__any_sync(0xffffffffu, grow)) {
                    const float m_new = grow ? m_tile : m_used;
                    const float alpha = grow ? fast_exp2((m_used - m_new) * cs) : 1.0f;
                    l_run *= alpha;
                    m_used = m_new;
                    if (jj > 0) {
                        // With two S buffers PV_X(jj-1) may still be accumulating into O_X: wait for its commit.
                        // (PV_X(jj-2) is known complete — the commit that published S_X(jj) covers it — and PV_X(jj)
                        // cannot be issued before this warpgroup publishes P_X(jj), so the barrier is exactly at
                        // phase gh-1 or gh and the parity wait is unambiguous.)
                        mbar_wait(&bars[L::bPVDone + X], (gh - 1) & 1);
                        tc_fence_after();
#pragma unroll 1
                        for (int c8 = 0; c8 < D / 8; ++c8) {
                            uint32_t o8[8];
                            tmem_ld8(o_addr + c8 * 8, o8);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 8; ++i) o8[i] = __float_as_uint(__uint_as_float(o8[i]) * alpha);
                            tmem_st8(o_addr + c8 * 8, o8);
                        }
                    }
                }

                const float neg = -m_used * cs;
                const uint64_t neg2 = pack_f32x2(neg, neg);
                uint64_t lsum2 = pack_f32x2(0.f, 0.f);
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const uint64_t x2 = ffma2(pack_f32x2(s[2 * i], s[2 * i + 1]), cs2, neg2);
                    float p0, p1;
                    if ((i % GTA_POLY_DEN) < GTA_POLY_NUM) {
                        poly_exp2x2(x2, p0, p1);
                    } else {
                        float x0, x1;
                        unpack_f32x2(x2, x0, x1);
                        p0 = fast_exp2(x0); p1 = fast_exp2(x1);
                    }
                    lsum2 = fadd2(lsum2, pack_f32x2(p0, p1));
                    sreg[i] = pack_bf16x2(p0, p1);           // in place: pair i is consumed before slot i is reused
                }
                tmem_st32(s_addr, sreg);
                float ls0, ls1;
                unpack_f32x2(lsum2, ls0, ls1);
                l_run += ls0 + ls1;
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(&bars[L::bPFull + X * 2 + sb]);
            }

            // ---- epilogue of this item.  Everything it needs from global memory is requested BEFORE waiting for the
            // last PV: the view matrices and the row's SO(2) table.
            const long long d_t1 = dbg ? clock64() : 0;
            const int t = ic.p * 256 + X * 128 + r;
            const bool valid = t < a.Tq;
            const int tt = valid ? t : a.Tq - 1;
            ViewReps vr;
            So2Chunk sc[D / 8];
            if (a.v_transform) {
                const size_t view = static_cast<size_t>(ic.b) * a.Nq + tt / a.tpvq;
                load_view_reps(vr, a.hd, a.se3_q + view * 16, a.so3_q + view * 34);
                const float* so2 = a.so2_q + (static_cast<size_t>(ic.b) * a.Tq + tt) * a.C * 2;
#pragma unroll
                for (int c = 0; c < D / 8; ++c) sc[c] = load_so2_chunk(so2, c, a.hd);
            }
            mbar_wait(&bars[L::bOFinal + X], cnt & 1);
            const long long d_t2 = dbg ? clock64() : 0;
            ++cnt;
            tc_fence_after();
            const float inv_l = 1.0f / l_run;
            TOut* orow = reinterpret_cast<TOut*>(a.out) + ((static_cast<int64_t>(ic.b) * a.Tq + tt) * a.H + ic.h) * D;
#pragma unroll
            for (int cb = 0; cb < D / 32; ++cb) {
                uint32_t o[32];
                tmem_ld32(o_addr + cb * 32, o);
                tmem_ld_wait();
                if (cb == D / 32 - 1) {                       // O_X fully read: the next item's PV_X(0) may overwrite it
                    tc_fence_before();
                    mbar_arrive(&bars[L::bOFree + X]);
                }
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    const int c = cb * 4 + cc;
                    float x[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) x[i] = __uint_as_float(o[cc * 8 + i]) * inv_l;
                    if (a.v_transform) apply_rep_chunk_pre<kModeOut>(x, c, a.hd, vr, sc[c], tc);
                    if (valid) store_chunk<TOut>(orow + c * 8, x);
                }
            }
            if (a.lse && valid)
                a.lse[(static_cast<int64_t>(ic.b) * a.H + ic.h) * a.Tq + t] = m_used * a.scale + logf(l_run);
            if (dbg) {
                const long long d_t3 = clock64();
                d_loop += d_t1 - d_t0; d_wait_o += d_t2 - d_t1; d_epi += d_t3 - d_t2; ++d_items;
            }
        }
        if (dbg) {
            dbg[0] = clock64() - d_start; dbg[1] = d_loop; dbg[2] = d_epi; dbg[3] = d_wait_s; dbg[4] = d_wait_o;
            dbg[5] = d_items;
        }
    } else {
      setmaxnreg_dec<96>();
      if (warp >= 10) {
        // =========================================================== Q stager (runs one item ahead)
        const int r0 = threadIdx.x - 320;    // 0..63; this thread stages rows r0 and r0 + 64 of each tile
        uint32_t cntx[2] = {0, 0};          // items staged per tile slot (buffer = cnt & 1, phase = (cnt >> 1) & 1)
#pragma unroll 1
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const Item4 ic = decode_item4(item, npairs, a.H, a.Tq);
#pragma unroll 1
            for (int X = 0; X < 2; ++X) {
                if (X == 1 && !ic.has_b) continue;
                const uint32_t c_ = cntx[X]++;
                const int buf = c_ & 1;
                if (c_ >= 2) mbar_wait(&bars[L::bQFree + buf * 2 + X], ((c_ >> 1) - 1) & 1);
                uint8_t* sQ = smem + L::kQ + (buf * 2 + X) * L::kTile;
#pragma unroll 1
                for (int rr = 0; rr < 2; ++rr) {
                    const int r = r0 + rr * 64;
                    const int t = ic.p * 256 + X * 128 + r;
                    const bool valid = t < a.Tq;
                    const int tt = valid ? t : a.Tq - 1;
                    const size_t view = static_cast<size_t>(ic.b) * a.Nq + tt / a.tpvq;
                    const float* so2 = a.so2_q + (static_cast<size_t>(ic.b) * a.Tq + tt) * a.C * 2;
                    const TIn* qrow = reinterpret_cast<const TIn*>(a.q) + static_cast<int64_t>(ic.b) * a.q_sb +
                                      static_cast<int64_t>(ic.h) * a.q_sh + static_cast<int64_t>(tt) * a.q_st;
                    constexpr int NC = D / 8;
                    constexpr int G = (sizeof(TIn) == 2) ? NC : ((NC % 6 == 0) ? 6 : 4);   // <= 48 registers of raw data
                    const float* se3 = a.se3_q + view * 16;
                    const float* so3 = a.so3_q + view * 34;
#pragma unroll 1
                    for (int g = 0; g < NC / G; ++g) {
                        RawChunk<TIn> raw[G];
#pragma unroll
                        for (int i = 0; i < G; ++i) {
                            zero_raw(raw[i]);
                            if (valid) load_raw(qrow + (g * G + i) * 8, raw[i]);
                        }
#pragma unroll
                        for (int i = 0; i < G; ++i) {
                            float x[8];
                            raw_to_f32(raw[i], x);
                            apply_rep_chunk<kModeQ>(x, g * G + i, a.hd, se3, so3, so2, tc);
                            *reinterpret_cast<uint4*>(sQ + tile_sw64_offset(r, g * G + i)) = pack_chunk_bf16(x);
                        }
                    }
                }
                fence_proxy_async_smem();
                mbar_arrive(&bars[L::bQFull + buf * 2 + X]);
            }
        }
      } else if (warp == 8) {
            // ======================================================= UMMA issuer
            constexpr uint32_t idesc_qk = make_idesc_bf16(128, 64, 0, 0);
            constexpr uint32_t idesc_pv = make_idesc_bf16(128, D, 0, 1);
            uint32_t gk = 0;                   // 128-key tile images consumed by this CTA (K/V ring position)
            uint32_t ghx[2] = {0, 0};          // half tiles per query tile slot (S buffer / s_full / p_full phases)
            uint32_t cntx[2] = {0, 0};         // items per query tile slot (Q buffer / o_free phase)
            long long* dbg = (a.dbg && lane == 0) ? a.dbg + static_cast<size_t>(blockIdx.x) * 16 : nullptr;
            long long w_k = 0, w_v = 0, w_p = 0, w_of = 0, w_q = 0;
#define GTA_TIMED_WAIT(acc, ...)                                 \
    do {                                                         \
        const long long t0_ = dbg ? clock64() : 0;               \
        __VA_ARGS__;                                             \
        if (dbg) acc += clock64() - t0_;                         \
    } while (0)

#pragma unroll 1
            for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
                const Item4 ic = decode_item4(item, npairs, a.H, a.Tq);
                const int nx = ic.has_b ? 2 : 1;
                uint32_t q_addr[2];
                for (int X = 0; X < nx; ++X)
                    q_addr[X] = smem_u32(smem + L::kQ + ((cntx[X] & 1) * 2 + X) * L::kTile);

                // S_X(jj) = Q_X K'(jj)^T for half tile jj (keys 64*jj ..) into S buffer (ghx[X] + jj) & 1
                auto issue_qk = [&](int X, int jj) {
                    const int s = (gk + (jj >> 1)) % NS;
                    const uint32_t sb = (ghx[X] + jj) & 1;
                    if (lane == 0) {
                        const uint32_t k_addr = smem_u32(smem + L::kK + s * L::kTile) + (jj & 1) * 4096u;
                        const uint32_t d_addr = tmem_base + k4TmemS + X * 128 + sb * 64;
#pragma unroll
                        for (int kk = 0; kk < D / 16; ++kk)
                            umma_ss(d_addr, desc_kmajor_sw64(q_addr[X], kk), desc_kmajor_sw64(k_addr, kk), idesc_qk, kk > 0);
                        // last reader of this K' tile image: second half (or the item's last half tile) of the last slot
                        if (X == nx - 1 && ((jj & 1) || jj == nh - 1)) umma_commit(&bars[L::bKEmpty + s]);
                        if (jj == nh - 1) umma_commit(&bars[L::bQFree + (cntx[X] & 1) * 2 + X]);
                        umma_commit(&bars[L::bSFull + X * 2 + sb]);
                    }
                    __syncwarp();
                };
                // O_X += P_X(jj) V'(jj)
                auto issue_pv = [&](int X, int jj) {
                    const int s = (gk + (jj >> 1)) % NS;
                    const uint32_t gh_ = ghx[X] + jj;
                    const uint32_t sb = gh_ & 1;
                    GTA_TIMED_WAIT(w_p, mbar_wait(&bars[L::bPFull + X * 2 + sb], (gh_ >> 1) & 1));
                    if (jj == 0 && cntx[X] > 0) GTA_TIMED_WAIT(w_of, mbar_wait(&bars[L::bOFree + X], (cntx[X] - 1) & 1));
                    tc_fence_after();
                    if (lane == 0) {
                        const uint32_t v_addr = smem_u32(smem + L::kV + s * L::kTile);
                        const uint32_t d_addr = tmem_base + k4TmemO + X * 128;
                        const uint32_t p_addr = tmem_base + k4TmemS + X * 128 + sb * 64;
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            umma_ts(d_addr, p_addr + kk * 8, desc_mnmajor_sw64(v_addr, (jj & 1) * 4 + kk), idesc_pv,
                                    (jj > 0 || kk > 0) ? 1u : 0u);
                        if (X == nx - 1 && ((jj & 1) || jj == nh - 1)) umma_commit(&bars[L::bVEmpty + s]);
                        if (jj == nh - 1) umma_commit(&bars[L::bOFinal + X]);
                        umma_commit(&bars[L::bPVDone + X]);
                    }
                    __syncwarp();
                };

                // prologue of the item: fill both S buffers of every query tile
                GTA_TIMED_WAIT(w_k, mbar_wait(&bars[L::bKFull + gk % NS], (gk / NS) & 1));
                for (int X = 0; X < nx; ++X) {
                    const uint32_t c_ = cntx[X];
                    GTA_TIMED_WAIT(w_q, mbar_wait(&bars[L::bQFull + (c_ & 1) * 2 + X], (c_ >> 1) & 1));
                    tc_fence_after();
                    issue_qk(X, 0);
                }
                if (nh > 1)
                    for (int X = 0; X < nx; ++X) issue_qk(X, 1);
#pragma unroll 1
                for (int jj = 0; jj < nh; ++jj) {
                    if ((jj & 1) == 0)
                        GTA_TIMED_WAIT(w_v, mbar_wait(&bars[L::bVFull + (gk + (jj >> 1)) % NS], ((gk + (jj >> 1)) / NS) & 1));
                    if (jj + 2 < nh && (jj & 1) == 0) {
                        const uint32_t g2 = gk + ((jj + 2) >> 1);
                        GTA_TIMED_WAIT(w_k, mbar_wait(&bars[L::bKFull + g2 % NS], (g2 / NS) & 1));
                    }
                    for (int X = 0; X < nx; ++X) {
                        issue_pv(X, jj);
                        if (jj + 2 < nh) issue_qk(X, jj + 2);
                    }
                }
                gk += n;
                for (int X = 0; X < nx; ++X) { ghx[X] += nh; ++cntx[X]; }
            }
            if (dbg) { dbg[8] = w_k; dbg[9] = w_v; dbg[10] = w_p; dbg[11] = w_of; dbg[12] = w_q; }
#undef GTA_TIMED_WAIT
      } else if (warp == 9) {
            // ======================================================= bulk-copy producer
            uint32_t gk = 0;
#pragma unroll 1
            for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
                const Item4 ic = decode_item4(item, npairs, a.H, a.Tq);
                const size_t blob0 = (static_cast<size_t>(ic.b) * a.H + ic.h) * n;
#pragma unroll 1
                for (int j = 0; j < n; ++j, ++gk) {
                    const int s = gk % NS;
                    if (gk >= NS) mbar_wait(&bars[L::bKEmpty + s], ((gk / NS) - 1) & 1);
                    if (lane == 0) {
                        mbar_arrive_expect_tx(&bars[L::bKFull + s], L::kTile);
                        bulk_g2s(smem + L::kK + s * L::kTile, a.ws_k + (blob0 + j) * L::kTile, L::kTile, &bars[L::bKFull + s]);
                    }
                    if (gk >= NS) mbar_wait(&bars[L::bVEmpty + s], ((gk / NS) - 1) & 1);
                    if (lane == 0) {
                        mbar_arrive_expect_tx(&bars[L::bVFull + s], L::kTile);
                        bulk_g2s(smem + L::kV + s * L::kTile, a.ws_v + (blob0 + j) * L::kTile, L::kTile, &bars[L::bVFull + s]);
                    }
                    __syncwarp();
                }
            }
      }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

template <typename TIn, typename TOut, int D>
static int launch4_one(const AttnArgs& a, const GtaAttnParams& p, cudaStream_t st) {
    using L = Attn4Cfg<D>;
    auto kern = attn_fwd4_kernel<TIn, TOut, D>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(L::kBytes));
    if (e != cudaSuccess) return set_error(GTA_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (num_sms <= 0) num_sms = 148;
    }
    const int npairs = (p.Tq + 255) / 256;
    const long long nitems = static_cast<long long>(p.B) * p.H * npairs;
    if (nitems > 0x7fffffffLL) return set_error(GTA_ERR_UNSUPPORTED, "too many work items");
    const int grid = static_cast<int>(nitems < num_sms ? nitems : num_sms);
    kern<<<grid, kThreads4, L::kBytes, st>>>(a, npairs, static_cast<int>(nitems));
    return check_launch("gta_attn_fwd");
}

template <typename TIn, typename TOut>
static int launch4_d(const AttnArgs& a, const GtaAttnParams& p, cudaStream_t st) {
    switch (p.D) {
        case 32: return launch4_one<TIn, TOut, 32>(a, p, st);
        case 64: return launch4_one<TIn, TOut, 64>(a, p, st);
        case 96: return launch4_one<TIn, TOut, 96>(a, p, st);
    }
    return set_error(GTA_ERR_UNSUPPORTED, "persistent pipeline supports head dims 32/64/96");
}

int launch_attn_fwd_v3(const GtaAttnParams& p, cudaStream_t st) {
    const AttnArgs a = make_attn_args(p);
    const bool ib = p.in_dtype == GTA_DTYPE_BF16, ob = p.out_dtype == GTA_DTYPE_BF16;
    if (ib && ob) return launch4_d<__nv_bfloat16, __nv_bfloat16>(a, p, st);
    if (ib && !ob) return launch4_d<__nv_bfloat16, float>(a, p, st);
    if (!ib && ob) return launch4_d<float, __nv_bfloat16>(a, p, st);
    return launch4_d<float, float>(a, p, st);
}

}  // namespace gta
