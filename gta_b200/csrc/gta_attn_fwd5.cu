// Fused GTA attention forward, v4 pipeline (default for head dims <= 96): persistent CTAs, two 128-query tiles per
// work item, and S / P DECOUPLED so that the next score tile is computed WHILE the softmax of the current one runs.
//
// v2 measurements (tools/phase_timing2.py): P aliased the S accumulator, so QK_X(j+1) could only be issued after
// PV_X(j); every softmax warpgroup then waited PV+QK (+ the other tile's MMAs queued in front) ~1200 clk per key tile
// and the UMMA issuer waited ~28 % of the time for P.  Here:
//   * the softmax warpgroup releases S_X as soon as it has been copied to registers (s_free) and QK_X(j+1) is issued
//     right then, overlapping the exp2 phase;
//   * P no longer aliases S: ONE P buffer in the last 64 TMEM columns is shared by both query tiles in strict
//     alternation (their softmax phases are offset by half a cycle anyway) — TMEM is exactly full:
//     S_A S_B 256 + O_A O_B 192 + P 64 = 512 columns; both PV products use the TS form;
//   * each query tile has its OWN UMMA issuer warp running the stream [s_free(j) -> QK(j+1)] [p_full(j) -> PV(j)]
//     with blocking waits on its own barriers, so a slow tile never blocks the other (a single polling issuer was
//     tried first: mbarrier.test_wait costs ~150 clk and the serial event loop became the bottleneck);
//   * K'/V' stages are released by BOTH streams (k_empty/v_empty count 2), PV completion is published per tile
//     (pv_done) for the lazy accumulator rescale and (p_free) for the shared P buffer.
// The launcher only selects this kernel when every work item has both query tiles (ceil(Tq/128) even); other shapes
// run the v2 pipeline.
//
//   warps 0-3 / 4-7  softmax warpgroups A / B (thread i <-> query row i <-> TMEM lane i), also run the epilogue
//   warps 8 / 9      UMMA issuers of tile A / B      warp 10   bulk-copy producer (2-stage K'/V' ring)
//   warp  11         Q stager for the NEXT item (rho_q^{-T} in fp32 registers, double-buffered operand tiles)
//
// Reference semantics: source/utils/gta.py:92-279 and source/layers.py:202-211.
#include <cmath>

#include "attn_common.cuh"

namespace gta {

constexpr int kThreads5 = 384;
constexpr int k5StagerThreads = 32;
constexpr uint32_t k5TmemS = 0;       // + X*128
constexpr uint32_t k5TmemO = 256;     // + X*96
constexpr uint32_t k5TmemPA = 448;    // P of tile A (64 columns)
constexpr float k5RescaleThreshold = 8.0f;
#ifndef GTA_POLY_NUM
#define GTA_POLY_NUM 0
#endif
#ifndef GTA_POLY_DEN
#define GTA_POLY_DEN 4
#endif

template <int D>
struct Attn5Cfg {
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
        bSFull = bVEmpty + kStages,        // [X] commit after QK_X
        bSFree = bSFull + 2,               // [X] count 128: S_X copied to registers
        bPFull = bSFree + 2,               // [X] count 128: P_X written
        bPVDone = bPFull + 2,              // [X] commit after every PV_X
        bOFinal = bPVDone + 2,             // [X] commit after the item's last PV_X
        bOFree = bOFinal + 2,              // [X] count 128: O_X read out
        bPFree = bOFree + 2,               // commit after every PV (either stream): the shared P buffer may be rewritten
        bCount = bPFree + 1
    };
    static constexpr uint32_t kTmemSlot = kBars + bCount * 8;
    static constexpr uint32_t kUsed = kTmemSlot + 16;
    static constexpr uint32_t kBytes = (kUsed + 1024 > 120u * 1024u) ? kUsed + 1024 : 120u * 1024u;
};

struct Item5 {
    int b, h, p;
    bool has_b;
};
__device__ __forceinline__ Item5 decode_item5(int item, int npairs, int H, int Tq) {
    Item5 c;
    c.p = item % npairs;
    const int bh = item / npairs;
    c.h = bh % H;
    c.b = bh / H;
    c.has_b = (c.p * 256 + 128) < Tq;
    return c;
}

template <typename TIn, typename TOut, int D>
__global__ void __launch_bounds__(kThreads5, 1) attn_fwd5_kernel(const AttnArgs a, const int npairs, const int nitems) {
    using L = Attn5Cfg<D>;
    constexpr int NS = L::kStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBars);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kTmemSlot);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = a.ntiles_k;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) {
            mbar_init(&bars[L::bQFull + i], k5StagerThreads);
            mbar_init(&bars[L::bQFree + i], 1);
        }
        for (int x = 0; x < 2; ++x) {
            mbar_init(&bars[L::bSFull + x], 1);
            mbar_init(&bars[L::bSFree + x], 128);
            mbar_init(&bars[L::bPFull + x], 128);
            mbar_init(&bars[L::bPVDone + x], 1);
            mbar_init(&bars[L::bOFinal + x], 1);
            mbar_init(&bars[L::bOFree + x], 128);
        }
        mbar_init(&bars[L::bPFree], 1);
        for (int s = 0; s < NS; ++s) {
            mbar_init(&bars[L::bKFull + s], 1);
            mbar_init(&bars[L::bVFull + s], 1);
            mbar_init(&bars[L::bKEmpty + s], 2);     // released by BOTH UMMA streams
            mbar_init(&bars[L::bVEmpty + s], 2);
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
        const uint32_t s_addr = lane_base + k5TmemS + X * 128;
        const uint32_t o_addr = lane_base + k5TmemO + X * 96;
        const uint32_t pa_addr = lane_base + k5TmemPA;
        uint32_t T0 = 0;       // CTA-global index of the current item's first key tile (items without a B tile count too)
        const float cs = a.scale_log2;
        const uint64_t cs2 = pack_f32x2(cs, cs);
        uint32_t gt = 0;      // key tiles processed by this warpgroup (s_full / s_free / p_full / pv_done phases)
        uint32_t cnt = 0;     // items processed by this warpgroup (o_final phase)
        long long* dbg = (a.dbg && threadIdx.x == 0) ? a.dbg + static_cast<size_t>(blockIdx.x) * 16 : nullptr;
        long long d_loop = 0, d_epi = 0, d_wait_s = 0, d_wait_o = 0, d_items = 0, d_wait_pv = 0;
        const long long d_start = dbg ? clock64() : 0;

#pragma unroll 1
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const Item5 ic = decode_item5(item, npairs, a.H, a.Tq);
            const uint32_t Tbase = T0;
            T0 += n;
            if (X == 1 && !ic.has_b) continue;
            float m_used = -INFINITY, l_run = 0.f;
            const long long d_t0 = dbg ? clock64() : 0;

#pragma unroll 1
            for (int j = 0; j < n; ++j, ++gt) {
                if (j == n - 1 && a.v_transform) {
                    // pull this row's output-rotation operands into L1 one key tile before the epilogue needs them
                    const int t_ = ic.p * 256 + X * 128 + r;
                    const int tt_ = t_ < a.Tq ? t_ : a.Tq - 1;
                    const size_t view_ = static_cast<size_t>(ic.b) * a.Nq + tt_ / a.tpvq;
                    if (a.hd.se3) prefetch_l1(a.se3_q + view_ * 16);
                    if (a.hd.so3) { prefetch_l1(a.so3_q + view_ * 34); prefetch_l1(a.so3_q + view_ * 34 + 32); }
                    if (a.hd.so2) {
                        const float* so2_ = a.so2_q + (static_cast<size_t>(ic.b) * a.Tq + tt_) * a.C * 2;
                        for (int off = 0; off < a.C * 2; off += 32) prefetch_l1(so2_ + off);
                    }
                }
                const long long d_w0 = dbg ? clock64() : 0;
                mbar_wait(&bars[L::bSFull + X], gt & 1);
                if (dbg) d_wait_s += clock64() - d_w0;
                tc_fence_after();
                uint32_t sreg[128];
                tmem_ld32(s_addr, sreg);
                tmem_ld32(s_addr + 32, sreg + 32);
                tmem_ld32(s_addr + 64, sreg + 64);
                tmem_ld32(s_addr + 96, sreg + 96);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(&bars[L::bSFree + X]);          // QK_X(j+1) may overwrite S_X now
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

                // PV_X of the previous tile must be complete before O_X is rescaled and before P_X is overwritten.
                if (gt > 0) {
                    const long long d_w1 = dbg ? clock64() : 0;
                    mbar_wait(&bars[L::bPVDone + X], (gt - 1) & 1);
                    if (dbg) d_wait_pv += clock64() - d_w1;
                    tc_fence_after();
                }
                const bool grow = (m_tile - m_used) * cs > k5RescaleThreshold;   // always true on the item's first tile
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

                const float neg = -m_used * cs;
                const uint64_t neg2 = pack_f32x2(neg, neg);
                uint64_t lsum2 = pack_f32x2(0.f, 0.f);
#pragma unroll
                for (int half = 0; half < 2; ++half) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const uint64_t x2 = ffma2(pack_f32x2(s[half * 64 + 2 * i], s[half * 64 + 2 * i + 1]), cs2, neg2);
                        float p0, p1;
                        if ((i % GTA_POLY_DEN) < GTA_POLY_NUM) {
                            poly_exp2x2(x2, p0, p1);
                        } else {
                            float x0, x1;
                            unpack_f32x2(x2, x0, x1);
                            p0 = fast_exp2(x0); p1 = fast_exp2(x1);
                        }
                        lsum2 = fadd2(lsum2, pack_f32x2(p0, p1));
                        sreg[half * 64 + i] = pack_bf16x2(p0, p1);      // in place (pair i is consumed before slot i)
                    }
                }
                float ls0, ls1;
                unpack_f32x2(lsum2, ls0, ls1);
                l_run += ls0 + ls1;
                // The ONE P buffer (64 TMEM columns) is shared by both query tiles in strict alternation
                // A(T) B(T) A(T+1) B(T+1) ...: use u = 2T + X waits until the PV product of use u-1 has consumed it
                // (for items without a B tile the B issuer passes the turn on).  This warpgroup's own previous use u-2
                // is already known complete through pv_done, so the parity wait is unambiguous.
                const uint32_t u = 2 * (Tbase + j) + X;
                if (u > 0) {
                    mbar_wait(&bars[L::bPFree], (u - 1) & 1);
                    tc_fence_after();
                }
                tmem_st32(pa_addr, sreg);
                tmem_st32(pa_addr + 32, sreg + 64);
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(&bars[L::bPFull + X]);
            }

            // ---- epilogue of this item: prefetch the row's reps, drain O to registers, release O, then finish.
            const long long d_t1 = dbg ? clock64() : 0;
            const int t = ic.p * 256 + X * 128 + r;
            const bool valid = t < a.Tq;
            const int tt = valid ? t : a.Tq - 1;
            // The output rotation walks the head row block type by block type with rolled loops (8 accumulator columns
            // per step straight from TMEM), so only ONE kind of rep data is live at a time: the view matrices are requested
            // before the wait for the last PV, the per-token SO(2) entries one chunk ahead of their use.  (A fully
            // unrolled epilogue kept M, W and all SO(2) chunks live next to 96 accumulator values and spilled ~150
            // local-memory loads per 32-column block.)
            const int c_se3 = a.hd.triv >> 3, n_se3 = a.hd.se3 >> 3, c_so3 = c_se3 + n_se3, n_so3 = a.hd.so3 >> 3;
            const int c_so2 = c_so3 + n_so3;
            // (all of this row's rep data was pulled into L1 one key tile ago, so each block loads its operands right
            //  before use and nothing has to stay live across the wait)
            const size_t view = static_cast<size_t>(ic.b) * a.Nq + tt / a.tpvq;
            const float* so2 = a.so2_q + (static_cast<size_t>(ic.b) * a.Tq + tt) * a.C * 2;
            mbar_wait(&bars[L::bOFinal + X], cnt & 1);
            const long long d_t2 = dbg ? clock64() : 0;
            ++cnt;
            tc_fence_after();
            const float inv_l = 1.0f / l_run;
            TOut* orow = reinterpret_cast<TOut*>(a.out) + ((static_cast<int64_t>(ic.b) * a.Tq + tt) * a.H + ic.h) * D;
            // O columns are fetched 8 at a time, one chunk AHEAD of their use (tcgen05.ld is asynchronous until
            // tcgen05.wait::ld), so the TMEM round trip overlaps the rotation of the previous chunk.
            uint32_t ocur[8];
            tmem_ld8(o_addr, ocur);
            auto next_o = [&](int c, float* x) {          // returns chunk c (already in flight), starts chunk c + 1
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = __uint_as_float(ocur[i]) * inv_l;
                if (c + 1 < D / 8) tmem_ld8(o_addr + (c + 1) * 8, ocur);   // consumed above; in-order issue makes the reuse safe
            };
            // 32-byte stores (STG.256): a thread owns a whole 2*D-byte output row, so every 16-byte store is its own
            // L1/L2 transaction (the v2 epilogue was bound by ~2.5 clk per such transaction); pairing two chunks halves
            // the transaction count and writes full sectors.  `pend` carries the even chunk across block-type sections.
            uint4 pend = make_uint4(0, 0, 0, 0);
            auto emit = [&](int c, const float* x) {
                if (sizeof(TOut) == 4) {
                    if (valid)
                        st_global_v8(orow + c * 8, make_uint4(__float_as_uint(x[0]), __float_as_uint(x[1]), __float_as_uint(x[2]), __float_as_uint(x[3])),
                                     make_uint4(__float_as_uint(x[4]), __float_as_uint(x[5]), __float_as_uint(x[6]), __float_as_uint(x[7])));
                } else {
                    const uint4 pk = pack_chunk_bf16(x);
                    if (c & 1) { if (valid) st_global_v8(orow + (c - 1) * 8, pend, pk); }
                    else pend = pk;
                }
            };
            const int c_rot = a.v_transform ? c_se3 : D / 8;       // chunks below c_rot are stored as they are
#pragma unroll 1
            for (int c = 0; c < c_rot; ++c) {
                float x[8];
                next_o(c, x);
                emit(c, x);
            }
            if (a.v_transform) {
                if (c_so3 > c_se3) {
                    float M[16];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 q4 = __ldg(reinterpret_cast<const float4*>(a.se3_q + view * 16) + i);
                        M[4 * i] = q4.x; M[4 * i + 1] = q4.y; M[4 * i + 2] = q4.z; M[4 * i + 3] = q4.w;
                    }
#pragma unroll 1
                    for (int c = c_se3; c < c_so3; ++c) {
                        float x[8];
                        next_o(c, x);
                        se3_apply(x, M, tc);
                        emit(c, x);
                    }
                }
                if (c_so2 > c_so3) {
                    float W[34];
#pragma unroll
                    for (int i = 0; i < 17; ++i) {
                        const float2 q2 = __ldg(reinterpret_cast<const float2*>(a.so3_q + view * 34) + i);
                        W[2 * i] = q2.x; W[2 * i + 1] = q2.y;
                    }
#pragma unroll 1
                    for (int c = c_so3; c < c_so2; ++c) {
                        float x[8];
                        next_o(c, x);
                        so3_apply<true>(x, W);
                        emit(c, x);
                    }
                }
                So2Chunk sc_cur = load_so2_chunk(so2, c_so2, a.hd);
#pragma unroll 1
                for (int c = c_so2; c < D / 8; ++c) {
                    So2Chunk sc_nxt = sc_cur;
                    if (c + 1 < D / 8) sc_nxt = load_so2_chunk(so2, c + 1, a.hd);
                    float x[8];
                    next_o(c, x);
                    const float cs8[8] = {sc_cur.a.x, sc_cur.a.y, sc_cur.a.z, sc_cur.a.w, sc_cur.b.x, sc_cur.b.y, sc_cur.b.z, sc_cur.b.w};
                    so2_apply<true>(x, cs8);
                    emit(c, x);
                    sc_cur = sc_nxt;
                }
            }
            tc_fence_before();
            mbar_arrive(&bars[L::bOFree + X]);                  // O_X fully read: the next item's PV_X(0) may overwrite it
            if (a.lse && valid)
                a.lse[(static_cast<int64_t>(ic.b) * a.H + ic.h) * a.Tq + t] = m_used * a.scale + logf(l_run);
            if (dbg) {
                const long long d_t3 = clock64();
                d_loop += d_t1 - d_t0; d_wait_o += d_t2 - d_t1; d_epi += d_t3 - d_t2; ++d_items;
            }
        }
        if (dbg) {
            dbg[0] = clock64() - d_start; dbg[1] = d_loop; dbg[2] = d_epi; dbg[3] = d_wait_s; dbg[4] = d_wait_o;
            dbg[5] = d_items; dbg[6] = d_wait_pv;
        }
    } else {
      setmaxnreg_dec<96>();
      if (warp == 11) {
        // =========================================================== Q stager (runs one item ahead), one warp
        uint32_t cntx[2] = {0, 0};          // items staged per tile slot (buffer = cnt & 1, phase = (cnt >> 1) & 1)
#pragma unroll 1
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const Item5 ic = decode_item5(item, npairs, a.H, a.Tq);
#pragma unroll 1
            for (int X = 0; X < 2; ++X) {
                if (X == 1 && !ic.has_b) continue;
                const uint32_t c_ = cntx[X]++;
                const int buf = c_ & 1;
                if (c_ >= 2) mbar_wait(&bars[L::bQFree + buf * 2 + X], ((c_ >> 1) - 1) & 1);
                uint8_t* sQ = smem + L::kQ + (buf * 2 + X) * L::kTile;
#pragma unroll 1
                for (int rr = 0; rr < 4; ++rr) {
                    const int r = lane + rr * 32;
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
      } else if (warp <= 9) {
            // ======================================================= UMMA issuers: warp 8 drives query tile A, warp 9
            // tile B.  Each is a straight-line stream  QK(0) { [s_free(j) -> QK(j+1)] [p_full(j) -> PV(j)] }  with
            // blocking waits on its own barriers only.  K'/V' stages are released by BOTH streams (k_empty/v_empty
            // have count 2): stream B walks through the tiles of items that have no B tile and releases them unused.
            const int X = warp - 8;
            constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);
            constexpr uint32_t idesc_pv = make_idesc_bf16(128, D, 0, 1);
            const uint32_t bar0 = smem_u32(bars);
            const uint32_t q_base = smem_u32(smem + L::kQ), k_base = smem_u32(smem + L::kK), v_base = smem_u32(smem + L::kV);
            const uint32_t s_tmem = tmem_base + k5TmemS + X * 128, o_tmem = tmem_base + k5TmemO + X * 96;
            uint32_t T = 0;        // CTA-global key tile index (K/V ring position)
            uint32_t gt = 0;       // key tiles this stream has processed (s_free / p_full phases)
            uint32_t cnt = 0;      // items this stream has processed (Q buffer, q_full / o_free phases)
            long long* dbg = (a.dbg && lane == 0) ? a.dbg + static_cast<size_t>(blockIdx.x) * 16 : nullptr;
            long long w_s = 0, w_p = 0, w_kv = 0;
#define GTA_TIMED_WAIT(acc, ...)                                 \
    do {                                                         \
        const long long t0_ = dbg ? clock64() : 0;               \
        __VA_ARGS__;                                             \
        if (dbg) acc += clock64() - t0_;                         \
    } while (0)

            auto issue_qk = [&](int j, uint32_t qlo, uint32_t qhi) {      // S_X = Q_X K'(T+j)^T
                const int s = (T + j) % NS;
                if (elect_one()) {
                    const uint64_t kd = desc_kmajor_sw64(k_base + s * L::kTile, 0);
                    const uint32_t klo = static_cast<uint32_t>(kd), khi = static_cast<uint32_t>(kd >> 32);
#pragma unroll
                    for (int kk = 0; kk < D / 16; ++kk)
                        umma_ss_lohi(s_tmem, qlo + kstep_kmajor_sw64(kk), qhi, klo + kstep_kmajor_sw64(kk), khi, idesc_qk, kk > 0);
                    umma_commit_addr(bar0 + (L::bKEmpty + s) * 8);
                    if (j == n - 1) umma_commit_addr(bar0 + (L::bQFree + (cnt & 1) * 2 + X) * 8);
                    umma_commit_addr(bar0 + (L::bSFull + X) * 8);
                }
                __syncwarp();
            };
            auto issue_pv = [&](int j) {                                    // O_X += P_X V'(T+j)
                const int s = (T + j) % NS;
                if (elect_one()) {
                    const uint64_t vd = desc_mnmajor_sw64(v_base + s * L::kTile, 0);
                    const uint32_t vlo = static_cast<uint32_t>(vd), vhi = static_cast<uint32_t>(vd >> 32);
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {
                        const uint32_t acc = (j > 0 || kk > 0) ? 1u : 0u;
                        umma_ts_lohi(o_tmem, tmem_base + k5TmemPA + kk * 8, vlo + kstep_mnmajor_sw64(kk), vhi, idesc_pv, acc);
                    }
                    umma_commit_addr(bar0 + L::bPFree * 8);
                    umma_commit_addr(bar0 + (L::bVEmpty + s) * 8);
                    umma_commit_addr(bar0 + (L::bPVDone + X) * 8);
                    if (j == n - 1) umma_commit_addr(bar0 + (L::bOFinal + X) * 8);
                }
                __syncwarp();
            };

#pragma unroll 1
            for (int item = blockIdx.x; item < nitems; item += gridDim.x, T += n) {
                const Item5 ic = decode_item5(item, npairs, a.H, a.Tq);
                const uint64_t qd = desc_kmajor_sw64(q_base + ((cnt & 1) * 2 + X) * L::kTile, 0);
                const uint32_t qlo = static_cast<uint32_t>(qd), qhi = static_cast<uint32_t>(qd >> 32);
                mbar_wait(&bars[L::bQFull + (cnt & 1) * 2 + X], (cnt >> 1) & 1);
                GTA_TIMED_WAIT(w_kv, mbar_wait(&bars[L::bKFull + T % NS], (T / NS) & 1));
                tc_fence_after();
                issue_qk(0, qlo, qhi);
#pragma unroll 1
                for (int j = 0; j < n; ++j) {
                    // S-step: S_X(j) has been copied to registers -> overwrite it with the next score tile
                    GTA_TIMED_WAIT(w_s, mbar_wait(&bars[L::bSFree + X], (gt + j) & 1));
                    if (j + 1 < n) {
                        GTA_TIMED_WAIT(w_kv, mbar_wait(&bars[L::bKFull + (T + j + 1) % NS], ((T + j + 1) / NS) & 1));
                        tc_fence_after();
                        issue_qk(j + 1, qlo, qhi);
                    }
                    // P-step: P_X(j) is ready (and O_X of the previous item has been read out)
                    GTA_TIMED_WAIT(w_p, mbar_wait(&bars[L::bPFull + X], (gt + j) & 1));
                    GTA_TIMED_WAIT(w_kv, mbar_wait(&bars[L::bVFull + (T + j) % NS], ((T + j) / NS) & 1));
                    if (j == 0 && cnt > 0) mbar_wait(&bars[L::bOFree + X], (cnt - 1) & 1);
                    tc_fence_after();
                    issue_pv(j);
                }
                gt += n; ++cnt;
            }
            if (dbg) { dbg[8 + X * 3] = w_s; dbg[9 + X * 3] = w_p; dbg[10 + X * 3] = w_kv; }
#undef GTA_TIMED_WAIT
      } else {
            // ======================================================= bulk-copy producer (warp 10)
            uint32_t gk = 0;
#pragma unroll 1
            for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
                const Item5 ic = decode_item5(item, npairs, a.H, a.Tq);
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
static int launch5_one(const AttnArgs& a, const GtaAttnParams& p, cudaStream_t st) {
    using L = Attn5Cfg<D>;
    auto kern = attn_fwd5_kernel<TIn, TOut, D>;
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
    kern<<<grid, kThreads5, L::kBytes, st>>>(a, npairs, static_cast<int>(nitems));
    return check_launch("gta_attn_fwd");
}

template <typename TIn, typename TOut>
static int launch5_d(const AttnArgs& a, const GtaAttnParams& p, cudaStream_t st) {
    switch (p.D) {
        case 32: return launch5_one<TIn, TOut, 32>(a, p, st);
        case 64: return launch5_one<TIn, TOut, 64>(a, p, st);
        case 96: return launch5_one<TIn, TOut, 96>(a, p, st);
    }
    return set_error(GTA_ERR_UNSUPPORTED, "persistent pipeline supports head dims 32/64/96");
}

int launch_attn_fwd_v4(const GtaAttnParams& p, cudaStream_t st) {
    const AttnArgs a = make_attn_args(p);
    const bool ib = p.in_dtype == GTA_DTYPE_BF16, ob = p.out_dtype == GTA_DTYPE_BF16;
    if (ib && ob) return launch5_d<__nv_bfloat16, __nv_bfloat16>(a, p, st);
    if (ib && !ob) return launch5_d<__nv_bfloat16, float>(a, p, st);
    if (!ib && ob) return launch5_d<float, __nv_bfloat16>(a, p, st);
    return launch5_d<float, float>(a, p, st);
}

}  // namespace gta
