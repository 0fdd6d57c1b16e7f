// Fused GTA attention forward, v3 pipeline (default for head dims <= 96): the persistent two-tile pipeline of
// gta_attn_fwd3.cu with the per-item PROLOGUE (Q staging) and EPILOGUE (1/l, rho_q^{-1}, store) moved off the
// softmax warpgroups onto a fourth warpgroup, so that the S -> P -> PV -> QK chain of the next item starts the moment
// the last key tile of the current one is done.  (Measured on the v2 pipeline: the epilogue held the softmax
// warpgroups 6.2 k of 37 k clocks per item at the MSN shape, during which the tensor pipe drained.)
//
//   warps 0-3   softmax warpgroup A   thread i <-> query row i of tile A <-> TMEM lane i
//   warps 4-7   softmax warpgroup B
//   warp  8     UMMA issuer (one elected lane)       warp 9   bulk-copy producer (K'/V' tile images, 2-stage ring)
//   warps 10-11 idle (they only keep the warpgroup's register budget small)
//   warps 12-15 pre/post warpgroup: thread i <-> row i of BOTH tiles.  Per item k:
//                 epilogue(k): wait (m,l) of the row + the last PV, drain O_X from TMEM 8 columns at a time, apply
//                              1/l and rho_q^{-1} in fp32 registers, write the bf16 row into the Q' buffer the item
//                              just finished with (free: its last QK has completed), release O_X, and hand the row to
//                              the bulk-copy engine (cp.async.bulk shared -> global, one 2*D-byte row per thread);
//                 stage(k+2):  load the raw strided Q rows of the item after next, apply rho_q^{-T}, write the bf16
//                              UMMA operand tiles into the same (now drained) Q' buffers.
// Register split (setmaxnreg, 512 threads, 65 536 registers): softmax 176, issuer/producer warpgroup 80, pre/post 80
// (no spills in the softmax loop or the issue loop; checked in SASS per USETMAXREG region).
//
// GTA_SPLIT_P (compile-time): the softmax publishes P in two 64-key halves so the first four PV MMAs of a tile run
// while the second half of the exponentials is still being computed.
//
// Reference semantics: source/utils/gta.py:92-279 and source/layers.py:202-211.
#include <cmath>

#include "attn_common.cuh"

namespace gta {

constexpr int kThreads6 = 512;
constexpr uint32_t k6TmemSA = 0, k6TmemSB = 128, k6TmemOA = 256, k6TmemOB = 384;
constexpr float k6RescaleThreshold = 8.0f;   // log2 units
#ifndef GTA_POLY_NUM
#define GTA_POLY_NUM 0
#endif
#ifndef GTA_POLY_DEN
#define GTA_POLY_DEN 4
#endif
// register split (setmaxnreg): 256 * SOFTMAX + 128 * ISSUE + 128 * POST must equal 65 536
#ifndef GTA_V3_REGS_SOFTMAX
#define GTA_V3_REGS_SOFTMAX 176
#define GTA_V3_REGS_ISSUE 80
#define GTA_V3_REGS_POST 80
#endif
#ifndef GTA_SPLIT_P
#define GTA_SPLIT_P 1
#endif

__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int D>
struct Attn6Cfg {
    static constexpr int kStages = 2;
    static constexpr uint32_t kTile = 128u * D * 2u;
    static constexpr uint32_t kQ = 0;                          // [2 buffers][2 tiles]; doubles as the output staging
    static constexpr uint32_t kK = 4 * kTile;                  // [kStages]
    static constexpr uint32_t kV = kTile * (4 + kStages);      // [kStages]
    static constexpr uint32_t kLM = kTile * (4 + 2 * kStages); // float2 (m, l) [X][buf][128]
    static constexpr uint32_t kBars = kLM + 2 * 2 * 128 * 8;
    enum : int {
        bQFull = 0,                        // [buf][X]  count 128 (pre/post warpgroup)
        bQFree = 4,                        // [buf][X]  tcgen05.commit after the item's last QK_X
        bKFull = 8,                        // [kStages]
        bVFull = bKFull + kStages,
        bKEmpty = bVFull + kStages,
        bVEmpty = bKEmpty + kStages,
        bSFull = bVEmpty + kStages,        // [X] commit
        bPHalf = bSFull + 2,               // [X] count 128: keys 0..63 of P_X published (GTA_SPLIT_P)
        bPFull = bPHalf + 2,               // [X] count 128: all of P_X published
        bOFinal = bPFull + 2,              // [X] commit after the item's last PV_X
        bOFree = bOFinal + 2,              // [X] count 128: O_X drained
        bLFull = bOFree + 2,               // [X] count 128: (m, l) of the item's rows written
        bCount = bLFull + 2
    };
    static constexpr uint32_t kTmemSlot = kBars + bCount * 8;
    static constexpr uint32_t kUsed = kTmemSlot + 16;
    // at least 120 KB so that only one CTA fits an SM (the kernel allocates all 512 TMEM columns)
    static constexpr uint32_t kBytes = (kUsed + 1024 > 120u * 1024u) ? kUsed + 1024 : 120u * 1024u;
};

struct ItemCoord6 {
    int b, h, p;
    bool has_b;
};
__device__ __forceinline__ ItemCoord6 decode_item6(int item, int npairs, int H, int Tq) {
    ItemCoord6 c;
    c.p = item % npairs;
    const int bh = item / npairs;
    c.h = bh % H;
    c.b = bh / H;
    c.has_b = (c.p * 256 + 128) < Tq;
    return c;
}

template <typename TIn, typename TOut, int D>
__global__ void __launch_bounds__(kThreads6, 1) attn_fwd6_kernel(const AttnArgs a, const int npairs, const int nitems) {
    using L = Attn6Cfg<D>;
    constexpr int NS = L::kStages;
    constexpr bool kBulkOut = sizeof(TOut) == 2;     // bf16 rows are staged in shared memory and bulk-copied out
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBars);
    float2* lm = reinterpret_cast<float2*>(smem + L::kLM);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kTmemSlot);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = a.ntiles_k;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) {
            mbar_init(&bars[L::bQFull + i], 128);
            mbar_init(&bars[L::bQFree + i], 1);
        }
        for (int x = 0; x < 2; ++x) {
            mbar_init(&bars[L::bSFull + x], 1);
            mbar_init(&bars[L::bPHalf + x], 128);
            mbar_init(&bars[L::bPFull + x], 128);
            mbar_init(&bars[L::bOFinal + x], 1);
            mbar_init(&bars[L::bOFree + x], 128);
            mbar_init(&bars[L::bLFull + x], 128);
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
        // =========================================================== softmax warpgroups
        setmaxnreg_inc<GTA_V3_REGS_SOFTMAX>();
        const int X = warp >> 2;
        const int r = threadIdx.x & 127;
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
        const uint32_t s_addr = lane_base + (X ? k6TmemSB : k6TmemSA);
        const uint32_t o_addr = lane_base + (X ? k6TmemOB : k6TmemOA);
        const float cs = a.scale_log2;
        const uint64_t cs2 = pack_f32x2(cs, cs);
        uint32_t gt = 0;      // tiles processed by this warpgroup (s_full / p_full phase)
        uint32_t cnt = 0;     // items processed by this warpgroup ((m,l) slot)
        long long* dbg = (a.dbg && threadIdx.x == 0) ? a.dbg + static_cast<size_t>(blockIdx.x) * 16 : nullptr;
        long long d_wait_s = 0, d_items = 0;
        const long long d_start = dbg ? clock64() : 0;

#pragma unroll 1
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const ItemCoord6 ic = decode_item6(item, npairs, a.H, a.Tq);
            if (X == 1 && !ic.has_b) continue;
            float m_used = -INFINITY, l_run = 0.f;

#pragma unroll 1
            for (int j = 0; j < n; ++j, ++gt) {
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

                const bool grow = (m_tile - m_used) * cs > k6RescaleThreshold;   // always true on the item's first tile
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
                    // packed in place: P pair i overwrites sreg[half*64 + i] after s[half*64 + 2i], s[.. + 2i+1] were consumed
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
                        sreg[half * 64 + i] = pack_bf16x2(p0, p1);
                    }
                    tmem_st32(s_addr + half * 32, sreg + half * 64);
                    if (GTA_SPLIT_P && half == 0) {
                        tmem_st_wait();
                        tc_fence_before();
                        mbar_arrive(&bars[L::bPHalf + X]);
                    }
                }
                float ls0, ls1;
                unpack_f32x2(lsum2, ls0, ls1);
                l_run += ls0 + ls1;
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(&bars[L::bPFull + X]);
            }
            // hand the row statistics to the pre/post warpgroup and move straight on to the next item
            lm[(X * 2 + (cnt & 1)) * 128 + r] = make_float2(m_used, l_run);
            mbar_arrive(&bars[L::bLFull + X]);
            ++cnt;
            ++d_items;
        }
        if (dbg) { dbg[0] = clock64() - d_start; dbg[3] = d_wait_s; dbg[5] = d_items; }
    } else if (warp >= 12) {
        // =========================================================== pre/post warpgroup: Q staging + epilogue
        setmaxnreg_dec<GTA_V3_REGS_POST>();
        const int r = threadIdx.x - 384;                       // row of both tiles / TMEM lane
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
        uint32_t sc[2] = {0, 0};           // items staged per tile slot  (buffer = sc & 1)
        uint32_t ec[2] = {0, 0};           // epilogues done per tile slot
        long long* dbg = (a.dbg && r == 0) ? a.dbg + static_cast<size_t>(blockIdx.x) * 16 : nullptr;
        long long d_epi = 0, d_wait_o = 0, d_stage = 0;

        auto stage_item = [&](int item) {
            const ItemCoord6 ic = decode_item6(item, npairs, a.H, a.Tq);
#pragma unroll 1
            for (int X = 0; X < 2; ++X) {
                if (X == 1 && !ic.has_b) continue;
                const uint32_t c_ = sc[X]++;
                const int buf = c_ & 1;
                // the buffer's previous occupant: its last QK has completed (q_free) and, for bf16 output, this
                // warpgroup's bulk stores out of it have finished reading (wait_group.read + barrier, done by the caller)
                if (c_ >= 2) mbar_wait(&bars[L::bQFree + buf * 2 + X], ((c_ >> 1) - 1) & 1);
                uint8_t* sQ = smem + L::kQ + (buf * 2 + X) * L::kTile;
                const int t = ic.p * 256 + X * 128 + r;
                const bool valid = t < a.Tq;
                const int tt = valid ? t : a.Tq - 1;
                const size_t view = static_cast<size_t>(ic.b) * a.Nq + tt / a.tpvq;
                const float* so2 = a.so2_q + (static_cast<size_t>(ic.b) * a.Tq + tt) * a.C * 2;
                const TIn* qrow = reinterpret_cast<const TIn*>(a.q) + static_cast<int64_t>(ic.b) * a.q_sb +
                                  static_cast<int64_t>(ic.h) * a.q_sh + static_cast<int64_t>(tt) * a.q_st;
                constexpr int NC = D / 8;
                constexpr int G = (sizeof(TIn) == 2) ? ((NC % 6 == 0) ? 6 : 4) : ((NC % 3 == 0) ? 3 : 2);   // <= 24 registers of raw data
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
                fence_proxy_async_smem();
                mbar_arrive(&bars[L::bQFull + buf * 2 + X]);
            }
        };

        auto epilogue_item = [&](int item) {
            const ItemCoord6 ic = decode_item6(item, npairs, a.H, a.Tq);
            const int c_se3 = a.hd.triv >> 3, n_se3 = a.hd.se3 >> 3, c_so3 = c_se3 + n_se3, n_so3 = a.hd.so3 >> 3;
            const int c_so2 = c_so3 + n_so3;
#pragma unroll 1
            for (int X = 0; X < 2; ++X) {
                if (X == 1 && !ic.has_b) continue;
                const uint32_t c_ = ec[X]++;
                const int buf = c_ & 1;
                const uint32_t o_addr = lane_base + (X ? k6TmemOB : k6TmemOA);
                const int t = ic.p * 256 + X * 128 + r;
                const bool valid = t < a.Tq;
                const int tt = valid ? t : a.Tq - 1;
                const size_t view = static_cast<size_t>(ic.b) * a.Nq + tt / a.tpvq;
                const float* so2 = a.so2_q + (static_cast<size_t>(ic.b) * a.Tq + tt) * a.C * 2;
                if (a.v_transform) {       // the row's output-rotation operands: pull them into L1 ahead of the waits
                    if (a.hd.se3) prefetch_l1(a.se3_q + view * 16);
                    if (a.hd.so3) { prefetch_l1(a.so3_q + view * 34); prefetch_l1(a.so3_q + view * 34 + 32); }
                    if (a.hd.so2)
                        for (int off = 0; off < a.C * 2; off += 32) prefetch_l1(so2 + off);
                }
                const long long d_t0 = dbg ? clock64() : 0;
                mbar_wait(&bars[L::bLFull + X], c_ & 1);
                const float2 ml = lm[(X * 2 + buf) * 128 + r];
                // staging area of this tile's output rows = the Q' buffer the item used (all its QK MMAs are complete)
                mbar_wait(&bars[L::bQFree + buf * 2 + X], (c_ >> 1) & 1);
                uint8_t* srow = smem + L::kQ + (buf * 2 + X) * L::kTile + static_cast<uint32_t>(r) * (D * 2);
                mbar_wait(&bars[L::bOFinal + X], c_ & 1);
                const long long d_t1 = dbg ? clock64() : 0;
                tc_fence_after();
                const float inv_l = 1.0f / ml.y;
                TOut* orow = reinterpret_cast<TOut*>(a.out) + ((static_cast<int64_t>(ic.b) * a.Tq + tt) * a.H + ic.h) * D;
                // O columns are fetched 8 at a time, one chunk AHEAD of their use (tcgen05.ld is asynchronous until
                // tcgen05.wait::ld), so the TMEM round trip overlaps the rotation of the previous chunk.
                uint32_t ocur[8];
                tmem_ld8(o_addr, ocur);
                auto next_o = [&](int c, float* x) {          // returns chunk c (already in flight), starts chunk c + 1
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 8; ++i) x[i] = __uint_as_float(ocur[i]) * inv_l;
                    if (c + 1 < D / 8) tmem_ld8(o_addr + (c + 1) * 8, ocur);
                };
                auto emit = [&](int c, const float* x) {
                    if (kBulkOut) {
                        *reinterpret_cast<uint4*>(srow + c * 16) = pack_chunk_bf16(x);
                    } else if (valid) {
                        st_global_v8(orow + c * 8, make_uint4(__float_as_uint(x[0]), __float_as_uint(x[1]), __float_as_uint(x[2]), __float_as_uint(x[3])),
                                     make_uint4(__float_as_uint(x[4]), __float_as_uint(x[5]), __float_as_uint(x[6]), __float_as_uint(x[7])));
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
                mbar_arrive(&bars[L::bOFree + X]);              // O_X fully read: the next item's PV_X(0) may overwrite it
                if (kBulkOut) {
                    fence_proxy_async_smem();                   // this thread's row -> visible to the bulk-copy engine
                    if (valid) bulk_s2g(orow, srow, D * 2);
                    bulk_commit_group();
                }
                if (a.lse && valid)
                    a.lse[(static_cast<int64_t>(ic.b) * a.H + ic.h) * a.Tq + t] = ml.x * a.scale + logf(ml.y);
                if (dbg) { d_wait_o += d_t1 - d_t0; d_epi += clock64() - d_t1; }
            }
        };

        // two items of look-ahead for Q, then per item: epilogue(k), stage(k + 2)
        int it0 = blockIdx.x;
        if (it0 < nitems) stage_item(it0);
        if (it0 + static_cast<int>(gridDim.x) < nitems) stage_item(it0 + gridDim.x);
#pragma unroll 1
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            epilogue_item(item);
            const long long nxt = static_cast<long long>(item) + 2LL * gridDim.x;
            if (nxt < nitems) {
                const long long d_s0 = dbg ? clock64() : 0;
                if (kBulkOut) {
                    bulk_wait_group_read0();                    // own rows left shared memory ...
                    named_bar_sync(1, 128);                     // ... and so did everybody else's
                }
                stage_item(static_cast<int>(nxt));
                if (dbg) d_stage += clock64() - d_s0;
            }
        }
        if (kBulkOut) bulk_wait_group0();
        if (dbg) { dbg[2] = d_epi; dbg[4] = d_wait_o; dbg[6] = d_stage; }
    } else {
      setmaxnreg_dec<GTA_V3_REGS_ISSUE>();
      if (warp == 8) {
            // ======================================================= UMMA issuer
            constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);
            constexpr uint32_t idesc_pv = make_idesc_bf16(128, D, 0, 1);
            uint32_t gk = 0;                   // global key-tile counter of this CTA (K/V ring position)
            uint32_t gtx[2] = {0, 0};          // tiles per softmax warpgroup (p_full phase)
            uint32_t cntx[2] = {0, 0};         // items per tile slot (Q buffer / o_free phase)
            const uint32_t bar0 = smem_u32(bars);
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
                const ItemCoord6 ic = decode_item6(item, npairs, a.H, a.Tq);
                const int nx = ic.has_b ? 2 : 1;
                uint32_t q_addr[2];
                for (int X = 0; X < nx; ++X) q_addr[X] = smem_u32(smem + L::kQ + ((cntx[X] & 1) * 2 + X) * L::kTile);

                auto issue_qk = [&](int X, int j) {
                    const int s = (gk + j) % NS;
                    if (elect_one()) {
                        const uint64_t qd = desc_kmajor_sw64(q_addr[X], 0);
                        const uint64_t kd = desc_kmajor_sw64(smem_u32(smem + L::kK + s * L::kTile), 0);
                        const uint32_t qlo = static_cast<uint32_t>(qd), qhi = static_cast<uint32_t>(qd >> 32);
                        const uint32_t klo = static_cast<uint32_t>(kd), khi = static_cast<uint32_t>(kd >> 32);
                        const uint32_t d_addr = tmem_base + (X ? k6TmemSB : k6TmemSA);
#pragma unroll
                        for (int kk = 0; kk < D / 16; ++kk)
                            umma_ss_lohi(d_addr, qlo + kstep_kmajor_sw64(kk), qhi, klo + kstep_kmajor_sw64(kk), khi, idesc_qk, kk > 0);
                        if (X == nx - 1) umma_commit_addr(bar0 + (L::bKEmpty + s) * 8);
                        if (j == n - 1) umma_commit_addr(bar0 + (L::bQFree + (cntx[X] & 1) * 2 + X) * 8);
                        umma_commit_addr(bar0 + (L::bSFull + X) * 8);
                    }
                    __syncwarp();
                };
                auto issue_pv = [&](int X, int j) {
                    const int s = (gk + j) % NS;
                    const uint32_t par = (gtx[X] + j) & 1;
                    const uint64_t vd = desc_mnmajor_sw64(smem_u32(smem + L::kV + s * L::kTile), 0);
                    const uint32_t vlo = static_cast<uint32_t>(vd), vhi = static_cast<uint32_t>(vd >> 32);
                    const uint32_t d_addr = tmem_base + (X ? k6TmemOB : k6TmemOA);
                    const uint32_t p_addr = tmem_base + (X ? k6TmemSB : k6TmemSA);
                    if (GTA_SPLIT_P) {
                        GTA_TIMED_WAIT(w_p, mbar_wait(&bars[L::bPHalf + X], par));
                        if (j == 0 && cntx[X] > 0) GTA_TIMED_WAIT(w_of, mbar_wait(&bars[L::bOFree + X], (cntx[X] - 1) & 1));
                        tc_fence_after();
                        if (elect_one()) {
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk)
                                umma_ts_lohi(d_addr, p_addr + kk * 8, vlo + kstep_mnmajor_sw64(kk), vhi, idesc_pv,
                                             (j > 0 || kk > 0) ? 1u : 0u);
                        }
                        __syncwarp();
                        GTA_TIMED_WAIT(w_p, mbar_wait(&bars[L::bPFull + X], par));
                        tc_fence_after();
                        if (elect_one()) {
#pragma unroll
                            for (int kk = 4; kk < 8; ++kk)
                                umma_ts_lohi(d_addr, p_addr + kk * 8, vlo + kstep_mnmajor_sw64(kk), vhi, idesc_pv, 1u);
                            if (X == nx - 1) umma_commit_addr(bar0 + (L::bVEmpty + s) * 8);
                            if (j == n - 1) umma_commit_addr(bar0 + (L::bOFinal + X) * 8);
                        }
                        __syncwarp();
                    } else {
                        GTA_TIMED_WAIT(w_p, mbar_wait(&bars[L::bPFull + X], par));
                        if (j == 0 && cntx[X] > 0) GTA_TIMED_WAIT(w_of, mbar_wait(&bars[L::bOFree + X], (cntx[X] - 1) & 1));
                        tc_fence_after();
                        if (elect_one()) {
#pragma unroll
                            for (int kk = 0; kk < 8; ++kk)
                                umma_ts_lohi(d_addr, p_addr + kk * 8, vlo + kstep_mnmajor_sw64(kk), vhi, idesc_pv,
                                             (j > 0 || kk > 0) ? 1u : 0u);
                            if (X == nx - 1) umma_commit_addr(bar0 + (L::bVEmpty + s) * 8);
                            if (j == n - 1) umma_commit_addr(bar0 + (L::bOFinal + X) * 8);
                        }
                        __syncwarp();
                    }
                };

                GTA_TIMED_WAIT(w_k, mbar_wait(&bars[L::bKFull + gk % NS], (gk / NS) & 1));
                for (int X = 0; X < nx; ++X) {
                    const uint32_t c_ = cntx[X];
                    GTA_TIMED_WAIT(w_q, mbar_wait(&bars[L::bQFull + (c_ & 1) * 2 + X], (c_ >> 1) & 1));
                    tc_fence_after();
                    issue_qk(X, 0);
                }
#pragma unroll 1
                for (int j = 0; j < n; ++j) {
                    GTA_TIMED_WAIT(w_v, mbar_wait(&bars[L::bVFull + (gk + j) % NS], ((gk + j) / NS) & 1));
                    if (j + 1 < n) GTA_TIMED_WAIT(w_k, mbar_wait(&bars[L::bKFull + (gk + j + 1) % NS], ((gk + j + 1) / NS) & 1));
                    for (int X = 0; X < nx; ++X) {
                        issue_pv(X, j);
                        if (j + 1 < n) issue_qk(X, j + 1);
                    }
                }
                gk += n;
                for (int X = 0; X < nx; ++X) { gtx[X] += n; ++cntx[X]; }
            }
            if (dbg) { dbg[8] = w_k; dbg[9] = w_v; dbg[10] = w_p; dbg[11] = w_of; dbg[12] = w_q; }
#undef GTA_TIMED_WAIT
      } else if (warp == 9) {
            // ======================================================= bulk-copy producer
            uint32_t gk = 0;
#pragma unroll 1
            for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
                const ItemCoord6 ic = decode_item6(item, npairs, a.H, a.Tq);
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
static int launch6_one(const AttnArgs& a, const GtaAttnParams& p, cudaStream_t st) {
    using L = Attn6Cfg<D>;
    auto kern = attn_fwd6_kernel<TIn, TOut, D>;
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
    if (nitems > 0x3fffffffLL) return set_error(GTA_ERR_UNSUPPORTED, "too many work items");
    const int grid = static_cast<int>(nitems < num_sms ? nitems : num_sms);
    kern<<<grid, kThreads6, L::kBytes, st>>>(a, npairs, static_cast<int>(nitems));
    return check_launch("gta_attn_fwd");
}

template <typename TIn, typename TOut>
static int launch6_d(const AttnArgs& a, const GtaAttnParams& p, cudaStream_t st) {
    switch (p.D) {
        case 32: return launch6_one<TIn, TOut, 32>(a, p, st);
        case 64: return launch6_one<TIn, TOut, 64>(a, p, st);
        case 96: return launch6_one<TIn, TOut, 96>(a, p, st);
    }
    return set_error(GTA_ERR_UNSUPPORTED, "persistent pipeline supports head dims 32/64/96");
}

int launch_attn_fwd_v3(const GtaAttnParams& p, cudaStream_t st) {
    const AttnArgs a = make_attn_args(p);
    const bool ib = p.in_dtype == GTA_DTYPE_BF16, ob = p.out_dtype == GTA_DTYPE_BF16;
    if (ib && ob) return launch6_d<__nv_bfloat16, __nv_bfloat16>(a, p, st);
    if (ib && !ob) return launch6_d<__nv_bfloat16, float>(a, p, st);
    if (!ib && ob) return launch6_d<float, __nv_bfloat16>(a, p, st);
    return launch6_d<float, float>(a, p, st);
}

}  // namespace gta
