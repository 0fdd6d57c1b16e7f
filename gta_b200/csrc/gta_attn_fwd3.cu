// Fused GTA attention forward, v2 pipeline: PERSISTENT CTAs (one per SM), two 128-query tiles per work item,
// Q staging of the NEXT item and the epilogue of the PREVIOUS item overlapped with the tensor-core main loop.
//
// Work item = (batch b, head h, query pair p): 256 query rows of one (b,h) against all Tk keys.
// CTA c processes items c, c+grid, c+2*grid, ... (all items cost the same, so static striding is balanced).
//
//   warps 0-3   softmax warpgroup A   thread i <-> query row i of tile A <-> TMEM lane i
//   warps 4-7   softmax warpgroup B
//   warp  8     UMMA issuer (one lane)
//   warp  9     bulk-copy producer for the K'/V' tile images (2-stage ring, runs ahead across item boundaries)
//   warps 10-11 Q stager: loads the raw strided Q rows of the NEXT item, applies rho_q^{-T} in fp32 registers and
//               writes the bf16 UMMA operand tiles into the other half of a double-buffered Q area
// Register split (setmaxnreg): softmax warpgroups 200, the third warpgroup 96 (384 threads, 65 536 registers).
//
// Per key tile the tensor pipe executes  PV_A(j) QK_A(j+1) PV_B(j) QK_B(j+1)  (see gta_attn_fwd2.cu); at an item
// boundary QK_X(0) of the next item is issued right after PV_X(n-1), so S of the next item is ready while the
// softmax warpgroup is still writing out the previous item's O.  O_X in TMEM is recycled through o_final/o_free.
//
// Reference semantics: source/utils/gta.py:92-279 and source/layers.py:202-211.
#include <cmath>

#include "attn_common.cuh"

namespace gta {

constexpr int kThreads3 = 384;
constexpr int kStagerThreads = 64;
constexpr uint32_t k3TmemSA = 0, k3TmemSB = 128, k3TmemOA = 256, k3TmemOB = 384;
constexpr float k3RescaleThreshold = 8.0f;   // log2 units
// Fraction of the exponentials evaluated with poly_exp2x2 instead of MUFU.EX2: pairs with (i % DEN) < NUM.
#ifndef GTA_POLY_NUM
#define GTA_POLY_NUM 0
#endif
#ifndef GTA_POLY_DEN
#define GTA_POLY_DEN 4
#endif
#ifndef GTA_OPTIMISTIC
#define GTA_OPTIMISTIC 0
#endif
// Q stager: raw q rows loaded with the L1 evict-first priority (measured variant, off by default).
#ifndef GTA_Q_EVICT_FIRST
#define GTA_Q_EVICT_FIRST 0
#endif
// Q stager: L1 prefetch of the next row (head dims <= 64), see the stager loop.
#ifndef GTA_Q_ROW_PREFETCH
#define GTA_Q_ROW_PREFETCH 1
#endif
constexpr bool kQRowPrefetch = GTA_Q_ROW_PREFETCH != 0;
// Epilogue: SO(2) table rows staged in shared memory by coalesced asynchronous copies (see the softmax loop's last key tile).
// MEASURED, OFF BY DEFAULT (-DGTA_SO2_STAGE=1 compiles it in): the so2 section of the epilogue drops from 2.97 k to 1.82 k clk per
// item at MSN, but the 28 KB of rows take the CTA from 198 to 226 KB of shared memory, i.e. they halve what is left of the
// 256 KB array as L1 for the stager's and the epilogue's global loads: the whole step is 3.5 % SLOWER at the MSN encoder
// shape (0.534 vs 0.516 ms, same box, profiles/r02_so2_stage_ab2.txt); neutral at D = 64 where shared memory is plentiful.
#ifndef GTA_SO2_STAGE
#define GTA_SO2_STAGE 0
#endif
constexpr bool kSo2Stage = GTA_SO2_STAGE != 0;

template <int D>
struct Attn3Cfg {
    static constexpr int kStages = 2;
    static constexpr uint32_t kTile = 128u * D * 2u;
    static constexpr uint32_t kQ = 0;                          // [2 buffers][2 tiles]
    static constexpr uint32_t kK = 4 * kTile;                  // [kStages]
    static constexpr uint32_t kV = kTile * (4 + kStages);      // [kStages]
    static constexpr uint32_t kBars = kTile * (4 + 2 * kStages);
    enum : int {
        bQFull = 0,                        // [buf][X]  count 128 (stager threads)
        bQFree = 4,                        // [buf][X]  tcgen05.commit after the item's last QK_X
        bKFull = 8,                        // [kStages]
        bVFull = bKFull + kStages,
        bKEmpty = bVFull + kStages,
        bVEmpty = bKEmpty + kStages,
        bSFull = bVEmpty + kStages,        // [X] commit
        bPFull = bSFull + 2,               // [X] count 128
        bOFinal = bPFull + 2,              // [X] commit after the item's last PV_X
        bOFree = bOFinal + 2,              // [X] count 128: O_X drained to registers
        bCount = bOFree + 2
    };
    static constexpr uint32_t kTmemSlot = kBars + bCount * 8;
    static constexpr uint32_t kUsed = kTmemSlot + 16;
    static constexpr uint32_t kBytes = (kUsed + 1024 > 120u * 1024u) ? kUsed + 1024 : 120u * 1024u;
    // Optional tail (launch3_one adds it to the dynamic size when it fits): per softmax warp 32 rows of the tokens' SO(2)
    // (cos, sin) table, row stride so2 + 4 floats (conflict-free 16-byte reads of a thread's own row).
    static constexpr uint32_t kSo2 = (kUsed + 15u) & ~15u;
    static constexpr uint32_t so2_stage_bytes(int so2_dims) { return 8u * 32u * static_cast<uint32_t>(so2_dims + 4) * 4u; }
};

struct ItemCoord {
    int b, h, p;
    bool has_b;
};
__device__ __forceinline__ ItemCoord decode_item(int item, int npairs, int H, int Tq) {
    ItemCoord c;
    c.p = item % npairs;
    const int bh = item / npairs;
    c.h = bh % H;
    c.b = bh / H;
    c.has_b = (c.p * 256 + 128) < Tq;
    return c;
}

template <typename TIn, typename TOut, int D>
__global__ void __launch_bounds__(kThreads3, 1) attn_fwd3_kernel(const AttnArgs a, const int npairs, const int nitems) {
    using L = Attn3Cfg<D>;
    constexpr int NS = L::kStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBars);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kTmemSlot);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = a.ntiles_k;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) {
            mbar_init(&bars[L::bQFull + i], kStagerThreads);
            mbar_init(&bars[L::bQFree + i], 1);
        }
        for (int x = 0; x < 2; ++x) {
            mbar_init(&bars[L::bSFull + x], 1);
            mbar_init(&bars[L::bPFull + x], 128);
            mbar_init(&bars[L::bOFinal + x], 1);
            mbar_init(&bars[L::bOFree + x], 128);
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
        const uint32_t s_addr = lane_base + (X ? k3TmemSB : k3TmemSA);
        const uint32_t o_addr = lane_base + (X ? k3TmemOB : k3TmemOA);
        const float cs = a.scale_log2;
        const uint64_t cs2 = pack_f32x2(cs, cs);
        uint32_t gt = 0;      // tiles processed by this warpgroup (s_full / p_full phase)
        uint32_t cnt = 0;     // items processed by this warpgroup (o_final phase)
        // optional phase clocks (GtaAttnParams.debug_clocks): [cta][16] accumulated over the CTA's items
        long long* dbg = (a.dbg && threadIdx.x == 0) ? a.dbg + static_cast<size_t>(blockIdx.x) * 16 : nullptr;
        long long d_loop = 0, d_epi = 0, d_wait_s = 0, d_wait_o = 0, d_items = 0;
        long long d_e[4] = {0, 0, 0, 0};
        const long long d_start = dbg ? clock64() : 0;

#pragma unroll 1
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const ItemCoord ic = decode_item(item, npairs, a.H, a.Tq);
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
                    if (kSo2Stage && a.hd.so2 && a.so2_stage) {
                        float* so2_sm_warp = reinterpret_cast<float*>(smem + L::kSo2) + warp * 32 * a.so2_stage;
                        // the warp's 32 token rows of the SO(2) table are one contiguous block: copy it with coalesced 16-byte
                        // asynchronous copies (lane = consecutive float4) instead of letting every thread fetch its own row in
                        // the epilogue (a warp-wide LDG.128 of 32 different lines costs 32 L1 wavefronts: the so2 chunks took
                        // 830 clk each against 270 for the SE(3) chunks)
                        __syncwarp();                            // every lane is done with the previous item's rows
                        const int q4 = a.hd.so2 >> 2;            // float4 per token
                        const int tb = ic.p * 256 + X * 128 + (warp & 3) * 32;
                        const int dq = 32 / q4, dm = 32 - dq * q4;
                        int row = lane / q4, col = lane - row * q4;
                        for (int i = 0; i < q4; ++i) {
                            const int tk_ = tb + row < a.Tq ? tb + row : a.Tq - 1;
                            cp_async16(so2_sm_warp + row * a.so2_stage + col * 4,
                                       a.so2_q + (static_cast<size_t>(ic.b) * a.Tq + tk_) * a.hd.so2 + col * 4);
                            col += dm; row += dq;
                            if (col >= q4) { col -= q4; ++row; }
                        }
                    } else if (a.hd.so2) {
                        const float* so2_ = a.so2_q + (static_cast<size_t>(ic.b) * a.Tq + tt_) * a.C * 2;
                        for (int off = 0; off < a.C * 2; off += 32) prefetch_l1(so2_ + off);
                    }
                }
                const long long d_w0 = dbg ? clock64() : 0;
                mbar_wait(&bars[L::bSFull + X], gt & 1);
                if (dbg) d_wait_s += clock64() - d_w0;
                tc_fence_after();
                uint32_t sreg[128];
#if GTA_OPTIMISTIC
                // measured variant, off by default (softmax_tile_optimistic, attn_common.cuh): every full tile after the
                // item's first takes its exponentials against the current reference while the scores still stream in
                if (j > 0 && (j < n - 1 || a.Tk - j * 128 == 128) &&
                    softmax_tile_optimistic<GTA_POLY_NUM, GTA_POLY_DEN>(s_addr, sreg, m_used, cs, cs2, k3RescaleThreshold, l_run)) {
                    tmem_st_wait();
                    tc_fence_before();
                    mbar_arrive(&bars[L::bPFull + X]);
                    continue;
                }
#endif
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

                const bool grow = (m_tile - m_used) * cs > k3RescaleThreshold;   // always true on the item's first tile
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
                    // packed in place: P pair i overwrites sreg[half*64 + i] after s[half*64 + 2i], s[.. + 2i+1] were consumed,
                    // so the store reuses the register block the load filled (no second 32-register block is needed)
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
                }
                float ls0, ls1;
                unpack_f32x2(lsum2, ls0, ls1);
                l_run += ls0 + ls1;
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
            const long long e0 = dbg ? clock64() : 0;
            const int c_rot = a.v_transform ? c_se3 : D / 8;       // chunks below c_rot are stored as they are
#pragma unroll 1
            for (int c = 0; c < c_rot; ++c) {
                float x[8];
                next_o(c, x);
                emit(c, x);
            }
            long long e1 = 0, e2 = 0;
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
                if (dbg) e1 = clock64();
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
                if (dbg) e2 = clock64();
                // this thread's row of the SO(2) staging rows (null: the table is read from global memory)
                const float* so2_sm = nullptr;
                if (kSo2Stage && a.so2_stage)
                    so2_sm = reinterpret_cast<const float*>(smem + L::kSo2) + (warp * 32 + lane) * a.so2_stage;
                auto get_so2 = [&](int c) {
                    if (so2_sm) {
                        So2Chunk r_;
                        const float4* p_ = reinterpret_cast<const float4*>(so2_sm + (c - c_so2) * 8);
                        r_.a = p_[0]; r_.b = p_[1];
                        return r_;
                    }
                    return load_so2_chunk(so2, c, a.hd);
                };
                if (so2_sm && c_so2 < D / 8) {
                    cp_async_wait_all();
                    __syncwarp();
                }
                So2Chunk sc_cur = get_so2(c_so2);
#pragma unroll 1
                for (int c = c_so2; c < D / 8; ++c) {
                    So2Chunk sc_nxt = sc_cur;
                    if (c + 1 < D / 8) sc_nxt = get_so2(c + 1);
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
                d_e[0] += e0 - d_t2; d_e[1] += e1 - e0; d_e[2] += e2 - e1; d_e[3] += d_t3 - e2;
            }
        }
        if (dbg) {
            dbg[0] = clock64() - d_start; dbg[1] = d_loop; dbg[2] = d_epi; dbg[3] = d_wait_s; dbg[4] = d_wait_o;
            dbg[5] = d_items; dbg[6] = d_e[0]; dbg[7] = d_e[1]; dbg[13] = d_e[2]; dbg[14] = d_e[3];
        }
    } else {
      setmaxnreg_dec<96>();
      if (warp >= 10) {
        // =========================================================== Q stager (runs one item ahead)
        const int r0 = threadIdx.x - 320;    // 0..63; this thread stages rows r0 and r0 + 64 of each tile
        uint32_t cntx[2] = {0, 0};          // items staged per tile slot (buffer = cnt & 1, phase = (cnt >> 1) & 1)
#pragma unroll 1
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const ItemCoord ic = decode_item(item, npairs, a.H, a.Tq);
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
                    if constexpr (kQRowPrefetch && D <= 64) {
                        // Short key loops (CLEVR: 5 key tiles per item) leave the two stager warps less time per item than
                        // their four dependent load round trips take (q_full was 31 % of the issuer's time at the CLEVR decoder
                        // shape): pull the NEXT row this thread will stage -- its raw q line and its SO(2) table line -- into L1
                        // now.  Only where shared memory leaves L1 room (D <= 64).
                        int nb_ = ic.b, nh_ = ic.h, nt_ = -1;
                        if (rr == 0) nt_ = t + 64;
                        else if (X == 0 && ic.has_b) nt_ = ic.p * 256 + 128 + r0;
                        else if (item + static_cast<int>(gridDim.x) < nitems) {
                            const ItemCoord nx_ = decode_item(item + gridDim.x, npairs, a.H, a.Tq);
                            nb_ = nx_.b; nh_ = nx_.h; nt_ = nx_.p * 256 + r0;
                        }
                        if (nt_ >= 0 && nt_ < a.Tq) {
                            prefetch_l1(reinterpret_cast<const TIn*>(a.q) + static_cast<int64_t>(nb_) * a.q_sb +
                                        static_cast<int64_t>(nh_) * a.q_sh + static_cast<int64_t>(nt_) * a.q_st);
                            if (a.hd.so2) prefetch_l1(a.so2_q + (static_cast<size_t>(nb_) * a.Tq + nt_) * a.C * 2);
                        }
                    }
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
#if GTA_Q_EVICT_FIRST
                            if (valid) load_raw_stream(qrow + (g * G + i) * 8, raw[i]);
#else
                            if (valid) load_raw(qrow + (g * G + i) * 8, raw[i]);
#endif
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
                const ItemCoord ic = decode_item(item, npairs, a.H, a.Tq);
                const int nx = ic.has_b ? 2 : 1;
                uint32_t q_addr[2];
                for (int X = 0; X < nx; ++X) {
                    const uint32_t c_ = cntx[X];
                    const int buf = c_ & 1;
                    q_addr[X] = smem_u32(smem + L::kQ + (buf * 2 + X) * L::kTile);
                }

                auto issue_qk = [&](int X, int j) {
                    const int s = (gk + j) % NS;
                    if (elect_one()) {
                        const uint64_t qd = desc_kmajor_sw64(q_addr[X], 0);
                        const uint64_t kd = desc_kmajor_sw64(smem_u32(smem + L::kK + s * L::kTile), 0);
                        const uint32_t qlo = static_cast<uint32_t>(qd), qhi = static_cast<uint32_t>(qd >> 32);
                        const uint32_t klo = static_cast<uint32_t>(kd), khi = static_cast<uint32_t>(kd >> 32);
                        const uint32_t d_addr = tmem_base + (X ? k3TmemSB : k3TmemSA);
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
                    GTA_TIMED_WAIT(w_p, mbar_wait(&bars[L::bPFull + X], (gtx[X] + j) & 1));
                    if (j == 0 && cntx[X] > 0) GTA_TIMED_WAIT(w_of, mbar_wait(&bars[L::bOFree + X], (cntx[X] - 1) & 1));
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t vd = desc_mnmajor_sw64(smem_u32(smem + L::kV + s * L::kTile), 0);
                        const uint32_t vlo = static_cast<uint32_t>(vd), vhi = static_cast<uint32_t>(vd >> 32);
                        const uint32_t d_addr = tmem_base + (X ? k3TmemOB : k3TmemOA);
                        const uint32_t p_addr = tmem_base + (X ? k3TmemSB : k3TmemSA);
#pragma unroll
                        for (int kk = 0; kk < 8; ++kk)
                            umma_ts_lohi(d_addr, p_addr + kk * 8, vlo + kstep_mnmajor_sw64(kk), vhi, idesc_pv,
                                         (j > 0 || kk > 0) ? 1u : 0u);
                        if (X == nx - 1) umma_commit_addr(bar0 + (L::bVEmpty + s) * 8);
                        if (j == n - 1) umma_commit_addr(bar0 + (L::bOFinal + X) * 8);
                    }
                    __syncwarp();
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
                const ItemCoord ic = decode_item(item, npairs, a.H, a.Tq);
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
static int launch3_one(const AttnArgs& a_in, const GtaAttnParams& p, cudaStream_t st) {
    using L = Attn3Cfg<D>;
    auto kern = attn_fwd3_kernel<TIn, TOut, D>;
    // SO(2) staging rows behind the fixed layout when they fit the 227 KB of a CTA (MSN: 24 so2 dims -> 28 KB, CLEVR: 32 -> 36 KB)
    AttnArgs a = a_in;
    uint32_t smem_bytes = L::kBytes;
    if (kSo2Stage && p.so2 > 0 && p.v_transform && L::kSo2 + L::so2_stage_bytes(p.so2) + 1024u <= 227u * 1024u) {
        a.so2_stage = p.so2 + 4;
        const uint32_t need = L::kSo2 + L::so2_stage_bytes(p.so2) + 1024u;
        if (need > smem_bytes) smem_bytes = need;
    }
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(227u * 1024u));
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
    kern<<<grid, kThreads3, smem_bytes, st>>>(a, npairs, static_cast<int>(nitems));
    return check_launch("gta_attn_fwd");
}

template <typename TIn, typename TOut>
static int launch3_d(const AttnArgs& a, const GtaAttnParams& p, cudaStream_t st) {
    switch (p.D) {
        case 32: return launch3_one<TIn, TOut, 32>(a, p, st);
        case 64: return launch3_one<TIn, TOut, 64>(a, p, st);
        case 96: return launch3_one<TIn, TOut, 96>(a, p, st);
    }
    return set_error(GTA_ERR_UNSUPPORTED, "persistent pipeline supports head dims 32/64/96");
}

int launch_attn_fwd_v2(const GtaAttnParams& p, cudaStream_t st) {
    const AttnArgs a = make_attn_args(p);
    const bool ib = p.in_dtype == GTA_DTYPE_BF16, ob = p.out_dtype == GTA_DTYPE_BF16;
    if (ib && ob) return launch3_d<__nv_bfloat16, __nv_bfloat16>(a, p, st);
    if (ib && !ob) return launch3_d<__nv_bfloat16, float>(a, p, st);
    if (!ib && ob) return launch3_d<float, __nv_bfloat16>(a, p, st);
    return launch3_d<float, float>(a, p, st);
}

}  // namespace gta
