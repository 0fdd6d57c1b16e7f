// Fused GTA attention BACKWARD for sm_100a: dQ, dK, dV and d(trans_coeff) of
//
//   O = rho_q^{-1} softmax((rho_q^{-T} Q)(rho_k K)^T * scale) (rho_k V)        (source/utils/gta.py:92-279)
//
// The block-diagonal reps are constant linear maps, so with Q' = rho_q^{-T} Q, K' = rho_k K, V' = rho_k V, O = rho_q^{-1} O':
//   dO' = rho_q^{-T} dO (the same map the forward applies to Q)      delta = rowsum(dO * O) = rowsum(dO' * O')
//   P = exp(Q'K'^T scale - lse),  dV' = P^T dO',  dP = dO' V'^T,  dS = P * (dP - delta) * scale,
//   dQ' = dS K',  dK' = dS^T Q',        dQ = rho_q^{-1} dQ',  dK = rho_k^T dK',  dV = rho_k^T dV'.
// scale_mask(trans_coeff) scales one column of the SE(3) matrices (gta.py:40-44), so d(trans_coeff) is a sum of
// per-4-vector terms evaluated where the un-rotated gradients are still in registers.
//
// Launches: (1) staging — K'/V' tile images (gta_rotate_kv.cu), Q'/dO' tile images + delta (+ the output-side
// trans_coeff term); (2) attn_bwd_dkv_kernel: one CTA per (b, h, 128-key tile) loops over the query tiles with
// S^T = K'Q'^T and dP^T = V'dO'^T (SS MMAs), P^T / dS^T written back IN PLACE as bf16 and consumed from tensor memory by
// dV' += P^T dO', dK' += dS^T Q' (TS MMAs, the operand tile images read MN-major); (3) attn_bwd_dq_kernel: one CTA per
// (b, h, 128-query tile) loops over the key tiles with S, dP (SS) and dQ' += dS K' (TS).  Recomputing S and dP in both
// kernels (7 instead of 5 MMAs per tile pair) avoids atomics on dQ and keeps every accumulator in tensor memory.
// Head dims <= 96 pass P / dS through shared memory (K-major operand tiles) so that the next tile's S / dP MMAs overlap
// the SIMT phase of the current one; kNW compute warpgroups split the columns of every S / dP tile.  Measured (B200, MSN
// shape, tools/bwd_phase_timing.py): the SIMT phase (~2.2 k clk per tile in the dK/dV kernel) and the per-CTA epilogue
// (~13 k clk) bound the kernels, not the tensor pipe (1.6 k clk per tile); 4 warpgroups instead of 2 were slower
// (register pressure in the epilogue), so kNW = 2.
#include <cmath>

#include "gta_attn_bwd.cuh"

namespace gta {


#ifndef GTA_BWD_NW
#define GTA_BWD_NW 2
#endif
constexpr int kNW = GTA_BWD_NW;                 // compute warpgroups per CTA (2 or 4); each owns kCW columns of every S / dP tile
constexpr int kCW = 128 / kNW;
constexpr int kBwdThreads = kNW * 128 + 64;
constexpr uint32_t kBwdTmemS = 0, kBwdTmemDP = 128, kBwdTmemAcc0 = 256, kBwdTmemAcc1 = 384;

template <int D>
struct BwdSmem {
    static constexpr uint32_t kTile = 128u * D * 2u;
    static constexpr uint32_t kFix0 = 0, kFix1 = kTile;          // the CTA's own two tiles (K',V' or Q',dO')
    static constexpr uint32_t kStg0 = 2 * kTile;                  // [2 stages] streamed tile 0 (Q' / K')
    static constexpr uint32_t kStg1 = 4 * kTile;                  // [2 stages] streamed tile 1 (dO' / V')
    // head dims <= 96: P / dS go through SHARED memory (two 32 KB K-major operand tiles, 128-byte swizzle) so that the
    // S / dP accumulators are free the moment they are in registers and the next tile's S / dP MMAs overlap the SIMT
    // phase; head dim 128 has no room for that and keeps P / dS in tensor memory (in place).
    static constexpr bool kPS = D <= 96;
    static constexpr uint32_t kP = 6 * kTile;                     // P  [2 blocks of 64][128 rows][128 B]
    static constexpr uint32_t kDS = kP + (kPS ? 32768u : 0u);     // dS
    static constexpr uint32_t kLD = kDS + (kPS ? 32768u : 0u);    // float [2 warpgroups][2 buffers][lse*log2e 64 | delta 64]
    static constexpr uint32_t kBars = kLD + 2 * 2 * 128 * 4;      // kNW warpgroups x 2 buffers x 2 kCW floats = 512 floats
    enum : int { bFix = 0, bFull = 1, bEmpty = 3, bSFull = 5, bPReady = 6, bDone = 7, bSFree = 8, bPdFree = 9, bCount = 10 };
    static constexpr uint32_t kTmemSlot = kBars + bCount * 8;
    static constexpr uint32_t kUsed = kTmemSlot + 16;
    static constexpr uint32_t kBytes = (kUsed + 1024 > 120u * 1024u) ? kUsed + 1024 : 120u * 1024u;
};

// TMEM column of the packed bf16 A operand for K-step kk (16 rows of the streamed tile): the first compute warpgroup packs
// columns 0..63 into 0..31, the second 64..127 into 64..95.
__device__ __forceinline__ constexpr uint32_t pk_off(int kk) { return (kk / (kCW / 16)) * kCW + (kk % (kCW / 16)) * 8u; }

// Shared skeleton of the two kernels.  kDKV = true : fixed tiles K'_j, V'_j; streamed Q'_i, dO'_i; rows = keys.
//                                      kDKV = false: fixed tiles Q'_i, dO'_i; streamed K'_j, V'_j; rows = queries.
template <typename TIn, typename TOut, int D, bool kDKV>
__global__ void __launch_bounds__(kBwdThreads, 1) attn_bwd_kernel(const BwdArgs a) {
    using L = BwdSmem<D>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBars);
    float* sLD = reinterpret_cast<float*>(smem + L::kLD);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kTmemSlot);

    const int tile = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nstream = kDKV ? a.ntq : a.ntk;
    const size_t bh = static_cast<size_t>(b) * a.H + h;
    const uint8_t* fix0 = (kDKV ? a.k_img + (bh * a.ntk + tile) * L::kTile : a.q_img + (bh * a.ntq + tile) * L::kTile);
    const uint8_t* fix1 = (kDKV ? a.v_img + (bh * a.ntk + tile) * L::kTile : a.do_img + (bh * a.ntq + tile) * L::kTile);
    const uint8_t* stg0 = (kDKV ? a.q_img + bh * a.ntq * L::kTile : a.k_img + bh * a.ntk * L::kTile);
    const uint8_t* stg1 = (kDKV ? a.do_img + bh * a.ntq * L::kTile : a.v_img + bh * a.ntk * L::kTile);

    if (threadIdx.x == 0) {
        mbar_init(&bars[L::bFix], 1);
        for (int s = 0; s < 2; ++s) { mbar_init(&bars[L::bFull + s], 1); mbar_init(&bars[L::bEmpty + s], 1); }
        mbar_init(&bars[L::bSFull], 1);
        mbar_init(&bars[L::bPReady], kNW * 128);
        mbar_init(&bars[L::bDone], 1);
        mbar_init(&bars[L::bSFree], kNW * 128);
        mbar_init(&bars[L::bPdFree], 1);
        fence_mbar_init();
    }
    if (warp == kNW * 4) {
        tmem_alloc(tmem_slot, kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

    if (warp < kNW * 4) {
        // =========================================================== kNW compute warpgroups: thread r <-> row r <-> TMEM
        // lane r in both; warpgroup w owns columns [64w, 64w+64) of every S / dP tile (no exchange is needed: lse and
        // delta are known), which halves the latency of the SIMT phase between the two MMA phases of a tile.
        const int wgc = warp >> 2;
        const int r = threadIdx.x & 127;
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
        const float tc = a.tc_ptr ? __ldg(a.tc_ptr) : 1.0f;
        const float cs = a.scale_log2;
        constexpr float kLog2e = 1.4426950408889634f;
        const float* lse_bh = a.lse + bh * a.Tq;
        const float* del_bh = a.delta + bh * a.Tq;
        float my_lse2 = 0.f, my_del = 0.f;                   // dQ kernel: this row's statistics
        // dKV kernel: per-column statistics of this warpgroup's kCW query columns, [buffer][lse*log2e | delta*scale]
        float* sLDw = sLD + wgc * (4 * kCW);
        // thread r fetches one of the 2*kCW values of query tile i.  The RAW value is returned and scaled only when it
        // is stored (col_scale): an arithmetic use right after the load would stall the in-order issue for a full
        // global-memory latency at the top of every tile (measured: 1.5 k clk per tile).
        auto col_stat = [&](int i) {
            const int t = i * 128 + wgc * kCW + (r % kCW);
            if (t >= a.Tq || r >= 2 * kCW) return 0.f;
            return r < kCW ? lse_bh[t] : del_bh[t];
        };
        const float col_scale = r < kCW ? kLog2e : a.scale;  // lse in log2 units; delta pre-scaled: dS = P * (dP*scale - delta*scale)
        if (!kDKV) {
            const int t = tile * 128 + r;
            if (t < a.Tq) { my_lse2 = lse_bh[t] * kLog2e; my_del = del_bh[t]; }
        } else {
            if (r < 2 * kCW) sLDw[r] = col_stat(0) * col_scale;
            bwd_bar_sync(1 + wgc);
        }
        // packed results go to the first 32 columns of the warpgroup's OWN 64-column range (columns 0..31 / 64..95):
        // never into columns the other warpgroup still has to read
        const uint32_t pk_col = wgc * kCW;
        const size_t cta = (static_cast<size_t>(blockIdx.z) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        long long* dbg = (a.dbg && threadIdx.x == 0) ? a.dbg + ((kDKV ? 0 : static_cast<size_t>(gridDim.x) * gridDim.y * gridDim.z) + cta) * 16 : nullptr;
        long long d_wait = 0, d_ld = 0, d_cmp = 0, d_st = 0, d_tail = 0;
        const long long d_start = dbg ? clock64() : 0;
#pragma unroll 1
        for (int i = 0; i < nstream; ++i) {
            float nx_stat = 0.f;
            if (kDKV && i + 1 < nstream) nx_stat = col_stat(i + 1);   // in flight during this tile
            if (i == nstream - 1) {
                // pull this row's epilogue operands (view matrices, raw se3 elements) into L1 while the last tile is computed
                // (not the per-row SO(2) tables: 3 lines per row x 256 rows would evict everything else from the ~30 KB of L1)
                const int T_ = kDKV ? a.Tk : a.Tq;
                const int tt_ = min(tile * 128 + r, T_ - 1);
                const size_t view_ = static_cast<size_t>(b) * (kDKV ? a.Nk : a.Nq) + tt_ / (kDKV ? a.tpvk : a.tpvq);
                if (a.hd.se3) prefetch_l1((kDKV ? a.se3_k : a.se3_q) + view_ * 16);
                if (a.hd.so3) {
                    prefetch_l1((kDKV ? a.so3_k : a.so3_q) + view_ * 34);
                    prefetch_l1((kDKV ? a.so3_k : a.so3_q) + view_ * 34 + 32);
                }
                if (a.dtc && a.hd.se3) {                     // ... and the raw se3 elements the trans_coeff term reads
                    const TIn* raw_ = kDKV
                        ? (wgc < kNW / 2 ? reinterpret_cast<const TIn*>(a.v) + b * a.v_sb + h * a.v_sh + static_cast<int64_t>(tt_) * a.v_st
                                    : reinterpret_cast<const TIn*>(a.k) + b * a.k_sb + h * a.k_sh + static_cast<int64_t>(tt_) * a.k_st)
                        : reinterpret_cast<const TIn*>(a.q) + b * a.q_sb + h * a.q_sh + static_cast<int64_t>(tt_) * a.q_st;
                    for (int e = a.hd.triv; e < a.hd.triv + a.hd.se3; e += 128 / static_cast<int>(sizeof(TIn)))
                        prefetch_l1(raw_ + e);
                    prefetch_l1(raw_ + a.hd.triv + a.hd.se3 - 1);
                }
            }
            const long long d0 = dbg ? clock64() : 0;
            mbar_wait(&bars[L::bSFull], i & 1);
            tc_fence_after();
            const long long d1 = dbg ? clock64() : 0;
            const float* ld = sLDw + (i & 1) * (2 * kCW);
            const int ncol = (kDKV ? a.Tq : a.Tk) - i * 128 - wgc * kCW;   // valid columns of this warpgroup's half
            uint32_t sr[kCW], dr[kCW];
#pragma unroll
            for (int q32 = 0; q32 < kCW / 32; ++q32) {
                tmem_ld32(lane_base + kBwdTmemS + wgc * kCW + q32 * 32, sr + q32 * 32);
                tmem_ld32(lane_base + kBwdTmemDP + wgc * kCW + q32 * 32, dr + q32 * 32);
            }
            tmem_ld_wait();
            if (L::kPS) {                                    // S / dP are in registers: the next tile's MMAs may overwrite them
                tc_fence_before();
                mbar_arrive(&bars[L::bSFree]);
            }
            const long long d2 = dbg ? clock64() : 0;
            // dS = P * (dP - delta) * scale = P * (dP*scale - delta*scale): one FFMA + one FMUL per element
            const float my_dls = my_del * a.scale;
#pragma unroll
            for (int u4 = 0; u4 < kCW / 4; ++u4) {           // 4 columns per step (one 16-byte read of each statistic)
                float l2[4], dls[4];
                if (kDKV) {
                    const float4 a4 = *reinterpret_cast<const float4*>(ld + 4 * u4);
                    const float4 b4 = *reinterpret_cast<const float4*>(ld + kCW + 4 * u4);
                    l2[0] = a4.x; l2[1] = a4.y; l2[2] = a4.z; l2[3] = a4.w;
                    dls[0] = b4.x; dls[1] = b4.y; dls[2] = b4.z; dls[3] = b4.w;
                } else {
#pragma unroll
                    for (int w = 0; w < 4; ++w) { l2[w] = my_lse2; dls[w] = my_dls; }
                }
                float pv[4], dv[4];
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const int col = 4 * u4 + w;
                    pv[w] = fast_exp2(fmaf(__uint_as_float(sr[col]), cs, -l2[w]));
                    dv[w] = pv[w] * fmaf(__uint_as_float(dr[col]), a.scale, -dls[w]);
                }
                // in place: pairs (4u4, 4u4+1), (4u4+2, 4u4+3) land in slots 2u4, 2u4+1 <= the columns just consumed
                sr[2 * u4] = pack_bf16x2(pv[0], pv[1]); sr[2 * u4 + 1] = pack_bf16x2(pv[2], pv[3]);
                dr[2 * u4] = pack_bf16x2(dv[0], dv[1]); dr[2 * u4 + 1] = pack_bf16x2(dv[2], dv[3]);
            }
            if (ncol < kCW) {                                // ragged last tile: zero the packed columns past the end
#pragma unroll
                for (int u = 0; u < kCW / 2; ++u) {
                    if (2 * u + 1 >= ncol) {
                        const uint32_t keep = (2 * u < ncol) ? 0x0000FFFFu : 0u;
                        sr[u] &= keep; dr[u] &= keep;
                    }
                }
            }
            const long long d3 = dbg ? clock64() : 0;
            if (L::kPS) {
                // the previous tile's dV/dK (dQ) MMAs must be done reading the shared P / dS tiles
                if (i > 0) mbar_wait(&bars[L::bPdFree], (i - 1) & 1);
#pragma unroll
                for (int u = 0; u < kCW / 8; ++u) {          // chunks of 8 streamed rows (K elements)
                    const uint32_t off = tile_sw128_offset(r, wgc * (kCW / 8) + u);
                    if (kDKV) *reinterpret_cast<uint4*>(smem + L::kP + off) = make_uint4(sr[4 * u], sr[4 * u + 1], sr[4 * u + 2], sr[4 * u + 3]);
                    *reinterpret_cast<uint4*>(smem + L::kDS + off) = make_uint4(dr[4 * u], dr[4 * u + 1], dr[4 * u + 2], dr[4 * u + 3]);
                }
                fence_proxy_async_smem();
            } else {
                if (kCW == 64) {
                    if (kDKV) tmem_st32(lane_base + kBwdTmemS + pk_col, sr);
                    tmem_st32(lane_base + (kDKV ? kBwdTmemDP : kBwdTmemS) + pk_col, dr);
                } else {
                    if (kDKV) tmem_st16(lane_base + kBwdTmemS + pk_col, sr);
                    tmem_st16(lane_base + (kDKV ? kBwdTmemDP : kBwdTmemS) + pk_col, dr);
                }
                tmem_st_wait();
                tc_fence_before();
            }
            mbar_arrive(&bars[L::bPReady]);
            if (dbg) { const long long d4 = clock64(); d_wait += d1 - d0; d_ld += d2 - d1; d_cmp += d3 - d2; d_st += d4 - d3; }
            const long long d5 = dbg ? clock64() : 0;
            if (kDKV && i + 1 < nstream) {
                asm volatile("" : "+f"(nx_stat));            // first use of the loaded value HERE (keeps ptxas from scaling it early)
                if (r < 2 * kCW) sLDw[((i + 1) & 1) * (2 * kCW) + r] = nx_stat * col_scale;
                bwd_bar_sync(1 + wgc);
            }
            if (dbg) d_tail += clock64() - d5;
        }

        // ---- epilogue: accumulators -> registers (8 columns at a time) -> transposed / inverse rep -> global
        const long long d_loop_end = dbg ? clock64() : 0;
        mbar_wait(&bars[L::bDone], 0);
        tc_fence_after();
        const long long d_done = dbg ? clock64() : 0;
        const int T = kDKV ? a.Tk : a.Tq;
        const int t = tile * 128 + r;
        const bool valid = t < T;
        const int tt = valid ? t : T - 1;
        const size_t view = static_cast<size_t>(b) * (kDKV ? a.Nk : a.Nq) + tt / (kDKV ? a.tpvk : a.tpvq);
        const float* se3 = (kDKV ? a.se3_k : a.se3_q) + view * 16;
        const float* so3 = (kDKV ? a.so3_k : a.so3_q) + view * 34;
        const float* so2 = (kDKV ? a.so2_k : a.so2_q) + (static_cast<size_t>(b) * T + tt) * a.C * 2;
        const int c_se3 = a.hd.triv >> 3, c_so3 = c_se3 + (a.hd.se3 >> 3);
        float dtc_part = 0.f;
        // dKV kernel: warpgroup 0 finishes dV' (accumulator 0), warpgroup 1 dK' (accumulator 1);
        // dQ kernel: the two warpgroups split the columns of dQ' (accumulator 0).
        // dKV kernel: the first half of the warpgroups finishes dV' (accumulator 0), the second half dK' (accumulator 1),
        // each warpgroup a contiguous share of the columns; dQ kernel: all warpgroups split the columns of dQ'.
        constexpr int kShare = kDKV ? kNW / 2 : kNW;                 // warpgroups per accumulator
        const int part = wgc % kShare;
        const int c_lo = part * (D / 8) / kShare, c_hi = (part + 1) * (D / 8) / kShare;
        {
            const int which = kDKV ? wgc / kShare : 0;
            const uint32_t acc = lane_base + (which == 0 ? kBwdTmemAcc0 : kBwdTmemAcc1);
            const bool rotate = kDKV ? (which == 1 || a.v_transform) : true;
            const TIn* raw = kDKV
                ? (which == 0 ? reinterpret_cast<const TIn*>(a.v) + b * a.v_sb + h * a.v_sh + static_cast<int64_t>(tt) * a.v_st
                              : reinterpret_cast<const TIn*>(a.k) + b * a.k_sb + h * a.k_sh + static_cast<int64_t>(tt) * a.k_st)
                : reinterpret_cast<const TIn*>(a.q) + b * a.q_sb + h * a.q_sh + static_cast<int64_t>(tt) * a.q_st;
            constexpr uint32_t kPitch = D * sizeof(TOut) + (sizeof(TOut) == 2 ? 16 : 0);   // +16: conflict-free 16-byte row writes
            uint8_t* stage_base = smem + ((kDKV && which == 1) ? L::kStg1 : L::kStg0);
            uint8_t* stage_row = stage_base + static_cast<size_t>(r) * kPitch;
            const bool want_tc = rotate && valid && a.dtc != nullptr && a.hd.se3 > 0;
            // the row's view matrices once, in registers (the 128 score registers of the main loop are dead here): one
            // L2 round trip per row instead of one per chunk
            ViewReps vr;
            if (rotate || want_tc) load_view_reps(vr, a.hd, se3, so3);
            const float* M = vr.M;
            // one chunk ahead: the accumulator columns (tcgen05.ld is asynchronous until wait::ld) and, for the
            // trans_coeff term, the raw input chunk (a global load whose latency would otherwise be paid per chunk)
            uint32_t o8[8];
            constexpr int kAhead = 4;                        // raw chunks in flight: a global round trip is ~3 iterations
            RawChunk<TIn> ring[kAhead];
            auto want_raw = [&](int c) { return want_tc && c >= c_se3 && c < c_so3 && c < c_hi; };
#pragma unroll
            for (int u = 0; u < kAhead; ++u) {
                zero_raw(ring[u]);
                if (want_raw(c_lo + u)) load_raw(raw + (c_lo + u) * 8, ring[u]);
            }
            tmem_ld8(acc + c_lo * 8, o8);
            long long e_wait = 0, e_cmp = 0;
#pragma unroll 1
            for (int c = c_lo; c < c_hi; ++c) {
                const long long t0_ = dbg ? clock64() : 0;
                tmem_ld_wait();
                const long long t1_ = dbg ? clock64() : 0;
                float x[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) x[u] = __uint_as_float(o8[u]);
                const RawChunk<TIn> raw_cur = ring[0];
#pragma unroll
                for (int u = 0; u + 1 < kAhead; ++u) ring[u] = ring[u + 1];
                if (want_raw(c + kAhead)) load_raw(raw + (c + kAhead) * 8, ring[kAhead - 1]);
                const So2Chunk sc = rotate ? load_so2_chunk(so2, c, a.hd) : So2Chunk{};
                if (c + 1 < c_hi) tmem_ld8(acc + (c + 1) * 8, o8);
                if (want_tc && c >= c_se3 && c < c_so3) {
                    // d(trans_coeff): the un-rotated gradient times d(rep)/d(tc) applied to the raw input 4-vectors
                    float xin[8];
                    raw_to_f32(raw_cur, xin);
#pragma unroll
                    for (int v4 = 0; v4 < 2; ++v4) {
                        const float* g = x + 4 * v4;
                        const float* y = xin + 4 * v4;
                        if (kDKV) dtc_part += (g[0] * M[3] + g[1] * M[7] + g[2] * M[11]) * y[3];       // K', V': rows 0..2, column 3
                        else dtc_part += g[3] * (M[3] * y[0] + M[7] * y[1] + M[11] * y[2]);              // Q': transposed rep
                    }
                }
                if (rotate) {
                    if (kDKV) apply_rep_chunk_pre<kModeKVT>(x, c, a.hd, vr, sc, tc);
                    else apply_rep_chunk_pre<kModeOut>(x, c, a.hd, vr, sc, tc);
                }
                // rows are staged in shared memory (the streamed-tile buffers are idle now) and written out below with
                // consecutive threads on consecutive 16-byte pieces of a row: a thread-per-row store is one 16-byte
                // transaction per lane and instruction, which made this epilogue 25 % of the kernel.
                store_chunk<TOut>(reinterpret_cast<TOut*>(stage_row) + c * 8, x);
                if (dbg) { e_wait += t1_ - t0_; e_cmp += clock64() - t1_; }
            }
            const long long t2_ = dbg ? clock64() : 0;
            // hand the staged rows to the cooperative store
            asm volatile("bar.sync %0, %1;" ::"r"(3 + which), "r"(kShare * 128) : "memory");
            constexpr int kPieces = D * static_cast<int>(sizeof(TOut)) / 16;      // 16-byte pieces per row
            const int nthr = kShare * 128, tid = part * 128 + r;
            const int nrows = min(128, T - tile * 128);
            TOut* gbase = reinterpret_cast<TOut*>(kDKV ? (which == 0 ? a.dv : a.dk) : a.dq);
            for (int idx = tid; idx < nrows * kPieces; idx += nthr) {
                const int row = idx / kPieces, pc = idx - row * kPieces;
                const uint4 val = *reinterpret_cast<const uint4*>(stage_base + static_cast<size_t>(row) * kPitch + pc * 16);
                TOut* grow = gbase + ((static_cast<int64_t>(b) * T + tile * 128 + row) * a.H + h) * D;
                *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(grow) + pc * 16) = val;
            }
            if (dbg) { dbg[8] = e_wait; dbg[9] = e_cmp; dbg[10] = clock64() - t2_; dbg[11] = t2_ - d_done; dbg[12] = d_tail; dbg[13] = d_loop_end - d_start; }
        }
        if (a.dtc) {
            dtc_part = warp_sum(dtc_part);
            if (lane == 0 && dtc_part != 0.f) atomicAdd(a.dtc, dtc_part);
        }
        if (dbg) {
            const long long d_end = clock64();
            dbg[0] = d_end - d_start; dbg[1] = d_wait; dbg[2] = d_ld; dbg[3] = d_cmp; dbg[4] = d_st;
            dbg[5] = d_end - d_loop_end; dbg[6] = nstream; dbg[7] = d_done - d_loop_end;
        }
    } else if (warp == kNW * 4) {
        // =========================================================== UMMA issuer
        constexpr uint32_t idesc_ss = make_idesc_bf16(128, 128, 0, 0);
        constexpr uint32_t idesc_ts = make_idesc_bf16(128, D, 0, 1);
        const uint32_t f0 = smem_u32(smem + L::kFix0), f1 = smem_u32(smem + L::kFix1);
        mbar_wait(&bars[L::bFix], 0);
        const uint32_t p_sm = smem_u32(smem + L::kP), ds_sm = smem_u32(smem + L::kDS);
        // S / dP of streamed tile i:  dKV: S^T = K' Q'^T, dP^T = V' dO'^T;  dQ: S = Q' K'^T, dP = dO' V'^T  (A = fixed, B = streamed)
        auto issue_sdp = [&](int i) {
            const int s = i & 1;
            const uint32_t g0 = smem_u32(smem + L::kStg0 + s * L::kTile), g1 = smem_u32(smem + L::kStg1 + s * L::kTile);
            mbar_wait(&bars[L::bFull + s], (i >> 1) & 1);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < D / 16; ++kk)
                    umma_ss(tmem_base + kBwdTmemS, desc_kmajor_sw64(f0, kk), desc_kmajor_sw64(g0, kk), idesc_ss, kk > 0);
#pragma unroll
                for (int kk = 0; kk < D / 16; ++kk)
                    umma_ss(tmem_base + kBwdTmemDP, desc_kmajor_sw64(f1, kk), desc_kmajor_sw64(g1, kk), idesc_ss, kk > 0);
                umma_commit(&bars[L::bSFull]);
            }
            __syncwarp();
        };
        if (L::kPS) issue_sdp(0);
#pragma unroll 1
        for (int i = 0; i < nstream; ++i) {
            const int s = i & 1;
            const uint32_t g0 = smem_u32(smem + L::kStg0 + s * L::kTile), g1 = smem_u32(smem + L::kStg1 + s * L::kTile);
            if (L::kPS) {
                if (i + 1 < nstream) {                       // next tile's S / dP as soon as this tile's are in registers
                    mbar_wait(&bars[L::bSFree], i & 1);
                    issue_sdp(i + 1);
                }
            } else {
                issue_sdp(i);
            }
            mbar_wait(&bars[L::bPReady], i & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t accf = (i > 0) ? 1u : 0u;
                if (kDKV) {
                    // dV' += P^T dO'   (A = P^T, B = dO'_i read MN-major);  dK' += dS^T Q'
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {
                        if (L::kPS) umma_ss(tmem_base + kBwdTmemAcc0, desc_p_sw128(p_sm, kk), desc_mnmajor_sw64(g1, kk), idesc_ts, (kk > 0) ? 1u : accf);
                        else umma_ts(tmem_base + kBwdTmemAcc0, tmem_base + kBwdTmemS + pk_off(kk), desc_mnmajor_sw64(g1, kk), idesc_ts, (kk > 0) ? 1u : accf);
                    }
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {
                        if (L::kPS) umma_ss(tmem_base + kBwdTmemAcc1, desc_p_sw128(ds_sm, kk), desc_mnmajor_sw64(g0, kk), idesc_ts, (kk > 0) ? 1u : accf);
                        else umma_ts(tmem_base + kBwdTmemAcc1, tmem_base + kBwdTmemDP + pk_off(kk), desc_mnmajor_sw64(g0, kk), idesc_ts, (kk > 0) ? 1u : accf);
                    }
                } else {
                    // dQ' += dS K'     (A = dS, B = K'_j read MN-major)
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {
                        if (L::kPS) umma_ss(tmem_base + kBwdTmemAcc0, desc_p_sw128(ds_sm, kk), desc_mnmajor_sw64(g0, kk), idesc_ts, (kk > 0) ? 1u : accf);
                        else umma_ts(tmem_base + kBwdTmemAcc0, tmem_base + kBwdTmemS + pk_off(kk), desc_mnmajor_sw64(g0, kk), idesc_ts, (kk > 0) ? 1u : accf);
                    }
                }
                umma_commit(&bars[L::bEmpty + s]);
                if (L::kPS) umma_commit(&bars[L::bPdFree]);
                if (i == nstream - 1) umma_commit(&bars[L::bDone]);
            }
            __syncwarp();
        }
    } else {
        // =========================================================== bulk-copy producer
        if (lane == 0) {
            mbar_arrive_expect_tx(&bars[L::bFix], 2 * L::kTile);
            bulk_g2s(smem + L::kFix0, fix0, L::kTile, &bars[L::bFix]);
            bulk_g2s(smem + L::kFix1, fix1, L::kTile, &bars[L::bFix]);
        }
#pragma unroll 1
        for (int i = 0; i < nstream; ++i) {
            const int s = i & 1;
            if (i >= 2) mbar_wait(&bars[L::bEmpty + s], ((i >> 1) - 1) & 1);
            if (lane == 0) {
                mbar_arrive_expect_tx(&bars[L::bFull + s], 2 * L::kTile);
                bulk_g2s(smem + L::kStg0 + s * L::kTile, stg0 + static_cast<size_t>(i) * L::kTile, L::kTile, &bars[L::bFull + s]);
                bulk_g2s(smem + L::kStg1 + s * L::kTile, stg1 + static_cast<size_t>(i) * L::kTile, L::kTile, &bars[L::bFull + s]);
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kNW * 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ---------------------------------------------------------------------------------------------------- staging
// delta[b,h,t] = sum_d dO * O, and the output-side trans_coeff term: O_i = sum_j M_ij O'_j + tc * M_i3 * O'_3 (i < 3),
// O_3 = M_33 O'_3 with M = E_q  =>  d/dtc = sum_{i<3} dO_i M_i3 O_3 / M_33 per SE(3) 4-vector.
// 16 lanes per row (one 16-byte chunk each, D <= 128), two rows per warp: coalesced reads, shuffle reduction.
template <typename T>
__global__ void bwd_delta_kernel(const T* __restrict__ out, const T* __restrict__ dout, float* __restrict__ delta,
                                 float* dtc, const float* __restrict__ se3_q, int B, int Tq, int H, int D, int Nq, int tpvq,
                                 int triv, int se3, int v_transform) {
    // rows = (b, t, h) in memory order; 32-bit index arithmetic (the launcher guarantees B*Tq*H < 2^31): the 64-bit
    // divisions of a first version made this HBM pass instruction-bound (2.2 TB/s)
    const uint32_t total = static_cast<uint32_t>(B) * Tq * H;
    const uint32_t c = threadIdx.x & 15;
    const bool tc_lane = dtc && v_transform && se3 && static_cast<int>(c * 8) >= triv && static_cast<int>(c * 8) < triv + se3;
    float part = 0.f;
    // grid-stride over groups of 16 rows per block: the trans_coeff partial sums stay in registers and cost ONE atomic per
    // block at the end (one atomic per warp on a single address serialised in L2: +0.25 ms at the MSN shape)
    const uint32_t stride = (gridDim.x * blockDim.x) >> 4;
    for (uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) >> 4; row < ((total + 15u) & ~15u); row += stride) {
        float acc = 0.f;
        if (row < total && static_cast<int>(c * 8) < D) {
            float xo[8], xg[8];
            load_chunk<T>(out + static_cast<size_t>(row) * D + c * 8, xo);
            load_chunk<T>(dout + static_cast<size_t>(row) * D + c * 8, xg);
#pragma unroll
            for (int u = 0; u < 8; ++u) acc = fmaf(xo[u], xg[u], acc);
            if (tc_lane) {
                const uint32_t bt = row / H;
                const uint32_t b = bt / Tq, t = bt - b * Tq;
                const float* M = se3_q + (static_cast<size_t>(b) * Nq + t / tpvq) * 16;
#pragma unroll
                for (int v4 = 0; v4 < 2; ++v4)
                    part += (xg[4 * v4] * M[3] + xg[4 * v4 + 1] * M[7] + xg[4 * v4 + 2] * M[11]) * xo[4 * v4 + 3] / M[15];
            }
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (row < total && c == 0) {
            const uint32_t bt = row / H, h = row - bt * H;
            const uint32_t b = bt / Tq, t = bt - b * Tq;
            delta[(static_cast<size_t>(b) * H + h) * Tq + t] = acc;
        }
    }
    if (dtc) {
        __shared__ float red[8];
        part = warp_sum(part);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
        __syncthreads();
        if (threadIdx.x == 0) {
            float s = 0.f;
            for (int w = 0; w < 8; ++w) s += red[w];
            if (s != 0.f) atomicAdd(dtc, s);
        }
    }
}

// ---------------------------------------------------------------------------------------------------- host side
static size_t bwd_align(size_t v) { return (v + 1023) & ~static_cast<size_t>(1023); }

size_t attn_bwd_workspace_bytes(int B, int H, int Tq, int Tk, int D) {
    if (B <= 0 || H <= 0 || Tq <= 0 || Tk <= 0 || D <= 0) return 0;
    const size_t tile = kv_tile_bytes(D);
    const size_t ntq = num_kv_tiles(Tq), ntk = num_kv_tiles(Tk);
    return 2 * bwd_align(static_cast<size_t>(B) * H * ntq * tile) + bwd_align(2 * static_cast<size_t>(B) * H * ntk * tile) +
           bwd_align(static_cast<size_t>(B) * H * Tq * 4) + (bwd_fused_supported(D) ? bwd_align(bwd_dq_acc_bytes(B, H, Tq, D)) : 0);
}

template <typename TIn, typename TOut, int D>
static int launch_bwd_d(const BwdArgs& a, cudaStream_t st) {
    using L = BwdSmem<D>;
    auto kkv = attn_bwd_kernel<TIn, TOut, D, true>;
    auto kq = attn_bwd_kernel<TIn, TOut, D, false>;
    cudaError_t e = cudaFuncSetAttribute(kkv, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(L::kBytes));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kq, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(L::kBytes));
    if (e != cudaSuccess) return set_error(GTA_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    kkv<<<dim3(a.ntk, a.H, a.B), kBwdThreads, L::kBytes, st>>>(a);
    kq<<<dim3(a.ntq, a.H, a.B), kBwdThreads, L::kBytes, st>>>(a);
    return check_launch("gta_attn_bwd");
}

template <typename T>
static int launch_bwd_t(const BwdArgs& a, int D, cudaStream_t st) {
    switch (D) {
        case 32: return launch_bwd_d<T, T, 32>(a, st);
        case 64: return launch_bwd_d<T, T, 64>(a, st);
        case 96: return launch_bwd_d<T, T, 96>(a, st);
        case 128: return launch_bwd_d<T, T, 128>(a, st);
    }
    return set_error(GTA_ERR_UNSUPPORTED, "head dim %d", D);
}

int launch_attn_bwd(const GtaAttnBwdParams& bp, cudaStream_t st, const void* delta_dout) {
    const GtaAttnParams& p = bp.fwd;
    if (attn_needs_generic(p)) return launch_attn_bwd_generic(bp, st);   // t2 block / unaligned blocks (euclid_sim: unsupported)
    if (!delta_dout) delta_dout = bp.dout;
    if (p.out_dtype != p.in_dtype) return set_error(GTA_ERR_UNSUPPORTED, "gta_attn_bwd: out/dout must have the dtype of q/k/v");
    if (!p.lse || !bp.dout || !bp.dq || !bp.dk || !bp.dv) return set_error(GTA_ERR_INVALID, "gta_attn_bwd: null lse/dout/dq/dk/dv");
    const size_t need = attn_bwd_workspace_bytes(p.B, p.H, p.Tq, p.Tk, p.D);
    if (!bp.workspace || bp.workspace_bytes < need) return set_error(GTA_ERR_INVALID, "gta_attn_bwd: workspace too small (need %zu bytes)", need);
    if (reinterpret_cast<uintptr_t>(bp.workspace) & 1023) return set_error(GTA_ERR_INVALID, "gta_attn_bwd: workspace must be 1024-byte aligned");

    const size_t tile = kv_tile_bytes(p.D);
    const int ntq = num_kv_tiles(p.Tq), ntk = num_kv_tiles(p.Tk);
    uint8_t* ws = static_cast<uint8_t*>(bp.workspace);
    uint8_t* q_img = ws;
    uint8_t* do_img = q_img + bwd_align(static_cast<size_t>(p.B) * p.H * ntq * tile);
    uint8_t* kv_img = do_img + bwd_align(static_cast<size_t>(p.B) * p.H * ntq * tile);
    float* delta = reinterpret_cast<float*>(kv_img + bwd_align(2 * static_cast<size_t>(p.B) * p.H * ntk * tile));

    // K'/V' tile images: the forward's staging kernel on plain bf16 images
    GtaAttnParams kvp = p;
    kvp.workspace = kv_img;
    kvp.workspace_bytes = 2 * static_cast<size_t>(p.B) * p.H * ntk * tile;
    kvp.flags = GTA_FLAG_FAST_FP32;
    int rc = launch_rotate_kv(kvp, st);
    if (rc) return rc;

    const bool bf = p.in_dtype == GTA_DTYPE_BF16;
    rc = launch_rotate_q_do(p, bp.dout, q_img, do_img, st);      // Q' and dO' = rho_q^{-T} (q, dout): one launch, shared rep data
    if (rc) return rc;
    {
        const int64_t rows = static_cast<int64_t>(p.B) * p.Tq * p.H;
        if (rows >= (1LL << 31) - 16) return set_error(GTA_ERR_UNSUPPORTED, "gta_attn_bwd: B*Tq*H must be below 2^31");
        const int64_t want = (rows * 16 + 255) / 256;
        const unsigned nb = static_cast<unsigned>(want < 148 * 16 ? want : 148 * 16);
        float* dtc = p.se3 ? bp.dtrans_coeff : nullptr;
        if (bf) bwd_delta_kernel<__nv_bfloat16><<<nb, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(p.out), static_cast<const __nv_bfloat16*>(delta_dout),
                                                                    delta, dtc, p.reps.se3_q, p.B, p.Tq, p.H, p.D, p.Nq, p.Tq / p.Nq, p.triv, p.se3, p.v_transform);
        else bwd_delta_kernel<float><<<nb, 256, 0, st>>>(static_cast<const float*>(p.out), static_cast<const float*>(delta_dout), delta, dtc,
                                                         p.reps.se3_q, p.B, p.Tq, p.H, p.D, p.Nq, p.Tq / p.Nq, p.triv, p.se3, p.v_transform);
    }
    rc = check_launch("gta_attn_bwd (staging)");
    if (rc) return rc;

    BwdArgs a;
    a.q_img = q_img; a.do_img = do_img; a.k_img = kv_img; a.v_img = kv_img + static_cast<size_t>(p.B) * p.H * ntk * tile;
    a.lse = p.lse; a.delta = delta;
    a.dq = bp.dq; a.dk = bp.dk; a.dv = bp.dv;
    a.q = p.q; a.k = p.k; a.v = p.v;
    a.q_sb = p.q_stride_b; a.q_sh = p.q_stride_h; a.q_st = p.q_stride_t;
    a.k_sb = p.k_stride_b; a.k_sh = p.k_stride_h; a.k_st = p.k_stride_t;
    a.v_sb = p.v_stride_b; a.v_sh = p.v_stride_h; a.v_st = p.v_stride_t;
    a.dtc = p.se3 ? bp.dtrans_coeff : nullptr;
    a.B = p.B; a.H = p.H; a.Tq = p.Tq; a.Tk = p.Tk; a.Nq = p.Nq; a.Nk = p.Nk; a.tpvq = p.Tq / p.Nq; a.tpvk = p.Tk / p.Nk;
    a.ntq = ntq; a.ntk = ntk; a.C = p.so2 >> 1;
    a.hd = HeadDims{p.triv, p.se3, p.so3, p.so2};
    a.se3_q = p.reps.se3_q; a.so3_q = p.reps.so3_q; a.so2_q = p.reps.so2_q;
    a.se3_k = p.reps.se3_k; a.so3_k = p.reps.so3_k; a.so2_k = p.reps.so2_k;
    a.tc_ptr = p.trans_coeff;
    a.scale = p.scale; a.scale_log2 = p.scale * 1.4426950408889634f;
    a.v_transform = p.v_transform;
    a.dbg = p.debug_clocks;
    a.dq_acc = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(delta) + bwd_align(static_cast<size_t>(p.B) * p.H * p.Tq * 4));
    // head dims <= 96: ONE kernel for dK, dV and the dQ' partial sums (gta_attn_bwd2.cu); GTA_FLAG_BWD_SPLIT keeps the
    // dK/dV + dQ kernel pair (the only path for head dim 128)
    // (calls of less than one wave of key tiles are launch-bound: the pair's two launches beat memset + fused + finishing
    // kernel by ~6 % at BASELINE config 1; GTA_FLAG_SINGLE_LAUNCH forces the fused kernel there too)
    const bool tiny = static_cast<int64_t>(p.B) * p.H * ntk < 148 && !(p.flags & GTA_FLAG_SINGLE_LAUNCH);
    if (bwd_fused_supported(p.D) && !(p.flags & GTA_FLAG_BWD_SPLIT) && !tiny)
        return launch_bwd_fused(a, bf, p.D, (p.flags & GTA_FLAG_RUNTIME_LAYOUT) != 0, st);
    return bf ? launch_bwd_t<__nv_bfloat16>(a, p.D, st) : launch_bwd_t<float>(a, p.D, st);
}

}  // namespace gta
