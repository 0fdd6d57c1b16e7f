// Fused GTA attention backward for head dims <= 96 (maths in gta_attn_bwd.cu / source/utils/gta.py:92-279):
// ONE kernel computes dK', dV' (accumulated in tensor memory over the query tiles) AND the dQ' partial sums, so that S and
// dP are computed once per (key tile, query tile) pair — 5 MMAs per pair instead of the 7 of the dK/dV + dQ kernel pair.
//
//   CTA = (b, h, 128-key tile j), loops over the query tiles i (starting at i = j mod ntq, so that the CTAs of one (b,h)
//   touch different dQ' tiles at any time):
//     S^T = K'_j Q'_i^T, dP^T = V'_j dO'_i^T                 (SS MMAs, rows = keys)
//     P^T = exp2(S^T c - lse_i), dS^T = P^T (dP^T - delta_i) scale      (two compute warpgroups, thread = key row,
//                                                                        warpgroup w owns query columns [64w, 64w+64))
//     dV' += P^T dO'_i    (A = P^T as bf16 in tensor memory, B = dO'_i read MN-major)
//     dK' += dS^T Q'_i    (A = dS^T from shared memory, K-major, 128-byte swizzle)
//     dQ'_i(partial) = dS K'_j   (A = the SAME shared-memory tile read MN-major, B = K'_j read MN-major)
//   The dQ' partial is drained from tensor memory by a reducer warpgroup, in two halves through one half-size shared-memory
//   tile, and added to an fp32 accumulation buffer in global memory by bulk reductions (cp.reduce.async.bulk ... add.f32,
//   performed in L2); the
//   buffer's tile layout [D/4][128 rows][4 floats] is chosen so that both the staging writes and the finishing kernel's
//   reads are conflict-free / coalesced.  bwd_dq_finish_kernel then applies rho_q^{-1} (and the query-side trans_coeff
//   term) and writes dq.
//
// Tensor memory (512 columns): S^T 0..127 | dP^T 128..255, reused by dQ' (128..128+D) | dV' 256.. | dK' 352.. | P^T (bf16) 448..511.
// dQ' shares its columns with dP^T: dP^T(i) is dead once dS(i) is computed, and dP^T(i+1) is issued when the reducer has
// drained dQ'(i) — which happens while dV'(i) executes.  Issue order: S^T(i+1) as soon as S^T(i) is in registers; then, when P^T / dS^T
// of pair i are ready, dQ'(i), dV'(i), dP^T(i+1), dK'(i).
// Shared memory (D = 96, 226 KB): K', V' 48 KB | Q' ring of 3 72 KB | dO' ring of 2 48 KB | dQ' staging 24 KB | dS^T 32 KB |
// per-column statistics 2 KB.  The shared-memory pipe is the busiest unit of the kernel.
//
// 512 threads: warps 0-7 compute (setmaxnreg 184), warps 8-11 reducer (88), warp 12 UMMA issuer, warp 13 bulk-copy
// producer, warps 14-15 only complete the warpgroup (56).
#include "gta_attn_bwd.cuh"

namespace gta {

// Share of the exponentials evaluated with poly_exp2x2 instead of MUFU.EX2: groups of 4 columns with (u4 % DEN) < NUM.
#ifndef GTA_BWD2_POLY_NUM
#define GTA_BWD2_POLY_NUM 0
#endif
#ifndef GTA_BWD2_POLY_DEN
#define GTA_BWD2_POLY_DEN 4
#endif
constexpr int kFThreads = 512;
constexpr uint32_t kFTmemS = 0, kFTmemDP = 128, kFTmemDV = 256, kFTmemDK = 352, kFTmemP = 448;
constexpr uint32_t kSmemOptinMax = 232448u;   // 227 KB

template <int D>
struct FusedSmem {
    static constexpr int kStagesQ = 3, kStagesO = 2;
    static constexpr uint32_t kTile = 128u * D * 2u;
    static constexpr uint32_t kFix0 = 0, kFix1 = kTile;          // K'_j, V'_j
    // Q'_i is read by S^T(i) — issued a whole pair early — and last by dK'(i): three stages; dO'_i by dP^T(i) and last by
    // dV'(i), which completes early in its pair: two stages are enough
    static constexpr uint32_t kStgQ = 2 * kTile;                  // [3 stages] Q'_i
    static constexpr uint32_t kStgO = 5 * kTile;                  // [2 stages] dO'_i
    static constexpr uint32_t kDQ = 7 * kTile;                    // fp32 staging of HALF a dQ' partial [D/8][128 rows][16 B] (= kTile bytes)
    static constexpr uint32_t kDS = 8 * kTile;                    // dS^T [2 blocks of 64 queries][128 keys][128 B], 128-byte swizzle
    static constexpr uint32_t kLD = kDS + 32768u;                 // float [2 warpgroups][2 buffers][-lse*log2e 64 | -delta*scale 64]
    static constexpr uint32_t kBars = kLD + 2048u;
    enum : int { bFix = 0, bFullQ = 1, bEmptyQ = 4, bFullO = 7, bEmptyO = 9, bSFull = 11, bDPFull = 12, bPReady = 13, bPdFree = 14,
                 bDQFull = 15, bDQFree = 16, bDone = 17, bSFree = 18, bPFree = 19, bCount = 20 };
    static constexpr uint32_t kTmemSlot = kBars + bCount * 8;
    static constexpr uint32_t kUsed = kTmemSlot + 16;
    // the dynamic shared-memory window is 1024-byte aligned in practice; the kernel checks (and traps) if the slack is not enough
    static constexpr uint32_t kBytes = (kUsed + 1024u > kSmemOptinMax) ? kSmemOptinMax : kUsed + 1024u;
    static_assert(kUsed <= kSmemOptinMax, "shared-memory layout exceeds the 227 KB opt-in limit");
    static_assert(kDS % 1024u == 0, "128-byte-swizzled tile needs 1024-byte alignment");
};

// dst[i] += src[i] for `bytes` contiguous bytes of fp32, shared -> global, performed by the L2 (SASS: UBLKRED).  Measured against
// vector reductions from registers (red.global.add.v4.f32, 24 per thread and pair): those cost the pair loop +1.3 k clk.
__device__ __forceinline__ void bulk_reduce_add_f32(float* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// dS^T tile read MN-major (A operand of dQ' = dS K'): M = queries, contiguous in 64-element (128 B) swizzle atoms that are
// 16 KB apart (LBO); the 8-key groups along K are 1024 B apart (SBO); K-step kk = 16 keys = 2048 B.
__device__ __forceinline__ uint64_t desc_ds_mnmajor_sw128(uint32_t addr, int kk) {
    return make_smem_desc(addr + kk * 2048u, 16384u, 1024u, kLayoutSW128);
}

// LY: HeadLayout<...> (straight-line epilogue for a shipped head layout) or void (run-time layout from BwdArgs.hd).
template <typename TIn, typename TOut, int D, typename LY>
__global__ void __launch_bounds__(kFThreads, 1) attn_bwd_fused_kernel(const BwdArgs a) {
    using L = FusedSmem<D>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
    uint8_t* smem = smem_raw + pad;
    if (pad + L::kUsed > L::kBytes) {
        if (threadIdx.x == 0) printf("gta_b200: attn_bwd_fused_kernel: dynamic shared memory misaligned by %u bytes\n", pad);
        __trap();
    }
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBars);
    float* sLD = reinterpret_cast<float*>(smem + L::kLD);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kTmemSlot);

    const int tile = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int N = a.ntq;                                  // streamed (query) tiles
    const int i_start = tile % N;
    auto qtile = [&](int n) { const int i = i_start + n; return i >= N ? i - N : i; };
    const size_t bh = static_cast<size_t>(b) * a.H + h;
    const uint8_t* fix0 = a.k_img + (bh * a.ntk + tile) * L::kTile;
    const uint8_t* fix1 = a.v_img + (bh * a.ntk + tile) * L::kTile;
    const uint8_t* stg0 = a.q_img + bh * a.ntq * L::kTile;
    const uint8_t* stg1 = a.do_img + bh * a.ntq * L::kTile;

    if (threadIdx.x == 0) {
        mbar_init(&bars[L::bFix], 1);
        for (int s = 0; s < L::kStagesQ; ++s) { mbar_init(&bars[L::bFullQ + s], 1); mbar_init(&bars[L::bEmptyQ + s], 1); }
        for (int s = 0; s < L::kStagesO; ++s) { mbar_init(&bars[L::bFullO + s], 1); mbar_init(&bars[L::bEmptyO + s], 1); }
        mbar_init(&bars[L::bSFull], 1);
        mbar_init(&bars[L::bDPFull], 1);
        mbar_init(&bars[L::bPReady], 256);
        mbar_init(&bars[L::bPdFree], 1);
        mbar_init(&bars[L::bDQFull], 1);
        mbar_init(&bars[L::bDQFree], 128);
        mbar_init(&bars[L::bDone], 1);
        mbar_init(&bars[L::bSFree], 256);
        mbar_init(&bars[L::bPFree], 1);
        fence_mbar_init();
    }
    if (warp == 12) {
        tmem_alloc(tmem_slot, kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

    if (warp < 8) {
        // =========================================================== two compute warpgroups: thread r <-> key row r <-> TMEM lane r
        setmaxnreg_inc<184>();
        const int wgc = warp >> 2;
        const int r = threadIdx.x & 127;
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
        const float tc = a.tc_ptr ? __ldg(a.tc_ptr) : 1.0f;
        const float cs = a.scale_log2;
        constexpr float kLog2e = 1.4426950408889634f;
        const float* lse_bh = a.lse + bh * a.Tq;
        const float* del_bh = a.delta + bh * a.Tq;
        float* sLDw = sLD + wgc * 256;                        // [2 buffers][lse*log2e 64 | delta*scale 64]
        // thread r < 128 fetches one of the 128 statistics of query tile i (raw; scaled where it is stored, see gta_attn_bwd.cu)
        auto col_stat = [&](int i) {
            const int t = i * 128 + wgc * 64 + (r & 63);
            if (t >= a.Tq) return 0.f;
            return r < 64 ? lse_bh[t] : del_bh[t];
        };
        const float col_scale = r < 64 ? -kLog2e : -a.scale;   // stored NEGATED: both uses are fused multiply-adds
        const uint64_t cs2 = pack_f32x2(cs, cs), sc2 = pack_f32x2(a.scale, a.scale);
        sLDw[r] = col_stat(qtile(0)) * col_scale;
        bwd_bar_sync(1 + wgc);
        long long* dbg = (a.dbg && threadIdx.x == 0)
            ? a.dbg + ((static_cast<size_t>(blockIdx.z) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 16 : nullptr;
        long long d_ws = 0, d_lds = 0, d_exp = 0, d_wpd = 0, d_wdp = 0, d_ds = 0, d_tail = 0;
        const long long d_start = dbg ? clock64() : 0;
#pragma unroll 1
        for (int n = 0; n < N; ++n) {
            const int qi = qtile(n);
            float nx_stat = 0.f;
            if (n + 1 < N) nx_stat = col_stat(qtile(n + 1));  // in flight during this tile
            const float* ld = sLDw + (n & 1) * 128;
            const int ncol = a.Tq - qi * 128 - wgc * 64;      // valid query columns of this warpgroup's half
            if (n == (N > 1 ? N - 2 : 0)) {
                // the epilogue's global operands (this key row's raw K or V SE(3) block for the trans_coeff term, its angles)
                // were last touched by the staging kernels: pull them into L2 now, a DRAM round trip costs the epilogue 3-5 k clk
                const int tt_ = min(tile * 128 + r, a.Tk - 1);
                if (a.hd.so2) {
                    const float* p_ = a.so2_k + (static_cast<size_t>(b) * a.Tk + tt_) * a.C * 2;
                    prefetch_l2(p_);
                    prefetch_l2(p_ + a.C * 2 - 1);
                }
                if (a.dtc && a.hd.se3) {
                    const TIn* raw_ = (wgc == 0 ? reinterpret_cast<const TIn*>(a.v) + b * a.v_sb + h * a.v_sh + static_cast<int64_t>(tt_) * a.v_st
                                                : reinterpret_cast<const TIn*>(a.k) + b * a.k_sb + h * a.k_sh + static_cast<int64_t>(tt_) * a.k_st) + a.hd.triv;
                    prefetch_l2(raw_);
                    prefetch_l2(raw_ + a.hd.se3 - 1);
                }
            }
            const long long d0 = dbg ? clock64() : 0;
            mbar_wait(&bars[L::bSFull], n & 1);
            tc_fence_after();
            const long long d1 = dbg ? clock64() : 0;
            uint32_t sr[64], pk[32];
            tmem_ld32(lane_base + kFTmemS + wgc * 64, sr);
            tmem_ld32(lane_base + kFTmemS + wgc * 64 + 32, sr + 32);
            tmem_ld_wait();
            const long long d1b = dbg ? clock64() : 0;
            tc_fence_before();
            mbar_arrive(&bars[L::bSFree]);                    // S^T is in registers: the next pair's S^T may overwrite it
            // P^T = exp2(S^T c - lse*log2e): kept as fp32 in sr (for dS), packed to bf16 in pk (A operand of dV')
#pragma unroll
            for (int u4 = 0; u4 < 16; ++u4) {
#ifdef GTA_BWD2_NOLDS
                const ulonglong2 l4 = make_ulonglong2(0ull, 0ull);
#else
                const ulonglong2 l4 = *reinterpret_cast<const ulonglong2*>(ld + 4 * u4);   // -lse*log2e of 4 query columns
#endif
                const uint64_t x01 = ffma2(pack_f32x2(__uint_as_float(sr[4 * u4]), __uint_as_float(sr[4 * u4 + 1])), cs2, l4.x);
                const uint64_t x23 = ffma2(pack_f32x2(__uint_as_float(sr[4 * u4 + 2]), __uint_as_float(sr[4 * u4 + 3])), cs2, l4.y);
                float p0, p1, p2, p3;
                if ((u4 % GTA_BWD2_POLY_DEN) < GTA_BWD2_POLY_NUM) {
                    // a share of the exponentials on the FMA pipe (degree-3 polynomial, 7.5e-5 relative): the two compute warps of
                    // a scheduler need 1024 clk of MUFU per pair otherwise
                    poly_exp2x2(x01, p0, p1);
                    poly_exp2x2(x23, p2, p3);
                } else {
                    float x0, x1, x2, x3;
                    unpack_f32x2(x01, x0, x1);
                    unpack_f32x2(x23, x2, x3);
                    p0 = fast_exp2(x0); p1 = fast_exp2(x1); p2 = fast_exp2(x2); p3 = fast_exp2(x3);
                }
                sr[4 * u4] = __float_as_uint(p0); sr[4 * u4 + 1] = __float_as_uint(p1);
                sr[4 * u4 + 2] = __float_as_uint(p2); sr[4 * u4 + 3] = __float_as_uint(p3);
                pk[2 * u4] = pack_bf16x2(p0, p1); pk[2 * u4 + 1] = pack_bf16x2(p2, p3);
            }
            if (ncol < 64) {                                  // ragged last query tile: zero the columns past the end
#pragma unroll
                for (int u = 0; u < 32; ++u) {
                    if (2 * u + 1 >= ncol) pk[u] &= (2 * u < ncol) ? 0x0000FFFFu : 0u;
                }
            }
            const long long d2 = dbg ? clock64() : 0;
            // the previous pair's dV' MMAs must be done reading P^T (tensor memory)
            if (n > 0) mbar_wait(&bars[L::bPFree], (n - 1) & 1);
            tmem_st32(lane_base + kFTmemP + wgc * 32, pk);
            const long long d3 = dbg ? clock64() : 0;
            mbar_wait(&bars[L::bDPFull], n & 1);
            tc_fence_after();
            const long long d4 = dbg ? clock64() : 0;
            uint32_t dr[64];
            tmem_ld32(lane_base + kFTmemDP + wgc * 64, dr);
            tmem_ld32(lane_base + kFTmemDP + wgc * 64 + 32, dr + 32);
            tmem_ld_wait();
            // dS^T = P^T (dP^T scale - delta scale), packed in place: pair (2u, 2u+1) lands in slot u <= the columns just consumed
#pragma unroll
            for (int u4 = 0; u4 < 16; ++u4) {
#ifdef GTA_BWD2_NOLDS
                const ulonglong2 d4v = make_ulonglong2(0ull, 0ull);
#else
                const ulonglong2 d4v = *reinterpret_cast<const ulonglong2*>(ld + 64 + 4 * u4);   // -delta*scale of 4 query columns
#endif
                const uint64_t t01 = ffma2(pack_f32x2(__uint_as_float(dr[4 * u4]), __uint_as_float(dr[4 * u4 + 1])), sc2, d4v.x);
                const uint64_t t23 = ffma2(pack_f32x2(__uint_as_float(dr[4 * u4 + 2]), __uint_as_float(dr[4 * u4 + 3])), sc2, d4v.y);
                const uint64_t s01 = fmul2(pack_f32x2(__uint_as_float(sr[4 * u4]), __uint_as_float(sr[4 * u4 + 1])), t01);
                const uint64_t s23 = fmul2(pack_f32x2(__uint_as_float(sr[4 * u4 + 2]), __uint_as_float(sr[4 * u4 + 3])), t23);
                float s0, s1, s2, s3;
                unpack_f32x2(s01, s0, s1);
                unpack_f32x2(s23, s2, s3);
                dr[2 * u4] = pack_bf16x2(s0, s1); dr[2 * u4 + 1] = pack_bf16x2(s2, s3);
            }
            if (ncol < 64) {
#pragma unroll
                for (int u = 0; u < 32; ++u) {
                    if (2 * u + 1 >= ncol) dr[u] &= (2 * u < ncol) ? 0x0000FFFFu : 0u;
                }
            }
            // ... and its dQ' / dK' MMAs reading dS^T (shared memory): they finish last, and this wait comes last
            if (n > 0) mbar_wait(&bars[L::bPdFree], (n - 1) & 1);
#pragma unroll
            for (int u = 0; u < 8; ++u) {                     // chunks of 8 queries (K elements of dK', M elements of dQ')
                const uint32_t off = tile_sw128_offset(r, wgc * 8 + u);
                *reinterpret_cast<uint4*>(smem + L::kDS + off) = make_uint4(dr[4 * u], dr[4 * u + 1], dr[4 * u + 2], dr[4 * u + 3]);
            }
            fence_proxy_async_smem();
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&bars[L::bPReady]);
            const long long d5 = dbg ? clock64() : 0;
            if (n + 1 < N) {
                asm volatile("" : "+f"(nx_stat));            // first use of the loaded value HERE
                sLDw[((n + 1) & 1) * 128 + r] = nx_stat * col_scale;
                bwd_bar_sync(1 + wgc);
            }
            if (dbg) { d_ws += d1 - d0; d_lds += d1b - d1; d_exp += d2 - d1; d_wpd += d3 - d2; d_wdp += d4 - d3; d_ds += d5 - d4; d_tail += clock64() - d5; }
        }
        const long long d_loop_end = dbg ? clock64() : 0;

        // ---- epilogue: warpgroup 0 finishes dV' (rho_k^T only with v_transform), warpgroup 1 dK'.
        // Phase A: thread = key row drains its accumulator row to an fp32 staging tile in shared memory ([D/4][128 rows][16 B],
        // 2064-byte pitch: conflict-free for both phases).  Phase B: the warpgroup walks the tile as (row, 8-column chunk) items
        // with the chunk index fastest over the lanes, block type by block type (the walk of rotate_tile): every global access —
        // the token's angles, the raw input chunk of the trans_coeff term, the output — is contiguous over neighbouring lanes.
        // (A thread-per-row epilogue spent 3 k clk per CTA in the load/store unit on 32-line accesses alone.)
        const int T = a.Tk;
        const int t0 = tile * 128;
        const int w4 = warp & 3;
        const int which = wgc;
        const uint32_t acc = lane_base + (which == 0 ? kFTmemDV : kFTmemDK);
        const bool rotate = which == 1 || a.v_transform;
        const bool want_tc = rotate && a.dtc != nullptr && a.hd.se3 > 0;
        const TIn* rawbase = which == 0 ? reinterpret_cast<const TIn*>(a.v) + b * a.v_sb + h * a.v_sh
                                        : reinterpret_cast<const TIn*>(a.k) + b * a.k_sb + h * a.k_sh;
        const int64_t raw_st = which == 0 ? a.v_st : a.k_st;
        TOut* gbase = reinterpret_cast<TOut*>(which == 0 ? a.dv : a.dk) + (static_cast<int64_t>(b) * T * a.H + h) * D;
        const float* so2_b = a.so2_k + static_cast<size_t>(b) * T * a.C * 2;
        constexpr uint32_t kEPitch = 2064;
        uint8_t* stage = smem + L::kStgQ + which * (D / 4) * kEPitch;
        static_assert(2u * (D / 4) * kEPitch <= 5u * L::kTile, "epilogue staging must fit in the streamed-tile rings (the dQ' staging tile behind them may still be read)");
        // the view matrices of the warp's first row, requested before the wait for the last MMAs; rows of another view reload them
        int cached_view = min(t0 + w4 * 32, T - 1) / a.tpvk;
        ViewReps vr;
        if (rotate) load_view_reps(vr, a.hd, a.se3_k + (static_cast<size_t>(b) * a.Nk + cached_view) * 16,
                                   a.so3_k + (static_cast<size_t>(b) * a.Nk + cached_view) * 34);
        float dtc_part = 0.f;
        long long d_done = 0, d_drain = 0, d_pre = 0;
        if constexpr (!std::is_void<LY>::value) {
            // ---- head layout known at compile time: the same walk as straight-line code with constant block boundaries
            // (the run-time-layout loop below executes ~8x the instructions per item: index arithmetic with divisions, a 3-way
            // block-type branch, operand rings — with two warps per scheduler that is 20 k clk per CTA)
            constexpr int k1 = LY::c1, k2 = LY::c2, k3 = LY::c3, K = D / 8;
            constexpr int nSo2 = K - k3;
            static_assert(LY::D == D, "layout / head dim mismatch");
            // row / chunk of the lane's item k (constant divisors)
            auto item_rc = [&](auto Kc, int& row, int& ch) {
                constexpr int k = decltype(Kc)::value;
                constexpr int sg = k < k1 ? 0 : (k < k2 ? 1 : (k < k3 ? 2 : 3));
                constexpr int st = sg == 0 ? 0 : (sg == 1 ? k1 : (sg == 2 ? k2 : k3));
                constexpr int n_t = (sg == 0 ? k1 : (sg == 1 ? k2 : (sg == 2 ? k3 : K))) - st;
                const int idx = lane + 32 * (k - st);
                const int rr = idx / n_t;
                ch = st + idx - rr * n_t;
                row = w4 * 32 + rr;
            };
            So2Chunk pre_sc[nSo2 > 0 ? nSo2 : 1];
            static_for<0, nSo2>([&](auto U) {
                constexpr int u = decltype(U)::value;
                pre_sc[u].a = make_float4(1.f, 0.f, 1.f, 0.f); pre_sc[u].b = pre_sc[u].a;
                int row, ch;
                item_rc(std::integral_constant<int, k3 + u>{}, row, ch);
                if (rotate && t0 + row < T) pre_sc[u] = load_so2_chunk(so2_b + static_cast<size_t>(t0 + row) * a.C * 2, ch, a.hd);
            });
            d_pre = dbg ? clock64() : 0;
            mbar_wait(&bars[L::bDone], 0);
            tc_fence_after();
            d_done = dbg ? clock64() : 0;
#pragma unroll
            for (int t3 = 0; t3 < D / 32; ++t3) {
                uint32_t v[32];
                tmem_ld32(acc + t3 * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    *reinterpret_cast<uint4*>(stage + (t3 * 8 + u) * kEPitch + r * 16) = make_uint4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
            }
            bwd_bar_sync(3 + which);
            d_drain = dbg ? clock64() : 0;
            // one pass per view among the warp's 32 rows (one pass except where a view boundary crosses the warp): the view
            // matrices are warp-uniform inside a pass, so an item costs two compares instead of a reload branch
            const int t_first = t0 + w4 * 32, t_last = min(t_first + 31, T - 1);
            const int v_first = t_first < T ? t_first / a.tpvk : 0, v_last = t_first < T ? t_last / a.tpvk : -1;
#pragma unroll 1
            for (int vw = v_first; vw <= v_last; ++vw) {
                if (rotate && vw != cached_view) {
                    cached_view = vw;
                    load_view_reps(vr, a.hd, a.se3_k + (static_cast<size_t>(b) * a.Nk + vw) * 16,
                                   a.so3_k + (static_cast<size_t>(b) * a.Nk + vw) * 34);
                }
                const int t_lo = vw * a.tpvk, t_hi = min(t_lo + a.tpvk, T);
                const float inv_m33 = want_tc ? 1.0f / vr.M[15] : 0.f;
                static_for<0, K>([&](auto Kc) {
                    constexpr int k = decltype(Kc)::value;
                    constexpr int sg = k < k1 ? 0 : (k < k2 ? 1 : (k < k3 ? 2 : 3));
                    int row, ch;
                    item_rc(Kc, row, ch);
                    const int t = t0 + row;
                    const bool mine = t >= t_lo && t < t_hi;
                    const float4 xa = *reinterpret_cast<const float4*>(stage + (2 * ch) * kEPitch + row * 16);
                    const float4 xb = *reinterpret_cast<const float4*>(stage + (2 * ch + 1) * kEPitch + row * 16);
                    float x[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
                    if constexpr (sg == 1) {
                        if (want_tc) {
                            // d(trans_coeff): the un-rotated gradient times d(rep)/d(tc) applied to the raw input 4-vectors (rows 0..2,
                            // column 3).  The raw w components come from the K' / V' tile image that sits in shared memory anyway
                            // (w' = M33 w: no global load of the raw rows, whose 32-line accesses stalled the load/store unit)
                            const uint4 kv = *reinterpret_cast<const uint4*>(smem + (which == 0 ? L::kFix1 : L::kFix0) + tile_sw64_offset(row, ch));
                            const float part = (x[0] * vr.M[3] + x[1] * vr.M[7] + x[2] * vr.M[11]) * bf16_hi(kv.y) +
                                               (x[4] * vr.M[3] + x[5] * vr.M[7] + x[6] * vr.M[11]) * bf16_hi(kv.w);
                            if (mine) dtc_part += part * inv_m33;
                        }
                        if (rotate) se3_apply_T(x, vr.M, tc);
                    } else if constexpr (sg == 2) {
                        if (rotate) so3_apply<true>(x, vr.W);
                    } else if constexpr (sg == 3) {
                        const So2Chunk& sc = pre_sc[k - k3];
                        const float cs8[8] = {sc.a.x, sc.a.y, sc.a.z, sc.a.w, sc.b.x, sc.b.y, sc.b.z, sc.b.w};
                        if (rotate) so2_apply<true>(x, cs8);
                    }
                    if (mine) store_chunk<TOut>(gbase + static_cast<int64_t>(t) * a.H * D + ch * 8, x);
                });
            }
        } else {
            // lane-item order (see phase B below): item k of a lane, k in [0, D/8), lies in the same block type for every lane
            const int k1 = a.hd.triv >> 3, k2 = k1 + (a.hd.se3 >> 3), k3 = k2 + (a.hd.so3 >> 3);
            auto item_rc = [&](int k, int& row, int& ch) {
                const int sg = k < k1 ? 0 : (k < k2 ? 1 : (k < k3 ? 2 : 3));
                const int st = sg == 0 ? 0 : (sg == 1 ? k1 : (sg == 2 ? k2 : k3));
                const int n_t = (sg == 0 ? k1 : (sg == 1 ? k2 : (sg == 2 ? k3 : D / 8))) - st;
                const int idx = lane + 32 * (k - st);
                const int rr = idx / n_t;
                ch = st + idx - rr * n_t;
                row = (t0 + w4 * 32 + rr < T) ? w4 * 32 + rr : -1;
            };
            // global operands of the lane's first kPre SE(3) items (raw input chunks of the trans_coeff term) and first kPre SO(2)
            // items (the token's angles): requested HERE, before the wait for the last MMAs — under the load of the reductions a
            // global round trip costs 1-2 k clk, and phase B has nothing to hide it behind
            constexpr int kPre = 4;
            RawChunk<TIn> pre_rw[kPre];
            So2Chunk pre_sc[kPre];
    #pragma unroll
            for (int u = 0; u < kPre; ++u) {
                zero_raw(pre_rw[u]);
                pre_sc[u].a = make_float4(1.f, 0.f, 1.f, 0.f); pre_sc[u].b = pre_sc[u].a;
                int row, ch;
                if (want_tc && k1 + u < k2) {
                    item_rc(k1 + u, row, ch);
                    if (row >= 0) load_raw(rawbase + (t0 + row) * raw_st + ch * 8, pre_rw[u]);
                }
                if (rotate && k3 + u < D / 8) {
                    item_rc(k3 + u, row, ch);
                    if (row >= 0) pre_sc[u] = load_so2_chunk(so2_b + static_cast<size_t>(t0 + row) * a.C * 2, ch, a.hd);
                }
            }
            mbar_wait(&bars[L::bDone], 0);
            tc_fence_after();
            d_done = dbg ? clock64() : 0;
    #pragma unroll
            for (int t3 = 0; t3 < D / 32; ++t3) {
                uint32_t v[32];
                tmem_ld32(acc + t3 * 32, v);
                tmem_ld_wait();
    #pragma unroll
                for (int u = 0; u < 8; ++u)
                    *reinterpret_cast<uint4*>(stage + (t3 * 8 + u) * kEPitch + r * 16) = make_uint4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
            }
            bwd_bar_sync(3 + which);
            d_drain = dbg ? clock64() : 0;
            // Phase B.  Each lane holds exactly n_t items of a block type with n_t chunks: inside the block type the warp's items are
            // (row rr, chunk cc), index rr * n_t + cc in [0, 32 n_t), lane l holding indices l, l + 32, ...  A compact loop (the
            // straight-line version of this walk was instruction-fetch bound); the shared-memory reads of item k+1 are issued before
            // item k is rotated and stored; global operands come from the prefetched registers (items past the first kPre of a block
            // type load theirs in place).
            const int view0 = t0 / a.tpvk, rem0 = t0 - view0 * a.tpvk;
            const bool one_boundary = a.tpvk >= 128;           // a tile then crosses at most one view boundary
            int row_c, ch_c, row_n = -1, ch_n = 0;
            float4 xa_c = make_float4(0.f, 0.f, 0.f, 0.f), xb_c = xa_c, xa_n = xa_c, xb_n = xa_c;
            item_rc(0, row_c, ch_c);
            if (row_c >= 0) {
                xa_c = *reinterpret_cast<const float4*>(stage + (2 * ch_c) * kEPitch + row_c * 16);
                xb_c = *reinterpret_cast<const float4*>(stage + (2 * ch_c + 1) * kEPitch + row_c * 16);
            }
    #pragma unroll 1
            for (int k = 0; k < D / 8; ++k) {
                if (k + 1 < D / 8) {
                    item_rc(k + 1, row_n, ch_n);
                    if (row_n >= 0) {
                        xa_n = *reinterpret_cast<const float4*>(stage + (2 * ch_n) * kEPitch + row_n * 16);
                        xb_n = *reinterpret_cast<const float4*>(stage + (2 * ch_n + 1) * kEPitch + row_n * 16);
                    }
                }
                const int sg = k < k1 ? 0 : (k < k2 ? 1 : (k < k3 ? 2 : 3));
                // this item's prefetched operands (slot 0 of its ring, then the ring moves up); late items load in place
                RawChunk<TIn> rw_c = pre_rw[0];
                So2Chunk sc_c = pre_sc[0];
                if (sg == 1) {
    #pragma unroll
                    for (int u = 0; u + 1 < kPre; ++u) pre_rw[u] = pre_rw[u + 1];
                    if (k - k1 >= kPre && want_tc && row_c >= 0) load_raw(rawbase + (t0 + row_c) * raw_st + ch_c * 8, rw_c);
                } else if (sg == 3) {
    #pragma unroll
                    for (int u = 0; u + 1 < kPre; ++u) pre_sc[u] = pre_sc[u + 1];
                    if (k - k3 >= kPre && rotate && row_c >= 0)
                        sc_c = load_so2_chunk(so2_b + static_cast<size_t>(t0 + row_c) * a.C * 2, ch_c, a.hd);
                }
                if (row_c >= 0) {
                    const int t = t0 + row_c;
                    float x[8] = {xa_c.x, xa_c.y, xa_c.z, xa_c.w, xb_c.x, xb_c.y, xb_c.z, xb_c.w};
                    if (rotate && (sg == 1 || sg == 2)) {
                        const int vrow = one_boundary ? view0 + (rem0 + row_c >= a.tpvk ? 1 : 0) : t / a.tpvk;
                        if (vrow != cached_view) {
                            cached_view = vrow;
                            load_view_reps(vr, a.hd, a.se3_k + (static_cast<size_t>(b) * a.Nk + vrow) * 16,
                                           a.so3_k + (static_cast<size_t>(b) * a.Nk + vrow) * 34);
                        }
                    }
                    if (sg == 1) {
                        if (want_tc) {
                            // d(trans_coeff): the un-rotated gradient times d(rep)/d(tc) applied to the raw input 4-vectors (rows 0..2, column 3)
                            float xin[8];
                            raw_to_f32(rw_c, xin);
    #pragma unroll
                            for (int v4 = 0; v4 < 2; ++v4)
                                dtc_part += (x[4 * v4] * vr.M[3] + x[4 * v4 + 1] * vr.M[7] + x[4 * v4 + 2] * vr.M[11]) * xin[4 * v4 + 3];
                        }
                        if (rotate) se3_apply_T(x, vr.M, tc);
                    } else if (sg == 2) {
                        if (rotate) so3_apply<true>(x, vr.W);
                    } else if (sg == 3) {
                        const float cs8[8] = {sc_c.a.x, sc_c.a.y, sc_c.a.z, sc_c.a.w, sc_c.b.x, sc_c.b.y, sc_c.b.z, sc_c.b.w};
                        if (rotate) so2_apply<true>(x, cs8);
                    }
                    store_chunk<TOut>(gbase + static_cast<int64_t>(t) * a.H * D + ch_c * 8, x);
                }
                row_c = row_n; ch_c = ch_n; xa_c = xa_n; xb_c = xb_n;
            }
        }
        if (a.dtc) {
            dtc_part = warp_sum(dtc_part);
            if (lane == 0 && dtc_part != 0.f) atomicAdd(a.dtc, dtc_part);
        }
        if (dbg) {
            const long long d_end = clock64();
            dbg[0] = d_end - d_start; dbg[1] = d_ws; dbg[2] = d_exp; dbg[3] = d_wpd; dbg[4] = d_wdp; dbg[5] = d_ds; dbg[6] = N;
            dbg[7] = d_tail; dbg[8] = d_loop_end - d_start; dbg[9] = d_done - d_loop_end; dbg[10] = d_drain - d_done; dbg[11] = d_end - d_drain; dbg[12] = d_lds; dbg[13] = d_pre ? d_pre - d_loop_end : 0;
        }
    } else if (warp < 12) {
        // =========================================================== reducer warpgroup: thread r <-> query row r of the dQ' partial
        // The partial leaves in two halves of D/2 columns through ONE half-size staging tile: tensor memory -> registers -> shared
        // memory -> bulk reduction into the fp32 accumulation tile.  dQ' is released when its second half is in registers.
        setmaxnreg_dec<88>();
        const int r = threadIdx.x - 256;
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
        uint8_t* stage = smem + L::kDQ;
        float* acc_bh = a.dq_acc + bh * a.ntq * (128u * D);
        constexpr int kHalf = D / 2;
#pragma unroll 1
        for (int n = 0; n < N; ++n) {
            float* dst = acc_bh + static_cast<size_t>(qtile(n)) * (128u * D);
            mbar_wait(&bars[L::bDQFull], n & 1);
            tc_fence_after();
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                uint32_t v[kHalf];
                if constexpr (kHalf >= 32) tmem_ld32(lane_base + kFTmemDP + hf * kHalf, v);
                if constexpr (kHalf % 32 == 16) tmem_ld16(lane_base + kFTmemDP + hf * kHalf + (kHalf / 32) * 32, v + (kHalf / 32) * 32);
                tmem_ld_wait();
                if (hf == 1) {                                // dQ' is in registers: dP^T of the next pair may overwrite it
                    tc_fence_before();
                    mbar_arrive(&bars[L::bDQFree]);
                }
                if (n > 0 || hf > 0) {                        // the previous bulk reduction must be done READING the staging tile
                    if (r == 0) bulk_wait_read_all();
                    bwd_bar_sync(5);
                }
#pragma unroll
                for (int u = 0; u < kHalf / 4; ++u)
                    *reinterpret_cast<uint4*>(stage + u * 2048 + r * 16) = make_uint4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
                fence_proxy_async_smem();
                bwd_bar_sync(5);
                if (r == 0) {
                    bulk_reduce_add_f32(dst + hf * (kHalf / 4) * 512, stage, 128u * kHalf * 4u);
                    bulk_commit_group();
                }
            }
        }
        if (r == 0) bulk_wait_all();
    } else {
        setmaxnreg_dec<56>();
        if (warp == 12) {
            // ======================================================= UMMA issuer
            constexpr uint32_t idesc_ss = make_idesc_bf16(128, 128, 0, 0);     // S^T, dP^T
            constexpr uint32_t idesc_kn = make_idesc_bf16(128, D, 0, 1);       // dV' (A in tensor memory), dK' (A K-major)
            constexpr uint32_t idesc_nn = make_idesc_bf16(128, D, 1, 1);       // dQ' (A = dS^T tile read MN-major)
            const uint32_t f0 = smem_u32(smem + L::kFix0), f1 = smem_u32(smem + L::kFix1);
            const uint32_t ds_sm = smem_u32(smem + L::kDS);
            mbar_wait(&bars[L::bFix], 0);
            // ring slots of pair n: Q' n % 3, dO' n % 2, with the parities of their full barriers — kept as counters
            auto issue_s = [&](int sq, int ph) {
                const uint32_t g0 = smem_u32(smem + L::kStgQ + sq * L::kTile);
                mbar_wait(&bars[L::bFullQ + sq], ph);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < D / 16; ++kk)
                        umma_ss(tmem_base + kFTmemS, desc_kmajor_sw64(f0, kk), desc_kmajor_sw64(g0, kk), idesc_ss, kk > 0);
                    umma_commit(&bars[L::bSFull]);
                }
                __syncwarp();
            };
            auto issue_dp = [&](int so, int ph) {
                const uint32_t g1 = smem_u32(smem + L::kStgO + so * L::kTile);
                mbar_wait(&bars[L::bFullO + so], ph);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < D / 16; ++kk)
                        umma_ss(tmem_base + kFTmemDP, desc_kmajor_sw64(f1, kk), desc_kmajor_sw64(g1, kk), idesc_ss, kk > 0);
                    umma_commit(&bars[L::bDPFull]);
                }
                __syncwarp();
            };
            issue_s(0, 0);
            issue_dp(0, 0);
            int sq = 0, sq1 = 1, phq1 = 0;                    // sq = n % 3; sq1 = (n + 1) % 3 with parity phq1
            int so = 0, so1 = 1, pho1 = 0;                    // so = n % 2; so1 = (n + 1) % 2 with parity pho1
#pragma unroll 1
            for (int n = 0; n < N; ++n) {
                const uint32_t g0 = smem_u32(smem + L::kStgQ + sq * L::kTile), g1 = smem_u32(smem + L::kStgO + so * L::kTile);
                if (n + 1 < N) {                             // next pair's S^T as soon as this pair's is in registers
                    mbar_wait(&bars[L::bSFree], n & 1);
                    tc_fence_after();
                    issue_s(sq1, phq1);
                }
                mbar_wait(&bars[L::bPReady], n & 1);
                tc_fence_after();
                const uint32_t accf = (n > 0) ? 1u : 0u;
                if (elect_one()) {
                    // dQ'(partial) = dS K'_j first: the reducer drains it while dV' executes
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk)
                        umma_ss(tmem_base + kFTmemDP, desc_ds_mnmajor_sw128(ds_sm, kk), desc_mnmajor_sw64(f0, kk), idesc_nn, kk > 0);
                    umma_commit(&bars[L::bDQFull]);
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk)
                        umma_ts(tmem_base + kFTmemDV, tmem_base + kFTmemP + kk * 8, desc_mnmajor_sw64(g1, kk), idesc_kn, (kk > 0) ? 1u : accf);
                    umma_commit(&bars[L::bEmptyO + so]);      // dO'_n has no reader left
                    umma_commit(&bars[L::bPFree]);            // ... and P^T may be overwritten
                }
                __syncwarp();
                if (n + 1 < N) {                             // dP^T of the next pair reuses the dQ' columns
                    mbar_wait(&bars[L::bDQFree], n & 1);
                    tc_fence_after();
                    issue_dp(so1, pho1);
                }
                if (elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk)
                        umma_ss(tmem_base + kFTmemDK, desc_p_sw128(ds_sm, kk), desc_mnmajor_sw64(g0, kk), idesc_kn, (kk > 0) ? 1u : accf);
                    umma_commit(&bars[L::bEmptyQ + sq]);
                    umma_commit(&bars[L::bPdFree]);
                    if (n == N - 1) umma_commit(&bars[L::bDone]);
                }
                __syncwarp();
                sq = sq1;
                if (++sq1 == L::kStagesQ) { sq1 = 0; phq1 ^= 1; }
                so = so1;
                if (++so1 == L::kStagesO) { so1 = 0; pho1 ^= 1; }
            }
        } else if (warp == 13) {
            // ======================================================= bulk-copy producer
            if (lane == 0) {
                mbar_arrive_expect_tx(&bars[L::bFix], 2 * L::kTile);
                bulk_g2s(smem + L::kFix0, fix0, L::kTile, &bars[L::bFix]);
                bulk_g2s(smem + L::kFix1, fix1, L::kTile, &bars[L::bFix]);
            }
            int sq = 0, phq = 1, so = 0, pho = 1;             // ring slots; parities of the previous completion of their empty barriers
#pragma unroll 1
            for (int n = 0; n < N; ++n) {
                const int qi = qtile(n);
                if (n >= L::kStagesQ) mbar_wait(&bars[L::bEmptyQ + sq], phq);
                if (lane == 0) {
                    mbar_arrive_expect_tx(&bars[L::bFullQ + sq], L::kTile);
                    bulk_g2s(smem + L::kStgQ + sq * L::kTile, stg0 + static_cast<size_t>(qi) * L::kTile, L::kTile, &bars[L::bFullQ + sq]);
                }
                __syncwarp();
                if (n >= L::kStagesO) mbar_wait(&bars[L::bEmptyO + so], pho);
                if (lane == 0) {
                    mbar_arrive_expect_tx(&bars[L::bFullO + so], L::kTile);
                    bulk_g2s(smem + L::kStgO + so * L::kTile, stg1 + static_cast<size_t>(qi) * L::kTile, L::kTile, &bars[L::bFullO + so]);
                }
                __syncwarp();
                if (++sq == L::kStagesQ) { sq = 0; phq ^= 1; }
                if (++so == L::kStagesO) { so = 0; pho ^= 1; }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 12) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// dq = rho_q^{-1} dQ' from the fp32 accumulation tiles (+ the query-side trans_coeff term).  HBM-bound: 4 D bytes read +
// sizeof(TOut) D written per row.  The (row, 8-column chunk) walk of the fused kernel's epilogue: the chunk index is fastest
// over the lanes inside a block type, so the accumulation tile ([D/4][128 rows][4 floats]), the token's angles, the raw q
// chunks of the trans_coeff term and the output rows are all read / written in contiguous pieces by neighbouring lanes
// (a thread-per-row version with a shared-memory transpose ran at 3.4 TB/s).
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(128) bwd_dq_finish_kernel(const BwdArgs a, const int D) {
    const int tile = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int lane = threadIdx.x & 31, w4 = threadIdx.x >> 5;
    const size_t bh = static_cast<size_t>(b) * a.H + h;
    const float* acc = a.dq_acc + (bh * a.ntq + tile) * (128u * D);
    const int t0 = tile * 128;
    const int K = D >> 3;
    const int k1 = a.hd.triv >> 3, k2 = k1 + (a.hd.se3 >> 3), k3 = k2 + (a.hd.so3 >> 3);
    const float tc = a.tc_ptr ? __ldg(a.tc_ptr) : 1.0f;
    const TIn* rawbase = reinterpret_cast<const TIn*>(a.q) + b * a.q_sb + h * a.q_sh;
    TOut* gbase = reinterpret_cast<TOut*>(a.dq) + (static_cast<int64_t>(b) * a.Tq * a.H + h) * D;
    const float* so2_b = a.so2_q + static_cast<size_t>(b) * a.Tq * a.C * 2;
    const bool want_tc = a.dtc != nullptr && a.hd.se3 > 0;
    int cached_view = -1;
    ViewReps vr;
    float dtc_part = 0.f;
#pragma unroll 1
    for (int k = 0; k < K; ++k) {
        const int sg = k < k1 ? 0 : (k < k2 ? 1 : (k < k3 ? 2 : 3));
        const int st = sg == 0 ? 0 : (sg == 1 ? k1 : (sg == 2 ? k2 : k3));
        const int n_t = (sg == 0 ? k1 : (sg == 1 ? k2 : (sg == 2 ? k3 : K))) - st;
        const int idx = lane + 32 * (k - st);
        const int rr = idx / n_t;
        const int ch = st + idx - rr * n_t;
        const int row = w4 * 32 + rr;
        const int t = t0 + row;
        if (t >= a.Tq) continue;
        const float4 xa = __ldg(reinterpret_cast<const float4*>(acc + (2 * ch) * 512 + row * 4));
        const float4 xb = __ldg(reinterpret_cast<const float4*>(acc + (2 * ch + 1) * 512 + row * 4));
        float x[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
        if (sg == 1 || sg == 2) {
            const int vrow = t / a.tpvq;
            if (vrow != cached_view) {
                cached_view = vrow;
                load_view_reps(vr, a.hd, a.se3_q + (static_cast<size_t>(b) * a.Nq + vrow) * 16, a.so3_q + (static_cast<size_t>(b) * a.Nq + vrow) * 34);
            }
        }
        if (sg == 1) {
            if (want_tc) {
                // Q' = (E_q msk)^T q: d/dtc = g_3 (M_03 q_0 + M_13 q_1 + M_23 q_2) per SE(3) 4-vector
                RawChunk<TIn> rc;
                load_raw(rawbase + t * a.q_st + ch * 8, rc);
                float xin[8];
                raw_to_f32(rc, xin);
#pragma unroll
                for (int v4 = 0; v4 < 2; ++v4)
                    dtc_part += x[4 * v4 + 3] * (vr.M[3] * xin[4 * v4] + vr.M[7] * xin[4 * v4 + 1] + vr.M[11] * xin[4 * v4 + 2]);
            }
            se3_apply(x, vr.M, tc);
        } else if (sg == 2) {
            so3_apply<true>(x, vr.W);
        } else if (sg == 3) {
            const So2Chunk sc = load_so2_chunk(so2_b + static_cast<size_t>(t) * a.C * 2, ch, a.hd);
            const float cs8[8] = {sc.a.x, sc.a.y, sc.a.z, sc.a.w, sc.b.x, sc.b.y, sc.b.z, sc.b.w};
            so2_apply<true>(x, cs8);
        }
        store_chunk<TOut>(gbase + static_cast<int64_t>(t) * a.H * D + ch * 8, x);
    }
    if (want_tc) {
        __shared__ float red[4];
        dtc_part = warp_sum(dtc_part);
        if (lane == 0) red[w4] = dtc_part;
        __syncthreads();
        if (threadIdx.x == 0) {
            const float s = red[0] + red[1] + red[2] + red[3];
            if (s != 0.f) atomicAdd(a.dtc, s);
        }
    }
}

// ---------------------------------------------------------------------------------------------------- host side
size_t bwd_dq_acc_bytes(int B, int H, int Tq, int D) {
    return static_cast<size_t>(B) * H * num_kv_tiles(Tq) * 128u * D * sizeof(float);
}

bool bwd_fused_supported(int D) { return D == 32 || D == 64 || D == 96; }

template <typename T, int D, typename LY>
static int launch_bwd_fused_d(const BwdArgs& a, cudaStream_t st) {
    using L = FusedSmem<D>;
    auto kern = attn_bwd_fused_kernel<T, T, D, LY>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(L::kBytes));
    if (e != cudaSuccess) return set_error(GTA_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    e = cudaMemsetAsync(a.dq_acc, 0, bwd_dq_acc_bytes(a.B, a.H, a.Tq, D), st);
    if (e != cudaSuccess) return set_error(GTA_ERR_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
    kern<<<dim3(a.ntk, a.H, a.B), kFThreads, L::kBytes, st>>>(a);
    bwd_dq_finish_kernel<T, T><<<dim3(a.ntq, a.H, a.B), 128, 0, st>>>(a, D);
    return check_launch("gta_attn_bwd (fused)");
}

// the shipped head layouts get a straight-line epilogue (runs/msn/GTA/gta_so3, runs/msn/GTA/gta, runs/clevrtr/GTA/gta, BASELINE
// config 1 and its so3 variant); everything else (and GTA_FLAG_RUNTIME_LAYOUT) the run-time-layout code
template <typename T>
static int launch_bwd_fused_t(const BwdArgs& a, int D, bool runtime_layout, cudaStream_t st) {
    const HeadDims& hd = a.hd;
    auto is = [&](int tr, int se3, int so3, int so2) { return !runtime_layout && hd.triv == tr && hd.se3 == se3 && hd.so3 == so3 && hd.so2 == so2; };
    switch (D) {
        case 32:
            if (is(0, 16, 0, 16)) return launch_bwd_fused_d<T, 32, HeadLayout<0, 16, 0, 16>>(a, st);
            if (is(0, 16, 8, 8)) return launch_bwd_fused_d<T, 32, HeadLayout<0, 16, 8, 8>>(a, st);
            return launch_bwd_fused_d<T, 32, void>(a, st);
        case 64:
            if (is(0, 32, 0, 32)) return launch_bwd_fused_d<T, 64, HeadLayout<0, 32, 0, 32>>(a, st);
            return launch_bwd_fused_d<T, 64, void>(a, st);
        case 96:
            if (is(0, 48, 24, 24)) return launch_bwd_fused_d<T, 96, HeadLayout<0, 48, 24, 24>>(a, st);
            if (is(0, 48, 0, 48)) return launch_bwd_fused_d<T, 96, HeadLayout<0, 48, 0, 48>>(a, st);
            return launch_bwd_fused_d<T, 96, void>(a, st);
    }
    return set_error(GTA_ERR_UNSUPPORTED, "fused backward: head dim %d", D);
}

int launch_bwd_fused(const BwdArgs& a, bool bf16, int D, bool runtime_layout, cudaStream_t st) {
    return bf16 ? launch_bwd_fused_t<__nv_bfloat16>(a, D, runtime_layout, st) : launch_bwd_fused_t<float>(a, D, runtime_layout, st);
}

}  // namespace gta
