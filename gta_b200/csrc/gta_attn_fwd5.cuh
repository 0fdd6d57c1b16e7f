// Fused GTA attention forward, v4 pipeline: the persistent two-tile tcgen05 pipeline of gta_attn_fwd3.cu with the
// softmax warps freed of everything that is not the softmax:
//
//   warps 0-3 / 4-7   softmax warpgroup A / B, STREAMING: a 128-key score row is processed in four 32-column quarters
//                     (tcgen05.ld of quarter q+1 in flight while quarter q is exponentiated; P quarter q is packed to bf16
//                     and stored over the S columns it came from), so a thread holds 64 + 16 instead of 128 score
//                     registers                                                                   112 registers
//   warp  8           UMMA issuer            } as in gta_attn_fwd3.cu
//   warp  9           bulk-copy producer     }                                                     88 registers
//   warps 10-11       Q stager               }
//   warps 12-15       EPILOGUE warpgroup: thread i <-> row i of the tile being finished.  It drains the whole O row from
//                     tensor memory into registers with one round trip, releases O at once (o_free) — the next item's PV
//                     never waits for the output rotation — and then normalises, applies rho_q^{-1} and stores from
//                     registers while both softmax warpgroups are already in the next item            200 registers
//                     (2 * 112 + 88 + 200 = 512 = the whole register file for 4 x 128 threads)
//
// The row statistics (running sum l and reference maximum mu) travel from the softmax thread to the epilogue thread through
// a double-buffered shared-memory array and the `stat` barrier.
//
// Lazy rescaling in the streaming form: exponentials use the reference mu (log2 units) of the previous quarters; only when
// a quarter's maximum exceeds it by more than 2^8 is mu advanced — by a whole number of log2 units, so that the already
// stored P quarters of the tile and the O accumulator are rescaled by an exact power of two.
//
// The head layout [triv | se3 | so3 | so2] is a COMPILE-TIME parameter here (the epilogue runs straight-line from the 96
// accumulator registers); the layouts of the shipped configs are instantiated and anything else keeps the gta_attn_fwd3.cu
// pipeline.  Reference semantics: source/utils/gta.py:92-279 and source/layers.py:202-211.
#pragma once
#include <cmath>
#include <type_traits>

#include "attn_common.cuh"

namespace gta {

constexpr int kThreads5 = 512;
constexpr uint32_t k5TmemSA = 0, k5TmemSB = 128, k5TmemOA = 256, k5TmemOB = 384;
constexpr float k5RescaleThreshold = 8.0f;   // log2 units
#ifndef GTA5_REG_SOFTMAX
#define GTA5_REG_SOFTMAX 112
#endif
#ifndef GTA5_REG_ISSUE
#define GTA5_REG_ISSUE 88
#endif
#ifndef GTA5_REG_EPI
#define GTA5_REG_EPI 200
#endif
static_assert(2 * GTA5_REG_SOFTMAX + GTA5_REG_ISSUE + GTA5_REG_EPI <= 512, "register split exceeds the register file");

template <int D>
struct Attn5Cfg {
    static constexpr int kStages = 2;
    static constexpr uint32_t kTile = 128u * D * 2u;
    static constexpr uint32_t kQ = 0;                          // [2 buffers][2 tiles]
    static constexpr uint32_t kK = 4 * kTile;                  // [kStages]
    static constexpr uint32_t kV = kTile * (4 + kStages);      // [kStages]
    static constexpr uint32_t kStats = kTile * (4 + 2 * kStages);   // float2 [2 parities][2 tiles][128 rows]
    static constexpr uint32_t kBars = kStats + 2 * 2 * 128 * 8;
    enum : int {
        bQFull = 0,                        // [buf][X]  count 64 (stager threads)
        bQFree = 4,                        // [buf][X]  tcgen05.commit after the item's last QK_X
        bKFull = 8,                        // [kStages]
        bVFull = bKFull + kStages,
        bKEmpty = bVFull + kStages,
        bVEmpty = bKEmpty + kStages,
        bSFull = bVEmpty + kStages,        // [X] commit
        bPFull = bSFull + 2,               // [X] count 128
        bOFinal = bPFull + 2,              // [X] commit after the item's last PV_X
        bOFree = bOFinal + 2,              // [X] count 128: O_X drained to registers by the epilogue warpgroup
        bStat = bOFree + 2,                // [parity][X] count 128: row statistics of the item written
        bStatFree = bStat + 4,             // [parity][X] count 128: ... and read by the epilogue warpgroup
        bCount = bStatFree + 4
    };
    static constexpr uint32_t kTmemSlot = kBars + bCount * 8;
    static constexpr uint32_t kUsed = kTmemSlot + 16;
    static constexpr uint32_t kBytes = (kUsed + 1024 > 120u * 1024u) ? kUsed + 1024 : 120u * 1024u;
};

struct ItemCoord5 {
    int b, h, p;
    bool has_b;
};
__device__ __forceinline__ ItemCoord5 decode_item5(int item, int npairs, int H, int Tq) {
    ItemCoord5 c;
    c.p = item % npairs;
    const int bh = item / npairs;
    c.h = bh % H;
    c.b = bh / H;
    c.has_b = (c.p * 256 + 128) < Tq;
    return c;
}

__device__ __forceinline__ uint32_t mul_bf16x2(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}

template <typename TIn, typename TOut, typename LY>
__global__ void __launch_bounds__(kThreads5, 1) attn_fwd5_kernel(const AttnArgs a, const int npairs, const int nitems) {
    constexpr int D = LY::D;
    using L = Attn5Cfg<D>;
    constexpr int NS = L::kStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBars);
    float2* stats = reinterpret_cast<float2*>(smem + L::kStats);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kTmemSlot);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = a.ntiles_k;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) {
            mbar_init(&bars[L::bQFull + i], 64);
            mbar_init(&bars[L::bQFree + i], 1);
            mbar_init(&bars[L::bStat + i], 128);
            mbar_init(&bars[L::bStatFree + i], 128);
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
        // =========================================================== softmax warpgroups (streaming)
        setmaxnreg_dec<GTA5_REG_SOFTMAX>();
        const int X = warp >> 2;
        const int r = threadIdx.x & 127;
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
        const uint32_t s_addr = lane_base + (X ? k5TmemSB : k5TmemSA);
        const uint32_t o_addr = lane_base + (X ? k5TmemOB : k5TmemOA);
        const float cs = a.scale_log2;
        const uint64_t cs2 = pack_f32x2(cs, cs);
        uint32_t gt = 0;      // tiles processed by this warpgroup (s_full / p_full phase)
        uint32_t cnt = 0;     // items processed by this warpgroup (stat parity / phase)

#pragma unroll 1
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const ItemCoord5 ic = decode_item5(item, npairs, a.H, a.Tq);
            if (X == 1 && !ic.has_b) continue;
            float mu = -INFINITY;     // reference maximum of the row, log2 units (score * scale * log2 e)
            float l_run = 0.f;

#pragma unroll 1
            for (int j = 0; j < n; ++j, ++gt) {
                const int nvalid = (j == n - 1) ? a.Tk - j * 128 : 128;
                mbar_wait(&bars[L::bSFull + X], gt & 1);
                tc_fence_after();
                uint32_t sa[32], sb[32];
                uint64_t lsum2 = pack_f32x2(0.f, 0.f);
                tmem_ld32(s_addr, sa);

                // one 32-column quarter q held in `sq`; the load of quarter q + 1 is already in flight
                auto quarter = [&](uint32_t* sq, const int q, auto masked) {
                    float* s = reinterpret_cast<float*>(sq);
                    if constexpr (decltype(masked)::value) {      // ragged last key tile only (its own copy of the code)
#pragma unroll
                        for (int i = 0; i < 32; ++i) if (q * 32 + i >= nvalid) s[i] = -INFINITY;
                    }
                    float mx0 = fmax3(s[0], s[1], s[2]), mx1 = fmax3(s[3], s[4], s[5]);
#pragma unroll
                    for (int i = 6; i < 30; i += 6) {
                        mx0 = fmax3(mx0, s[i], s[i + 1]); mx1 = fmax3(mx1, s[i + 2], s[i + 3]);
                        mx0 = fmax3(mx0, s[i + 4], s[i + 5]);
                    }
                    const float qmax = fmax3(mx0, mx1, fmaxf(s[30], s[31])) * cs;
                    const bool grow = qmax - mu > k5RescaleThreshold;     // always true on the item's first quarter (mu = -inf)
                    if (__any_sync(0xffffffffu, grow)) {
                        // advance the reference by a whole number of log2 units: alpha is an exact power of two
                        const float mu_new = grow ? (mu == -INFINITY ? qmax : mu + ceilf(qmax - mu)) : mu;
                        const float alpha = grow ? fast_exp2(mu - mu_new) : 1.0f;
                        l_run *= alpha;
                        float l0, l1;
                        unpack_f32x2(lsum2, l0, l1);
                        lsum2 = pack_f32x2(l0 * alpha, l1 * alpha);
                        mu = mu_new;
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
                        if (q > 0) {
                            // the P quarters of this tile that are already stored (columns [0, 16 q))
                            tmem_st_wait();
                            const uint32_t al2 = pack_bf16x2(alpha, alpha);
#pragma unroll 1
                            for (int c8 = 0; c8 < 2 * q; ++c8) {
                                uint32_t p8[8];
                                tmem_ld8(s_addr + c8 * 8, p8);
                                tmem_ld_wait();
#pragma unroll
                                for (int i = 0; i < 8; ++i) p8[i] = mul_bf16x2(p8[i], al2);
                                tmem_st8(s_addr + c8 * 8, p8);
                            }
                        }
                    }
                    const uint64_t neg2 = pack_f32x2(-mu, -mu);
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const uint64_t x2 = ffma2(pack_f32x2(s[2 * i], s[2 * i + 1]), cs2, neg2);
                        float x0, x1;
                        unpack_f32x2(x2, x0, x1);
                        const float p0 = fast_exp2(x0), p1 = fast_exp2(x1);
                        lsum2 = fadd2(lsum2, pack_f32x2(p0, p1));
                        pk[i] = pack_bf16x2(p0, p1);
                    }
                    tmem_st16(s_addr + q * 16, pk);
                };

                auto tile_body = [&](auto masked) {
                    tmem_ld_wait32(sa);
                    tmem_ld32(s_addr + 32, sb);
                    quarter(sa, 0, masked);
                    tmem_ld_wait32(sb);
                    tmem_ld32(s_addr + 64, sa);
                    quarter(sb, 1, masked);
                    tmem_ld_wait32(sa);
                    tmem_ld32(s_addr + 96, sb);
                    quarter(sa, 2, masked);
                    tmem_ld_wait32(sb);
                    quarter(sb, 3, masked);
                };
                if (nvalid < 128) tile_body(std::true_type{}); else tile_body(std::false_type{});
                float ls0, ls1;
                unpack_f32x2(lsum2, ls0, ls1);
                l_run += ls0 + ls1;
                if (j == n - 1) {
                    // hand the row statistics to the epilogue warpgroup (double-buffered by item parity)
                    const uint32_t par = cnt & 1;
                    if (cnt >= 2) mbar_wait(&bars[L::bStatFree + par * 2 + X], ((cnt >> 1) - 1) & 1);
                    stats[(par * 2 + X) * 128 + r] = make_float2(l_run, mu);
                    mbar_arrive(&bars[L::bStat + par * 2 + X]);
                }
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(&bars[L::bPFull + X]);
            }
            ++cnt;
        }
    } else if (warp >= 12) {
        // =========================================================== epilogue warpgroup
        setmaxnreg_inc<GTA5_REG_EPI>();
        const int r = threadIdx.x - 384;                      // row of the tile = TMEM lane
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
        uint32_t cntx[2] = {0, 0};
#pragma unroll 1
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const ItemCoord5 ic = decode_item5(item, npairs, a.H, a.Tq);
#pragma unroll 1
            for (int X = 0; X < 2; ++X) {
                if (X == 1 && !ic.has_b) continue;
                const uint32_t c_ = cntx[X]++;
                const uint32_t par = c_ & 1;
                const int t = ic.p * 256 + X * 128 + r;
                const bool valid = t < a.Tq;
                const int tt = valid ? t : a.Tq - 1;
                const size_t view = static_cast<size_t>(ic.b) * a.Nq + tt / a.tpvq;
                const float* so2 = a.so2_q + (static_cast<size_t>(ic.b) * a.Tq + tt) * a.C * 2;
                const bool vt = a.v_transform != 0;
                // the row's view matrices are requested before the waits
                float M[LY::kSe3 ? 16 : 1], W[LY::kSo3 ? 34 : 1];
                if constexpr (LY::kSe3 > 0) {
                    if (vt) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float4 q4 = __ldg(reinterpret_cast<const float4*>(a.se3_q + view * 16) + i);
                            M[4 * i] = q4.x; M[4 * i + 1] = q4.y; M[4 * i + 2] = q4.z; M[4 * i + 3] = q4.w;
                        }
                    }
                }
                if constexpr (LY::kSo3 > 0) {
                    if (vt) {
#pragma unroll
                        for (int i = 0; i < 17; ++i) {
                            const float2 q2 = __ldg(reinterpret_cast<const float2*>(a.so3_q + view * 34) + i);
                            W[2 * i] = q2.x; W[2 * i + 1] = q2.y;
                        }
                    }
                }
                mbar_wait(&bars[L::bStat + par * 2 + X], (c_ >> 1) & 1);
                const float2 st2 = stats[(par * 2 + X) * 128 + r];
                mbar_arrive(&bars[L::bStatFree + par * 2 + X]);
                mbar_wait(&bars[L::bOFinal + X], c_ & 1);
                tc_fence_after();
                uint32_t o[D];
                const uint32_t o_addr = lane_base + (X ? k5TmemOB : k5TmemOA);
#pragma unroll
                for (int c32 = 0; c32 < D / 32; ++c32) tmem_ld32(o_addr + c32 * 32, o + c32 * 32);
#pragma unroll
                for (int c32 = 0; c32 < D / 32; ++c32) tmem_ld_wait32(o + c32 * 32);
                tc_fence_before();
                mbar_arrive(&bars[L::bOFree + X]);              // O_X is in registers: the next item's PV_X(0) may overwrite it

                const float inv_l = 1.0f / st2.x;
                TOut* orow = reinterpret_cast<TOut*>(a.out) + ((static_cast<int64_t>(ic.b) * a.Tq + tt) * a.H + ic.h) * D;
                uint4 pend = make_uint4(0, 0, 0, 0);
                static_for<0, D / 8>([&](auto icn) {
                    constexpr int c = decltype(icn)::value;
                    float x[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) x[i] = __uint_as_float(o[c * 8 + i]) * inv_l;
                    if (vt) {
                        if constexpr (c < LY::c1) {
                        } else if constexpr (c < LY::c2) {
                            se3_apply(x, M, tc);
                        } else if constexpr (c < LY::c3) {
                            so3_apply<true>(x, W);
                        } else {
                            const float4 c0 = __ldg(reinterpret_cast<const float4*>(so2 + (c - LY::c3) * 8));
                            const float4 c1v = __ldg(reinterpret_cast<const float4*>(so2 + (c - LY::c3) * 8) + 1);
                            const float cs8[8] = {c0.x, c0.y, c0.z, c0.w, c1v.x, c1v.y, c1v.z, c1v.w};
                            so2_apply<true>(x, cs8);
                        }
                    }
                    if (sizeof(TOut) == 4) {
                        if (valid)
                            st_global_v8(orow + c * 8, make_uint4(__float_as_uint(x[0]), __float_as_uint(x[1]), __float_as_uint(x[2]), __float_as_uint(x[3])),
                                         make_uint4(__float_as_uint(x[4]), __float_as_uint(x[5]), __float_as_uint(x[6]), __float_as_uint(x[7])));
                    } else {
                        const uint4 pk = pack_chunk_bf16(x);
                        if (c & 1) { if (valid) st_global_v8(orow + (c - 1) * 8, pend, pk); }
                        else pend = pk;
                    }
                });
                if (a.lse && valid)
                    a.lse[(static_cast<int64_t>(ic.b) * a.H + ic.h) * a.Tq + t] = st2.y * 0.6931471805599453f + logf(st2.x);
            }
        }
    } else {
      setmaxnreg_dec<GTA5_REG_ISSUE>();
      if (warp >= 10) {
        // =========================================================== Q stager (runs one item ahead)
        const int r0 = threadIdx.x - 320;    // 0..63; this thread stages rows r0 and r0 + 64 of each tile
        uint32_t cntx[2] = {0, 0};          // items staged per tile slot (buffer = cnt & 1, phase = (cnt >> 1) & 1)
#pragma unroll 1
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const ItemCoord5 ic = decode_item5(item, npairs, a.H, a.Tq);
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

#pragma unroll 1
            for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
                const ItemCoord5 ic = decode_item5(item, npairs, a.H, a.Tq);
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
                        const uint32_t d_addr = tmem_base + (X ? k5TmemSB : k5TmemSA);
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
                    mbar_wait(&bars[L::bPFull + X], (gtx[X] + j) & 1);
                    if (j == 0 && cntx[X] > 0) mbar_wait(&bars[L::bOFree + X], (cntx[X] - 1) & 1);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t vd = desc_mnmajor_sw64(smem_u32(smem + L::kV + s * L::kTile), 0);
                        const uint32_t vlo = static_cast<uint32_t>(vd), vhi = static_cast<uint32_t>(vd >> 32);
                        const uint32_t d_addr = tmem_base + (X ? k5TmemOB : k5TmemOA);
                        const uint32_t p_addr = tmem_base + (X ? k5TmemSB : k5TmemSA);
#pragma unroll
                        for (int kk = 0; kk < 8; ++kk)
                            umma_ts_lohi(d_addr, p_addr + kk * 8, vlo + kstep_mnmajor_sw64(kk), vhi, idesc_pv,
                                         (j > 0 || kk > 0) ? 1u : 0u);
                        if (X == nx - 1) umma_commit_addr(bar0 + (L::bVEmpty + s) * 8);
                        if (j == n - 1) umma_commit_addr(bar0 + (L::bOFinal + X) * 8);
                    }
                    __syncwarp();
                };

                mbar_wait(&bars[L::bKFull + gk % NS], (gk / NS) & 1);
                for (int X = 0; X < nx; ++X) {
                    const uint32_t c_ = cntx[X];
                    mbar_wait(&bars[L::bQFull + (c_ & 1) * 2 + X], (c_ >> 1) & 1);
                    tc_fence_after();
                    issue_qk(X, 0);
                }
#pragma unroll 1
                for (int j = 0; j < n; ++j) {
                    mbar_wait(&bars[L::bVFull + (gk + j) % NS], ((gk + j) / NS) & 1);
                    if (j + 1 < n) mbar_wait(&bars[L::bKFull + (gk + j + 1) % NS], ((gk + j + 1) / NS) & 1);
                    for (int X = 0; X < nx; ++X) {
                        issue_pv(X, j);
                        if (j + 1 < n) issue_qk(X, j + 1);
                    }
                }
                gk += n;
                for (int X = 0; X < nx; ++X) { gtx[X] += n; ++cntx[X]; }
            }
      } else if (warp == 9) {
            // ======================================================= bulk-copy producer
            uint32_t gk = 0;
#pragma unroll 1
            for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
                const ItemCoord5 ic = decode_item5(item, npairs, a.H, a.Tq);
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

template <typename TIn, typename TOut, typename LY>
static int launch5_one(const AttnArgs& a, const GtaAttnParams& p, cudaStream_t st) {
    using L = Attn5Cfg<LY::D>;
    auto kern = attn_fwd5_kernel<TIn, TOut, LY>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(L::kBytes));
    if (e != cudaSuccess) return set_error(GTA_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    static thread_local int num_sms = 0, cached_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != cached_dev) {
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (num_sms <= 0) num_sms = 148;
        cached_dev = dev;
    }
    const int npairs = (p.Tq + 255) / 256;
    const long long nitems = static_cast<long long>(p.B) * p.H * npairs;
    if (nitems > 0x7fffffffLL) return set_error(GTA_ERR_UNSUPPORTED, "too many work items");
    const int grid = static_cast<int>(nitems < num_sms ? nitems : num_sms);
    kern<<<grid, kThreads5, L::kBytes, st>>>(a, npairs, static_cast<int>(nitems));
    return check_launch("gta_attn_fwd");
}

template <typename LY>
static int launch5_layout(const GtaAttnParams& p, cudaStream_t st) {
    const AttnArgs a = make_attn_args(p);
    const bool ib = p.in_dtype == GTA_DTYPE_BF16, ob = p.out_dtype == GTA_DTYPE_BF16;
    if (ib && ob) return launch5_one<__nv_bfloat16, __nv_bfloat16, LY>(a, p, st);
    if (ib && !ob) return launch5_one<__nv_bfloat16, float, LY>(a, p, st);
    if (!ib && ob) return launch5_one<float, __nv_bfloat16, LY>(a, p, st);
    return launch5_one<float, float, LY>(a, p, st);
}

}  // namespace gta
