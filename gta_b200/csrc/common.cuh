// Shared host-side helpers of the C-ABI library (error reporting, launch checks) + internal launchers.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/gta_b200.h"

namespace gta {

int set_error(int code, const char* fmt, ...);
int check_launch(const char* what);

int launch_build_reps(const float* extr_q, const float* extr_k, const float* coord_q, const float* coord_k, int B,
                      int Nq, int Nk, int Tq, int Tk, int so2_nfreqs, float mfh, float mfw, int shared,
                      int so3_maxdeg, float* se3_q, float* se3_k, float* so3_q, float* so3_k, float* so2_q,
                      float* so2_k, cudaStream_t st);
int launch_so2_mats(const float* coord, int64_t n, int nfreqs, float mfh, float mfw, int shared, float* mats,
                    cudaStream_t st);
int launch_wigner(const float* R, int64_t n, float* d1, float* d2, cudaStream_t st);
int launch_se3_inverse(const float* extr, int64_t n, float* inv, cudaStream_t st);
int launch_t2_mats(const float* coord, int64_t n, float* mats, float* inv, cudaStream_t st);

size_t attn_bwd_workspace_bytes(int B, int H, int Tq, int Tk, int D);
// delta_dout: the gradient the delta kernel pairs with fwd.out (the generic path stages a rotated dout but takes delta from
// the un-rotated pair); nullptr = bp.dout
int launch_attn_bwd(const GtaAttnBwdParams& bp, cudaStream_t st, const void* delta_dout = nullptr);
size_t generic_bwd_workspace_bytes(const GtaAttnParams& p);
int launch_attn_bwd_generic(const GtaAttnBwdParams& bp, cudaStream_t st);

// Generic path (gta_generic.cu): t2 block, euclid_sim, head layouts with blocks that are not multiples of 8.
bool attn_needs_generic(const GtaAttnParams& p);
size_t generic_workspace_bytes(const GtaAttnParams& p);
int launch_attn_fwd_generic(const GtaAttnParams& p, cudaStream_t st);
size_t attn_probs_workspace_bytes(int B, int H, int Tq, int Tk, int D);
int launch_attn_probs(const GtaAttnParams& p, float* attn, cudaStream_t st);
int launch_rotate_debug_generic(const GtaAttnParams& p, float* qt, float* kt, float* vt, cudaStream_t st);

int validate_attn_params(const GtaAttnParams* p);
int launch_rotate_kv(const GtaAttnParams& p, cudaStream_t st);
int launch_rotate_q_do(const GtaAttnParams& p, const void* dout, uint8_t* q_img, uint8_t* do_img, cudaStream_t st);
int launch_rotate_debug(const GtaAttnParams& p, float* qt, float* kt, float* vt, cudaStream_t st);
int launch_attn_fwd(const GtaAttnParams& p, cudaStream_t st);
int launch_attn_fwd_v0(const GtaAttnParams& p, cudaStream_t st);   // dev library only (gta_dev_abi.cu)
int launch_attn_fwd_v2(const GtaAttnParams& p, cudaStream_t st);
int launch_attn_fwd_v3(const GtaAttnParams& p, bool fused, cudaStream_t st);
int launch_attn_fwd_v4(const GtaAttnParams& p, cudaStream_t st, bool* handled);
int launch_attn_fwd_v5(const GtaAttnParams& p, cudaStream_t st);
int launch_softmax_bench(int num, int den, int warps, int reps, int grid, const float* in, float* out, long long* clk,
                         cudaStream_t st);
int launch_umma_bench(int D, int mode, int reps, int grid, long long* out, cudaStream_t st);
int launch_umma_probe(const void* A, const void* Bm, const void* P, const void* V, int D, int p_in_tmem, float* outS,
                      float* outO, cudaStream_t st);

// fp32 inputs run the split-precision pipeline (bf16 hi + bf16 residual operands, 3 MMAs per product term) unless the
// caller asks for plain bf16 math with GTA_FLAG_FAST_FP32.
inline bool attn_is_split_precision(const GtaAttnParams& p) {
    return p.in_dtype == GTA_DTYPE_F32 && !(p.flags & GTA_FLAG_FAST_FP32);
}
int launch_attn_fwd_hp(const GtaAttnParams& p, cudaStream_t st);

// Scratch layout: [K' tiles | V' tiles], each tile image = D/32 column blocks x 128 rows x 64 B (bf16).
inline int num_kv_tiles(int Tk) { return (Tk + 127) / 128; }
inline size_t kv_tile_bytes(int D) { return static_cast<size_t>(128) * D * 2; }
// The fused single-launch kernel keeps one ready flag (int) per (batch, head, key tile) behind the tile images.
inline size_t kv_flags_offset(int B, int H, int Tk, int D) { return 2 * static_cast<size_t>(B) * H * num_kv_tiles(Tk) * kv_tile_bytes(D); }
inline size_t kv_flags_bytes(int B, int H, int Tk) { return (static_cast<size_t>(B) * H * num_kv_tiles(Tk) * sizeof(int) + 1023) / 1024 * 1024; }
// Which parameter sets the fused single-launch kernel (gta_attn_fwd4.cu) serves.
inline bool attn_is_fused_launch(const GtaAttnParams& p) {
    if (attn_is_split_precision(p) || p.D > 96) return false;
    if (p.flags & (GTA_FLAG_SKIP_STAGE | GTA_FLAG_STAGE_ONLY | GTA_FLAG_V1_PIPELINE | GTA_FLAG_V4_PIPELINE | GTA_FLAG_V5_PIPELINE |
                   GTA_FLAG_TWO_LAUNCH))
        return false;
    if (p.flags & GTA_FLAG_SINGLE_LAUNCH) return true;
    // automatic choice (measured on B200): the staging warps of the single-launch kernel rotate 2*Tk/Tq key tiles
    // per work item; large calls that need more than one per item run 2 % faster with the stand-alone staging kernel
    const double flops = 4.0 * p.B * p.H * static_cast<double>(p.Tq) * p.Tk * p.D;
    return !(2.0 * p.Tk > 1.0 * p.Tq && flops > 1e11);
}

}  // namespace gta
