// Shared by the backward kernels (gta_attn_bwd.cu: dK/dV kernel + dQ kernel; gta_attn_bwd2.cu: the fused kernel).
#pragma once
#include "attn_common.cuh"

namespace gta {

struct BwdArgs {
    const uint8_t* q_img; const uint8_t* do_img;     // [B*H*ntq] tile images of Q', dO'
    const uint8_t* k_img; const uint8_t* v_img;      // [B*H*ntk] tile images of K', V'
    const float* lse; const float* delta;            // [B,H,Tq]
    void* dq; void* dk; void* dv;                    // [B,T,H,D] contiguous
    const void* q; const void* k; const void* v;     // raw inputs (trans_coeff terms)
    int64_t q_sb, q_sh, q_st, k_sb, k_sh, k_st, v_sb, v_sh, v_st;
    float* dtc;
    int B, H, Tq, Tk, Nq, Nk, tpvq, tpvk, ntq, ntk, C;
    HeadDims hd;
    const float* se3_q; const float* so3_q; const float* so2_q;
    const float* se3_k; const float* so3_k; const float* so2_k;
    const float* tc_ptr;
    float scale, scale_log2;
    int v_transform;
    float* dq_acc;       // fused kernel: fp32 dQ' accumulation tiles [B*H*ntq][D/4][128][4], zeroed by the launcher
    long long* dbg;      // optional [2 kernels][num CTAs][16] clock64 phase sums (tools/bwd_phase_timing.py)
};

__device__ __forceinline__ void bwd_bar_sync(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Fused backward (gta_attn_bwd2.cu): one kernel for dK, dV and the dQ' partial sums + the kernel that finishes dQ.
size_t bwd_dq_acc_bytes(int B, int H, int Tq, int D);
bool bwd_fused_supported(int D);
int launch_bwd_fused(const BwdArgs& a, bool bf16, int D, bool runtime_layout, cudaStream_t st);

}  // namespace gta
