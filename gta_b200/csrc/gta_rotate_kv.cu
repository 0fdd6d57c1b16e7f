// K'/V' staging: K' = rho_k K, V' = rho_k V are rotated ONCE per (batch, head) and written as bf16 UMMA
// operand tile images (128 keys x D, 64-byte swizzle — see tile_sw64_offset) so that the attention kernel can
// fetch a whole tile with one bulk async copy and feed it to tcgen05.mma without touching it again.
// Reference semantics: source/utils/gta.py:160-168 (se3), :182-198 (so3), :203-218 (so2).
//
// HBM-bound by design: algorithmic bytes = (in_bytes + 2) * 2 * B*H*Tk*D; see DESIGN.md.
#include "rotate_tile.cuh"

namespace gta {

#ifndef GTA_ROT_MINB
#define GTA_ROT_MINB 1
#endif
template <typename T, bool kQSide>
__global__ void __launch_bounds__(128, GTA_ROT_MINB) rotate_kv_kernel(const RotArgs a) {
    const float tc = a.tc_ptr ? __ldg(a.tc_ptr) : 1.0f;
    rotate_tile<T, kQSide>(a, blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x >> 5, threadIdx.x & 31, tc);
}

// ---------------------------------------------------------------------------------------------------
// Debug / inspection: rotated q', k', v' as fp32 [B,H,T,D].
struct RotDbgArgs {
    const void* x; int64_t sb, sh, st;
    float* out;
    int B, H, T, D, N, tpv, C;
    HeadDims hd;
    const float* se3; const float* so3; const float* so2; const float* tc_ptr;
    int mode;      // RepMode
    int rotate;
};

template <typename T>
__global__ void rotate_debug_kernel(const RotDbgArgs a) {
    const int nch = a.D >> 3;
    int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t total = static_cast<int64_t>(a.B) * a.H * a.T * nch;
    if (i >= total) return;
    const int c = static_cast<int>(i % nch);
    int64_t r = i / nch;
    const int t = static_cast<int>(r % a.T); r /= a.T;
    const int h = static_cast<int>(r % a.H);
    const int b = static_cast<int>(r / a.H);
    const float tc = a.tc_ptr ? __ldg(a.tc_ptr) : 1.0f;
    float x[8];
    load_chunk<T>(reinterpret_cast<const T*>(a.x) + b * a.sb + h * a.sh + t * a.st + c * 8, x);
    if (a.rotate) {
        const size_t view = static_cast<size_t>(b) * a.N + t / a.tpv;
        const float* se3 = a.se3 + view * 16;
        const float* so3 = a.so3 + view * 34;
        const float* so2 = a.so2 + (static_cast<size_t>(b) * a.T + t) * a.C * 2;
        if (a.mode == kModeQ) apply_rep_chunk<kModeQ>(x, c, a.hd, se3, so3, so2, tc);
        else apply_rep_chunk<kModeKV>(x, c, a.hd, se3, so3, so2, tc);
    }
    store_chunk<float>(a.out + ((static_cast<int64_t>(b) * a.H + h) * a.T + t) * a.D + c * 8, x);
}

int launch_rotate_kv(const GtaAttnParams& p, cudaStream_t st) {
    const RotArgs a = make_rot_args_kv(p);
    dim3 grid(a.ntiles, p.H, p.B);
    if (p.in_dtype == GTA_DTYPE_BF16) rotate_kv_kernel<__nv_bfloat16, false><<<grid, 128, 0, st>>>(a);
    else rotate_kv_kernel<float, false><<<grid, 128, 0, st>>>(a);
    return check_launch("gta_rotate_kv");
}

// Backward staging of the query side: Q' and dO' tile images (bf16) from q (strided) and dout [B,Tq,H,D].
int launch_rotate_q_do(const GtaAttnParams& p, const void* dout, uint8_t* q_img, uint8_t* do_img, cudaStream_t st) {
    RotArgs a;
    a.k = p.q; a.v = dout;
    a.k_sb = p.q_stride_b; a.k_sh = p.q_stride_h; a.k_st = p.q_stride_t;
    a.v_sb = static_cast<int64_t>(p.Tq) * p.H * p.D; a.v_sh = p.D; a.v_st = static_cast<int64_t>(p.H) * p.D;
    a.ntiles = num_kv_tiles(p.Tq);
    a.ws_k = q_img; a.ws_v = do_img; a.lo_offset = 0;
    a.H = p.H; a.Tk = p.Tq; a.D = p.D; a.Nk = p.Nq; a.tpv = p.Tq / p.Nq;
    a.hd = HeadDims{p.triv, p.se3, p.so3, p.so2};
    a.se3_k = p.reps.se3_q; a.so3_k = p.reps.so3_q; a.so2_k = p.reps.so2_q;
    a.tc_ptr = p.trans_coeff; a.C = p.so2 >> 1; a.v_transform = p.v_transform;
    dim3 grid(a.ntiles, p.H, p.B);
    if (p.in_dtype == GTA_DTYPE_BF16) rotate_kv_kernel<__nv_bfloat16, true><<<grid, 128, 0, st>>>(a);
    else rotate_kv_kernel<float, true><<<grid, 128, 0, st>>>(a);
    return check_launch("gta_attn_bwd (Q'/dO' staging)");
}

int launch_rotate_debug(const GtaAttnParams& p, float* qt, float* kt, float* vt, cudaStream_t st) {
    for (int which = 0; which < 3; ++which) {
        float* out = which == 0 ? qt : (which == 1 ? kt : vt);
        if (!out) continue;
        RotDbgArgs a;
        a.x = which == 0 ? p.q : (which == 1 ? p.k : p.v);
        a.sb = which == 0 ? p.q_stride_b : (which == 1 ? p.k_stride_b : p.v_stride_b);
        a.sh = which == 0 ? p.q_stride_h : (which == 1 ? p.k_stride_h : p.v_stride_h);
        a.st = which == 0 ? p.q_stride_t : (which == 1 ? p.k_stride_t : p.v_stride_t);
        a.out = out; a.B = p.B; a.H = p.H; a.D = p.D;
        a.T = which == 0 ? p.Tq : p.Tk; a.N = which == 0 ? p.Nq : p.Nk; a.tpv = a.T / a.N;
        a.C = p.so2 >> 1; a.hd = HeadDims{p.triv, p.se3, p.so3, p.so2};
        a.se3 = which == 0 ? p.reps.se3_q : p.reps.se3_k;
        a.so3 = which == 0 ? p.reps.so3_q : p.reps.so3_k;
        a.so2 = which == 0 ? p.reps.so2_q : p.reps.so2_k;
        a.tc_ptr = p.trans_coeff; a.mode = which == 0 ? kModeQ : kModeKV;
        a.rotate = (which < 2) || p.v_transform;
        int64_t total = static_cast<int64_t>(a.B) * a.H * a.T * (a.D >> 3);
        unsigned blocks = static_cast<unsigned>((total + 255) / 256);
        if (p.in_dtype == GTA_DTYPE_BF16) rotate_debug_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(a);
        else rotate_debug_kernel<float><<<blocks, 256, 0, st>>>(a);
    }
    return check_launch("gta_rotate_debug");
}

}  // namespace gta
