// extern "C" entry points of libgta_b200.so (declared in include/gta_b200.h).  Plain pointers and sizes only.
#include <cstring>

#include "common.cuh"

namespace gta {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(GTA_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return GTA_OK;
}

int validate_attn_params(const GtaAttnParams* p) {
    if (!p) return set_error(GTA_ERR_INVALID, "null params");
    if (!p->q || !p->k || !p->v || !p->out) return set_error(GTA_ERR_INVALID, "null q/k/v/out pointer");
    if (p->B <= 0 || p->H <= 0 || p->Tq <= 0 || p->Tk <= 0) return set_error(GTA_ERR_INVALID, "empty B/H/Tq/Tk");
    if (p->D != 32 && p->D != 64 && p->D != 96 && p->D != 128)
        return set_error(GTA_ERR_UNSUPPORTED, "head dim %d not in {32,64,96,128}", p->D);
    if (p->triv < 0 || p->se3 < 0 || p->so3 < 0 || p->so2 < 0 || p->t2 < 0 ||
        p->triv + p->se3 + p->so3 + p->so2 + p->t2 != p->D)
        return set_error(GTA_ERR_INVALID, "f_dims (%d,%d,%d,%d,%d) must sum to head dim %d", p->triv, p->se3, p->so3,
                         p->so2, p->t2, p->D);
    if (p->se3 % (p->euclid ? 3 : 4) || p->so3 % 8 || p->so2 % 2 || p->t2 % 3)
        return set_error(GTA_ERR_INVALID, "f_dims: se3 must hold whole %d-vectors, so3 whole [3|5] groups, so2 pairs, t2 triples",
                         p->euclid ? 3 : 4);
    if (p->euclid && p->D > 96) return set_error(GTA_ERR_UNSUPPORTED, "euclid_sim needs head dim <= 96");
    if (p->t2 && (!p->reps.t2_q || !p->reps.t2_k)) return set_error(GTA_ERR_INVALID, "t2 block without t2 coordinates");
    if (p->euclid && p->se3 && !p->reps.se3_qi) return set_error(GTA_ERR_INVALID, "euclid_sim needs reps.se3_qi = inv(E_q)");
    if (p->Nq <= 0 || p->Nk <= 0 || p->Tq % p->Nq || p->Tk % p->Nk)
        return set_error(GTA_ERR_INVALID, "Tq/Tk must be divisible by the number of views");
    if (p->se3 && (!p->reps.se3_q || !p->reps.se3_k)) return set_error(GTA_ERR_INVALID, "se3 block without se3 reps");
    if (p->so3 && (!p->reps.so3_q || !p->reps.so3_k)) return set_error(GTA_ERR_INVALID, "so3 block without so3 reps");
    if (p->so2 && (!p->reps.so2_q || !p->reps.so2_k)) return set_error(GTA_ERR_INVALID, "so2 block without so2 reps");
    if ((p->in_dtype != GTA_DTYPE_BF16 && p->in_dtype != GTA_DTYPE_F32) ||
        (p->out_dtype != GTA_DTYPE_BF16 && p->out_dtype != GTA_DTYPE_F32))
        return set_error(GTA_ERR_UNSUPPORTED, "dtype must be bf16 or f32");
    const int64_t align = p->in_dtype == GTA_DTYPE_BF16 ? 8 : 4;  // 16-byte vector loads
    const int64_t strides[9] = {p->q_stride_b, p->q_stride_h, p->q_stride_t, p->k_stride_b, p->k_stride_h,
                                p->k_stride_t, p->v_stride_b, p->v_stride_h, p->v_stride_t};
    for (int i = 0; i < 9; ++i)
        if (strides[i] % align) return set_error(GTA_ERR_UNSUPPORTED, "q/k/v strides must keep rows 16-byte aligned");
    const uintptr_t ptrs[4] = {reinterpret_cast<uintptr_t>(p->q), reinterpret_cast<uintptr_t>(p->k),
                               reinterpret_cast<uintptr_t>(p->v), reinterpret_cast<uintptr_t>(p->out)};
    for (int i = 0; i < 4; ++i)
        if (ptrs[i] & 15) return set_error(GTA_ERR_UNSUPPORTED, "q/k/v/out must be 16-byte aligned");
    return GTA_OK;
}

}  // namespace gta

using namespace gta;

extern "C" {

const char* gta_last_error(void) { return g_err; }
// 5: gta_attn_bwd_workspace_bytes_p, backward of the generic-path layouts, fused backward kernel (larger backward workspace)
int gta_abi_version(void) { return 5; }

size_t gta_attn_fwd_workspace_bytes(int B, int H, int Tk, int D) {
    if (B <= 0 || H <= 0 || Tk <= 0 || D <= 0) return 0;
    return kv_flags_offset(B, H, Tk, D) + kv_flags_bytes(B, H, Tk);
}

size_t gta_attn_fwd_workspace_bytes_ex(int B, int H, int Tk, int D, int in_dtype, int flags) {
    if (B <= 0 || H <= 0 || Tk <= 0 || D <= 0) return 0;
    const size_t tiles = kv_flags_offset(B, H, Tk, D);
    return ((in_dtype == GTA_DTYPE_F32 && !(flags & GTA_FLAG_FAST_FP32)) ? 2 * tiles : tiles) + kv_flags_bytes(B, H, Tk);
}

size_t gta_attn_fwd_workspace_bytes_p(const GtaAttnParams* p) {
    if (!p || p->B <= 0 || p->H <= 0 || p->Tq <= 0 || p->Tk <= 0 || p->D <= 0) return 0;
    if (attn_needs_generic(*p)) return generic_workspace_bytes(*p);
    return gta_attn_fwd_workspace_bytes_ex(p->B, p->H, p->Tk, p->D, p->in_dtype, p->flags);
}

int gta_attn_fwd_pipeline(const GtaAttnParams* p) {
    if (!p) return GTA_ERR_INVALID;
    if (attn_needs_generic(*p)) return GTA_PIPELINE_GENERIC;
    if (attn_is_split_precision(*p)) return GTA_PIPELINE_SPLIT_PRECISION;
    return attn_is_fused_launch(*p) ? GTA_PIPELINE_SINGLE_LAUNCH : GTA_PIPELINE_TWO_LAUNCH;
}

int gta_attn_fwd(const GtaAttnParams* p, void* stream) {
    int rc = validate_attn_params(p);
    if (rc) return rc;
    const size_t need = gta_attn_fwd_workspace_bytes_p(p);
    if (!p->workspace || p->workspace_bytes < need)
        return set_error(GTA_ERR_INVALID, "workspace too small (need %zu bytes)", need);
    if (reinterpret_cast<uintptr_t>(p->workspace) & 1023) return set_error(GTA_ERR_INVALID, "workspace must be 1024-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (attn_needs_generic(*p)) return launch_attn_fwd_generic(*p, st);
    if (attn_is_fused_launch(*p)) return launch_attn_fwd_v3(*p, true, st);           // one launch: rotation + attention
    if ((p->flags & GTA_FLAG_SKIP_STAGE) && (p->flags & GTA_FLAG_V3_PRESTAGED) && !attn_is_split_precision(*p) && p->D <= 96)
        return launch_attn_fwd_v3(*p, false, st);
    if (!(p->flags & GTA_FLAG_SKIP_STAGE)) {
        rc = launch_rotate_kv(*p, st);
        if (rc) return rc;
    }
    if (p->flags & GTA_FLAG_STAGE_ONLY) return GTA_OK;
    if (attn_is_split_precision(*p)) return launch_attn_fwd_hp(*p, st);
    return launch_attn_fwd(*p, st);
}

size_t gta_attn_bwd_workspace_bytes(int B, int H, int Tq, int Tk, int D) { return attn_bwd_workspace_bytes(B, H, Tq, Tk, D); }
size_t gta_attn_bwd_workspace_bytes_p(const GtaAttnParams* p) {
    if (!p) return 0;
    return attn_needs_generic(*p) ? generic_bwd_workspace_bytes(*p) : attn_bwd_workspace_bytes(p->B, p->H, p->Tq, p->Tk, p->D);
}

int gta_attn_bwd(const GtaAttnBwdParams* p, void* stream) {
    if (!p) return set_error(GTA_ERR_INVALID, "null params");
    int rc = validate_attn_params(&p->fwd);
    if (rc) return rc;
    return launch_attn_bwd(*p, static_cast<cudaStream_t>(stream));
}

size_t gta_attn_probs_workspace_bytes(int B, int H, int Tq, int Tk, int D) { return attn_probs_workspace_bytes(B, H, Tq, Tk, D); }

int gta_attn_probs(const GtaAttnParams* p, float* attn, void* stream) {
    int rc = validate_attn_params(p);
    if (rc) return rc;
    return launch_attn_probs(*p, attn, static_cast<cudaStream_t>(stream));
}

int gta_rotate_debug(const GtaAttnParams* p, float* qt, float* kt, float* vt, void* stream) {
    int rc = validate_attn_params(p);
    if (rc) return rc;
    if (attn_needs_generic(*p)) return launch_rotate_debug_generic(*p, qt, kt, vt, static_cast<cudaStream_t>(stream));
    return launch_rotate_debug(*p, qt, kt, vt, static_cast<cudaStream_t>(stream));
}

int gta_build_reps(const float* extr_q, const float* extr_k, const float* coord_q, const float* coord_k, int B, int Nq,
                   int Nk, int Tq, int Tk, int so2_nfreqs, float max_freq_h, float max_freq_w, int shared_freqs,
                   int so3_maxdeg, float* se3_q, float* se3_k, float* so3_q, float* so3_k, float* so2_q, float* so2_k,
                   void* stream) {
    return launch_build_reps(extr_q, extr_k, coord_q, coord_k, B, Nq, Nk, Tq, Tk, so2_nfreqs, max_freq_h, max_freq_w,
                             shared_freqs, so3_maxdeg, se3_q, se3_k, so3_q, so3_k, so2_q, so2_k,
                             static_cast<cudaStream_t>(stream));
}

int gta_so2_mats(const float* coord, int64_t n, int nfreqs, float max_freq_h, float max_freq_w, int shared_freqs,
                 float* mats, void* stream) {
    return launch_so2_mats(coord, n, nfreqs, max_freq_h, max_freq_w, shared_freqs, mats, static_cast<cudaStream_t>(stream));
}

int gta_se3_inverse(const float* extr, int64_t n, float* inv, void* stream) {
    return launch_se3_inverse(extr, n, inv, static_cast<cudaStream_t>(stream));
}

int gta_t2_mats(const float* coord, int64_t n, float* mats, float* inv_mats, void* stream) {
    return launch_t2_mats(coord, n, mats, inv_mats, static_cast<cudaStream_t>(stream));
}

int gta_wigner_d(const float* R, int64_t n, float* d1, float* d2, void* stream) {
    return launch_wigner(R, n, d1, d2, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
