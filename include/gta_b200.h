/*
 * gta_b200 — C ABI of the B200-native geometric-transform-attention (GTA) path.
 *
 * Drop-in boundary for the hot path of autonomousvision/gta (paths relative to the reference root):
 *   gta_attn_fwd          replaces  multihead_geometric_transform_attention  source/utils/gta.py:92-279
 *                         together with AttnFn.forward                         source/layers.py:202-211
 *                         (sole call site: Attention.forward, source/layers.py:422-428)
 *   gta_build_reps        replaces  ImprovedSRTEncoder/Decoder.pre_compute_reps source/encoder.py:183-265,
 *                                                                              source/decoder.py:247-353
 *   gta_so2_mats          replaces  make_SO2mats                               source/utils/gta.py:47-69
 *   gta_wigner_d          replaces  rotmat_to_wigner_d_matrices (l = 1, 2)     source/utils/wigner_d.py:52-58
 *
 * Conventions: plain pointers and sizes, no torch types.  All pointers are DEVICE pointers unless
 * noted; `stream` is a cudaStream_t passed as void*.  Every entry point returns 0 on success or a
 * negative GTA_ERR_* code and never throws; gta_last_error() returns a thread-local message.
 * No ownership is transferred: the caller allocates every buffer (sizes below / gta_attn_fwd_workspace_bytes).
 */
#ifndef GTA_B200_H_
#define GTA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GTA_OK 0
#define GTA_ERR_INVALID (-1)      /* bad argument / unsupported shape */
#define GTA_ERR_CUDA (-2)         /* CUDA runtime error (message has the cudaError string) */
#define GTA_ERR_UNSUPPORTED (-3)  /* valid in the reference but not implemented here (head dim, fp16, backward of the generic path, ...) */

#define GTA_DTYPE_BF16 0
#define GTA_DTYPE_F32 1

/* Packed rep tables (fp32), produced by gta_build_reps or packed from the reference's `extras`:
 *   se3_q [B,Nq,16]  E_q row-major, UNSCALED (= extras['inv_se3rep_q']); the kernels apply
 *                    scale_mask(trans_coeff) on load                      (source/utils/gta.py:135-141)
 *   se3_k [B,Nk,16]  inv(E_k) row-major, unscaled (= extras['se3rep_k'])
 *   so3_q [B,Nq,34]  Wigner D_1 (9) | D_2 (25) of inv(E_q)[:3,:3]  (= extras['so3rep_q'])
 *   so3_k [B,Nk,34]
 *   so2_q [B,Tq,C,2] (cos, sin) of theta[t, j*2+axis], C = 2*nfreqs (= extras['so2rep_q'][...,0,0] / [...,1,0])
 *   so2_k [B,Tk,C,2]
 * Pointers of absent blocks may be NULL. */
typedef struct GtaReps {
    const float* se3_q;
    const float* se3_k;
    const float* so3_q;
    const float* so3_k;
    const float* so2_q;
    const float* so2_k;
    /* ablation blocks (generic path, see GtaAttnParams.t2 / .euclid) */
    const float* se3_qi; /* [B,Nq,16] inv(E_q) unscaled (= extras['se3rep_q']); only read when euclid != 0
                            (source/utils/gta.py:146-156 multiplies the query points by c_q, not by inv_c_q^T) */
    const float* t2_q;   /* [B,Tq,2] patch coordinates (x,y) of make_T2mats, = extras['t2rep_q'][...,2,:2]
                            (source/utils/gta.py:72-89) */
    const float* t2_k;   /* [B,Tk,2] */
} GtaReps;

typedef struct GtaAttnParams {
    /* q [B,H,Tq,D], k,v [B,H,Tk,D]: arbitrary batch/head/token strides (in ELEMENTS), unit stride in D.
     * These are the strided views the reference passes (source/layers.py:394-395). */
    const void* q;
    const void* k;
    const void* v;
    int64_t q_stride_b, q_stride_h, q_stride_t;
    int64_t k_stride_b, k_stride_h, k_stride_t;
    int64_t v_stride_b, v_stride_h, v_stride_t;
    /* out [B,Tq,H,D] contiguous (so 'b h n d -> b n (h d)' at layers.py:429 is a free view). */
    void* out;
    float* lse; /* optional [B,H,Tq] natural-log-sum-exp of the scaled logits; may be NULL */
    int B, H, Tq, Tk, D;
    int Nq, Nk;                /* views; token t belongs to view t / (T/N)  (gta.py:160-162) */
    int triv, se3, so3, so2;   /* f_dims, fixed order triv|se3|so3|so2|t2 (gta.py:115) */
    GtaReps reps;
    const float* trans_coeff;  /* DEVICE pointer to the layer's scalar parameter (layers.py:188-191); NULL => 1.0 */
    float scale;               /* attn_fn.scale / tau  (layers.py:209) */
    int in_dtype, out_dtype;   /* GTA_DTYPE_*.  bf16 in: bf16 tensor-core math, 1e-2 parity budget.  f32 in: split-precision
                                  (bf16 hi + residual, fp32 accumulation), 1e-3 budget, ~3x the tensor work. */
    int v_transform;           /* gta.py:156,168,277 */
    void* workspace;           /* >= gta_attn_fwd_workspace_bytes(...) bytes, 1024-byte aligned */
    size_t workspace_bytes;
    int flags;                 /* GTA_FLAG_* */
    long long* debug_clocks;   /* optional [num_CTAs,16] clock64 phase stamps of the attention kernel (tools/phase_timing*.py); NULL = off */
    int t2;                    /* f_dims['t2']: 3-vectors transformed by the per-token T(2) matrices (gta.py:221-238,272-274) */
    int euclid;                /* euclid_sim: se3 block = homogenised 3-vectors (gta.py:146-156,251-253) and the similarity is
                                  -0.5*|q'-k'|^2 (EuclidAttnFn, source/layers.py:213-224) */
} GtaAttnParams;

/* Which implementation serves a parameter set:
 *   fused path   : t2 == 0, euclid == 0 and every block a multiple of 8 elements (all GTA / GTA-so3 configs): rep
 *                  application fused into the staging pass, the Q stager and the epilogue of the attention kernel.
 *   generic path : anything else the reference accepts (t2, euclid_sim, blocks that are not multiples of 8 such as
 *                  runs/clevrtr/GTA/gta_t2 = triv 2 | se3 32 | t2 30): element-wise rep kernels before and after the same
 *                  tensor-core attention kernel; the euclid similarity is folded into the QK product through two extra
 *                  key columns holding -0.5*|k'|^2 (hi + residual), head dim padded by 32.  Needs the larger workspace
 *                  reported by gta_attn_fwd_workspace_bytes_p. */

#define GTA_FLAG_SKIP_STAGE 2  /* workspace already holds K'/V' of these inputs: launch only the attention kernel */
#define GTA_FLAG_STAGE_ONLY 4  /* launch only the K'/V' staging kernel (fills the workspace) */
#define GTA_FLAG_FAST_FP32 128 /* fp32 inputs: multiply in plain bf16 (1e-2 budget) instead of the split-precision path */
#define GTA_FLAG_V1_PIPELINE 16 /* second generation: two query tiles per CTA, non-persistent (default for D = 128) */
/* Pipeline selection for bf16 / fast-fp32 calls with D <= 96.  With neither bit set the library chooses from the shape
 * (gta_attn_fwd_pipeline reports the choice): ONE launch unless the call is large AND rotation-heavy — 2*Tk/Tq > 1 key
 * tiles to rotate per work item and more than 1e11 attention FLOPs — where the staging warps of the single-launch kernel
 * cannot keep up with the tensor pipe and the stand-alone staging kernel wins by 2 % (measurements in DESIGN.md). */
#define GTA_FLAG_SINGLE_LAUNCH 32 /* force ONE launch for K/V rotation + attention (gta_attn_fwd4.cu): the rotation is done by staging
                                   warps of the persistent attention kernel itself, K'/V' tile images go through L2 with per-tile
                                   ready flags; 0.89x the DRAM traffic of the two-launch pipeline */
#define GTA_FLAG_TWO_LAUNCH 1024 /* force the two-launch pipeline: K'/V' staging kernel + persistent attention kernel */
#define GTA_FLAG_V4_PIPELINE 256 /* two launches with the streaming-softmax / epilogue-warpgroup attention kernel (gta_attn_fwd5.cuh);
                                    head layouts without an instantiation fall back to the gta_attn_fwd3.cu kernel */
#define GTA_FLAG_V5_PIPELINE 512 /* two launches; attention kernel with the spare P buffer in tensor memory (gta_attn_fwd6.cu): QK_X(j+1) is
                                    issued while the exponentials of tile j still run, on every other key tile */
#define GTA_FLAG_RUNTIME_LAYOUT 2048 /* single-launch kernel: use the run-time-layout staging code even for a layout that has a
                                       compile-time-specialised instantiation (A/B measurement, tests) */
#define GTA_FLAG_BWD_SPLIT 4096 /* gta_attn_bwd: keep the dK/dV kernel + dQ kernel pair (S and dP computed twice) instead of the fused
                                  * kernel that accumulates the dQ partial sums with bulk reductions (head dims <= 96; calls of fewer than 148 key tiles
                                  * use the pair by default, GTA_FLAG_SINGLE_LAUNCH forces the fused kernel) */
#define GTA_FLAG_V3_PRESTAGED 64 /* with GTA_FLAG_SKIP_STAGE: run the single-launch kernel on an already staged workspace
                                   (its rotation warps idle) instead of the two-launch attention kernel */

/* Scratch for the rotated K'/V' operand tiles (bf16 inputs, or fp32 inputs with GTA_FLAG_FAST_FP32) and the per-tile
 * ready flags of the single-launch kernel. */
size_t gta_attn_fwd_workspace_bytes(int B, int H, int Tk, int D);
/* Same for any dtype/flags: fp32 inputs use split-precision (hi + residual) tile images, twice the size. */
size_t gta_attn_fwd_workspace_bytes_ex(int B, int H, int Tk, int D, int in_dtype, int flags);

/* Workspace for exactly this parameter set (covers the generic path; workspace fields of *p are ignored). */
size_t gta_attn_fwd_workspace_bytes_p(const GtaAttnParams* p);

/* Fused forward: O = rho_q^{-1} softmax((rho_q^{-T} Q)(rho_k K)^T * scale) (rho_k V). */
int gta_attn_fwd(const GtaAttnParams* p, void* stream);

/* Which pipeline gta_attn_fwd runs for *p (workspace fields are ignored). */
#define GTA_PIPELINE_TWO_LAUNCH 0      /* rotate_kv_kernel + attention kernel */
#define GTA_PIPELINE_SINGLE_LAUNCH 1   /* attn_fwd4_kernel: rotation + attention in one launch */
#define GTA_PIPELINE_SPLIT_PRECISION 2 /* fp32 inputs: staging + attn_fwd_hp_kernel */
#define GTA_PIPELINE_GENERIC 3         /* t2 / euclid_sim / unaligned blocks: element-wise rep kernels around the attention */
int gta_attn_fwd_pipeline(const GtaAttnParams* p);

/* Backward of gta_attn_fwd (what torch.autograd derives from source/utils/gta.py:92-279 + source/layers.py:207-211 in
 * the reference's training step, source/trainer.py:69-83): gradients w.r.t. q, k, v and the layer's trans_coeff
 * (source/layers.py:188-191; the SO(3) reps are detached, gta.py:194-197, the SO(2)/SE(3) matrices are data).
 *   fwd          the forward call's parameters: q/k/v (+strides), dims, reps, trans_coeff, scale, dtypes, v_transform;
 *                fwd.out = the forward OUTPUT [B,Tq,H,D] and fwd.lse = its log-sum-exp [B,H,Tq] (both required);
 *                fwd.workspace / flags are ignored.  Fused-path configurations only (no t2 / euclid).
 *   dout         [B,Tq,H,D] contiguous, dtype of out (= dtype of q/k/v)
 *   dq, dk, dv   [B,Tq,H,D] / [B,Tk,H,D] contiguous, dtype of q/k/v (written)
 *   dtrans_coeff optional device scalar, ACCUMULATED into (atomicAdd; zero it first)
 * Tensor-core math is bf16 with fp32 accumulation for both input dtypes. */
typedef struct GtaAttnBwdParams {
    GtaAttnParams fwd;
    const void* dout;
    void* dq;
    void* dk;
    void* dv;
    float* dtrans_coeff;
    void* workspace;          /* >= gta_attn_bwd_workspace_bytes(...), 1024-byte aligned */
    size_t workspace_bytes;
} GtaAttnBwdParams;

size_t gta_attn_bwd_workspace_bytes(int B, int H, int Tq, int Tk, int D);
/* ... of the call described by p->fwd-style parameters: also covers the generic-path layouts (t2 block, blocks that are not
 * multiples of 8), whose backward runs the element-wise rep passes around the tensor-core backward on dense operands. */
size_t gta_attn_bwd_workspace_bytes_p(const GtaAttnParams* p);
int gta_attn_bwd(const GtaAttnBwdParams* p, void* stream);

/* The attention map the reference returns as its second output (source/layers.py:207-211 `attn`, consumed only under
 * return_last_attmap, source/layers.py:478-480 / SURVEY T7): attn[b,h,i,j] = exp(q'_i . k'_j * scale - lse[b,h,i]), fp32
 * [B,H,Tq,Tk], from the same parameters as a previous gta_attn_fwd call whose p->lse was requested.  Materialises
 * B*H*Tq*Tk floats — a visualisation path (fp32 SIMT dot products), not a hot path.  Not defined for euclid_sim.
 * Workspace: gta_attn_probs_workspace_bytes (dense fp32 q', k'). */
size_t gta_attn_probs_workspace_bytes(int B, int H, int Tq, int Tk, int D);
int gta_attn_probs(const GtaAttnParams* p, float* attn, void* stream);

/* Rotated operands only (q' = rho_q^{-T} q etc.), fp32 [B,H,T,D] contiguous; testing / inspection. */
int gta_rotate_debug(const GtaAttnParams* p, float* qt, float* kt, float* vt, void* stream);

/* extrinsics [B,N,4,4] fp32 row-major, coords [B,T,2] fp32 -> packed tables (see GtaReps).
 * so3_maxdeg: 0 (skip) or 2.  Output pointers of skipped blocks may be NULL. */
int gta_build_reps(const float* extr_q, const float* extr_k, const float* coord_q, const float* coord_k,
                   int B, int Nq, int Nk, int Tq, int Tk, int so2_nfreqs, float max_freq_h, float max_freq_w,
                   int shared_freqs, int so3_maxdeg, float* se3_q, float* se3_k, float* so3_q, float* so3_k,
                   float* so2_q, float* so2_k, void* stream);

/* extr [n,4,4] -> inv [n,4,4] (general inverse in fp64, torch.linalg.inv at source/encoder.py:219). */
int gta_se3_inverse(const float* extr, int64_t n, float* inv, void* stream);

/* make_T2mats (source/utils/gta.py:72-89): coord [n,2] -> mats [n,3,3] = [[1,0,0],[0,1,0],[x,y,1]] and, when
 * inv_mats != NULL, their inverses [[1,0,0],[0,1,0],[-x,-y,1]] (torch.linalg.inv at source/encoder.py:212). */
int gta_t2_mats(const float* coord, int64_t n, float* mats, float* inv_mats, void* stream);

/* coord [n,2] -> mats [n, 2*nfreqs, 2, 2] in the reference's layout (freq-major, axis-minor pairs). */
int gta_so2_mats(const float* coord, int64_t n, int nfreqs, float max_freq_h, float max_freq_w, int shared_freqs,
                 float* mats, void* stream);

/* R [n,3,3] -> d1 [n,3,3], d2 [n,5,5] (ZYZ Euler angles with gimbal handling, D_l = Z J Z J Z). */
int gta_wigner_d(const float* R, int64_t n, float* d1, float* d2, void* stream);

const char* gta_last_error(void);
int gta_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* GTA_B200_H_ */
