// Generic rep application for the configurations the fused kernels do not cover: the T(2) block (`t2`), the euclid
// similarity (`euclid_sim`) and head layouts whose blocks are not multiples of 8 elements (e.g. runs/clevrtr/GTA/gta_t2:
// triv 2 | se3 32 | t2 30).  One thread per OUTPUT ELEMENT evaluates its row of the block-diagonal rep
//   q' = rho_q^{-T} q,  k' = rho_k k,  v' = rho_k v   (source/utils/gta.py:127-242)
// into dense, padded operand tensors; the tensor-core attention kernel then runs on them with an all-trivial head
// layout, and a second element-wise pass applies rho_q^{-1} to its fp32 output (gta.py:246-276).
//
// euclid_sim (EuclidAttnFn, source/layers.py:213-224): sim = q'.k' - |q'|^2/2 - |k'|^2/2.  The -|q'|^2/2 term is constant
// along the softmax axis and cancels; -|k'|^2/2 is folded into the QK product through two extra operand columns
// (k'[D] = hi, k'[D+1] = residual of -|k'|^2/2, q'[D] = q'[D+1] = 1), the head dim being padded by 32.
#include <algorithm>

#include "common.cuh"
#include "reps.cuh"

namespace gta {

struct GenDims {
    int triv, se3, so3, so2, t2, euclid;
};

struct GenArgs {
    const void* x;             // modes Q/KV: [B,H,T,D] strided; mode Out: fp32 [B,T,H,Da] contiguous
    int64_t sb, sh, st;
    void* out;                 // modes Q/KV: [B,H,T,Da] contiguous (TStore); mode Out: [B,T,H,D] (TStore)
    int B, H, T, D, Da, N, tpv, C;
    GenDims gd;
    const float* se3m;         // [B,N,16]: Q: E_q (euclid: inv E_q); KV: inv E_k; Out: E_q
    const float* so3m;         // [B,N,34]
    const float* so2cs;        // [B,T,C,2]
    const float* xy;           // [B,T,2]
    const float* tc_ptr;
    int mode;                  // RepMode
    int rotate;                // 0: copy (v_transform = False)
    int ones;                  // euclid, query side: columns D, D+1 = 1
    int out_bthd;              // modes Q/KV: write [B,T,H,Da] instead of [B,H,T,Da] (backward: dO' in the layout of dout)
    int se3_lin_t;             // euclid backward: the SE(3) block applies the TRANSPOSED 3x3 linear part of se3m (no translation)
    int sub_trans;             // euclid backward: SE(3) block only, y = x - tc * translation column of se3m (every other element copied)
    const void* bias_k;        // euclid backward, mode KVT: dense k' [B,H,T,Da]; the loaded gradient row becomes g[e] - g[D] * k'[e]
};

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// One output element e of the rep applied to a head row; ld(i) reads input element i of the same row.
template <typename Ld>
__device__ float gen_element(const GenArgs& a, int e, size_t view, size_t tok, float tc, Ld ld) {
    const GenDims& g = a.gd;
    const int mode = a.mode;
    if (e < g.triv) return ld(e);
    int e1 = e - g.triv;
    if (a.sub_trans && e1 >= g.se3) return ld(e);
    if (e1 < g.se3) {
        const float* M = a.se3m + view * 16;
        if (!g.euclid) {                       // 4-vectors, (M * scale_mask(tc)) or its transpose (gta.py:160-167,255-257)
            const int base = e - (e1 & 3), i = e1 & 3;
            const float x0 = ld(base), x1 = ld(base + 1), x2 = ld(base + 2), x3 = ld(base + 3);
            if (mode == kModeQ || mode == kModeKVT) {        // transpose of (M * scale_mask(tc))
                if (i < 3) return fmaf(M[i], x0, fmaf(M[4 + i], x1, M[8 + i] * x2));
                return fmaf(tc, fmaf(M[3], x0, fmaf(M[7], x1, M[11] * x2)), M[15] * x3);
            }
            if (i < 3) return fmaf(M[4 * i], x0, fmaf(M[4 * i + 1], x1, fmaf(M[4 * i + 2], x2, M[4 * i + 3] * tc * x3)));
            return M[15] * x3;
        }
        // euclid: homogenised 3-vectors, no transpose on the query side (gta.py:146-156,251-253): y = A x + t * tc
        const int i = e1 % 3, base = e - i;
        if (a.sub_trans) return ld(e) - M[4 * i + 3] * tc;
        const float x0 = ld(base), x1 = ld(base + 1), x2 = ld(base + 2);
        if (a.se3_lin_t) return fmaf(M[i], x0, fmaf(M[4 + i], x1, M[8 + i] * x2));       // backward: A^T g
        return fmaf(M[4 * i], x0, fmaf(M[4 * i + 1], x1, fmaf(M[4 * i + 2], x2, M[4 * i + 3] * tc)));
    }
    int e2 = e1 - g.se3;
    if (e2 < g.so3) {                          // [3 | 5] groups, Wigner D_1 / D_2 (gta.py:182-201,259-268)
        const float* W = a.so3m + view * 34;
        const int i = e2 & 7, base = e - i;
        const bool tr = mode == kModeOut || mode == kModeKVT;
        float s = 0.f;
        if (i < 3) {
            for (int j = 0; j < 3; ++j) s = fmaf(tr ? W[j * 3 + i] : W[i * 3 + j], ld(base + j), s);
        } else {
            const int ii = i - 3;
            for (int j = 0; j < 5; ++j) s = fmaf(tr ? W[9 + j * 5 + ii] : W[9 + ii * 5 + j], ld(base + 3 + j), s);
        }
        return s;
    }
    int e3 = e2 - g.so3;
    if (e3 < g.so2) {                          // pairs rotated by the token's angles (gta.py:203-219,269-271)
        const int pr = e3 >> 1, i = e3 & 1, base = e - i;
        const float* cs = a.so2cs + (tok * a.C + pr) * 2;
        const float c = cs[0], s = (mode == kModeOut || mode == kModeKVT) ? -cs[1] : cs[1];
        const float x0 = ld(base), x1 = ld(base + 1);
        return i == 0 ? fmaf(c, x0, -s * x1) : fmaf(s, x0, c * x1);
    }
    {                                          // t2: 3-vectors, T = [[1,0,0],[0,1,0],[x,y,1]] (gta.py:72-89,221-238,272-274)
        const int e4 = e3 - g.so2, i = e4 % 3, base = e - i;
        const float px = a.xy[tok * 2], py = a.xy[tok * 2 + 1];
        const float x0 = ld(base), x1 = ld(base + 1), x2 = ld(base + 2);
        if (mode == kModeQ) return i == 0 ? fmaf(-px, x2, x0) : (i == 1 ? fmaf(-py, x2, x1) : x2);   // (T^-1)^T
        if (mode == kModeKV) return i == 2 ? fmaf(px, x0, fmaf(py, x1, x2)) : (i == 0 ? x0 : x1);    // T
        if (mode == kModeKVT) return i == 0 ? fmaf(px, x2, x0) : (i == 1 ? fmaf(py, x2, x1) : x2);   // T^T (backward: dk = rho_k^T dk')
        return i == 2 ? fmaf(-px, x0, fmaf(-py, x1, x2)) : (i == 0 ? x0 : x1);                      // T^-1
    }
}

// modes Q / KV: strided [B,H,T,D] input (TIn) -> dense [B,H,T,Da] (TStore); columns >= D: 0 (or 1, see `ones`).
template <typename TIn, typename TStore>
__global__ void gen_rotate_in_kernel(const GenArgs a) {
    const int64_t total = static_cast<int64_t>(a.B) * a.H * a.T * a.Da;
    int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int e = static_cast<int>(i % a.Da);
    int64_t r = i / a.Da;
    const int t = static_cast<int>(r % a.T); r /= a.T;
    const int h = static_cast<int>(r % a.H);
    const int b = static_cast<int>(r / a.H);
    float y;
    if (e >= a.D) {
        y = (a.ones && e < a.D + 2) ? 1.0f : 0.0f;
    } else {
        const TIn* row = reinterpret_cast<const TIn*>(a.x) + b * a.sb + h * a.sh + t * a.st;
        auto ld = [&](int idx) { return to_f32<TIn>(row[idx]); };
        if (!a.rotate) y = ld(e);
        else y = gen_element(a, e, static_cast<size_t>(b) * a.N + t / a.tpv, static_cast<size_t>(b) * a.T + t,
                             a.tc_ptr ? __ldg(a.tc_ptr) : 1.0f, ld);
    }
    const int64_t o = a.out_bthd ? ((static_cast<int64_t>(b) * a.T + t) * a.H + h) * a.Da + e : i;
    reinterpret_cast<TStore*>(a.out)[o] = from_f32<TStore>(y);
}

// euclid: k'[D], k'[D+1] = -|k'|^2/2 as value + residual in the storage type (one thread per key row, on the
// STORED — already rounded — operand values so that the folded similarity is self-consistent).
template <typename TStore>
__global__ void gen_key_bias_kernel(TStore* __restrict__ kt, int64_t rows, int D, int Da) {
    int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    TStore* row = kt + i * Da;
    float s = 0.f;
    for (int e = 0; e < D; ++e) { const float v = to_f32<TStore>(row[e]); s = fmaf(v, v, s); }
    s *= -0.5f;
    const TStore hi = from_f32<TStore>(s);
    row[D] = hi;
    row[D + 1] = from_f32<TStore>(s - to_f32<TStore>(hi));
}

// mode Out (or, backward, KVT): [B,T,H,Da] (TIn; fp32 for the forward's O') -> [B,T,H,D] (TOut) with rho_q^{-1} (rho_k^T) applied.
template <typename TIn, typename TOut>
__global__ void gen_rotate_out_kernel(const GenArgs a) {
    const int64_t total = static_cast<int64_t>(a.B) * a.T * a.H * a.D;
    int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int e = static_cast<int>(i % a.D);
    int64_t r = i / a.D;
    const int h = static_cast<int>(r % a.H); r /= a.H;
    const int t = static_cast<int>(r % a.T);
    const int b = static_cast<int>(r / a.T);
    const TIn* row = reinterpret_cast<const TIn*>(a.x) + ((static_cast<int64_t>(b) * a.T + t) * a.H + h) * a.Da;
    // euclid_sim: the bias -|k'|^2/2 sits in operand column D, so its gradient is g[D] and dk' = g[:D] - g[D] k'
    const TIn* krow = a.bias_k ? reinterpret_cast<const TIn*>(a.bias_k) + ((static_cast<int64_t>(b) * a.H + h) * a.T + t) * a.Da : nullptr;
    const float gb = krow ? to_f32<TIn>(row[a.D]) : 0.f;
    auto ld = [&](int idx) { return krow ? fmaf(-gb, to_f32<TIn>(krow[idx]), to_f32<TIn>(row[idx])) : to_f32<TIn>(row[idx]); };
    float y;
    if (!a.rotate) y = ld(e);
    else y = gen_element(a, e, static_cast<size_t>(b) * a.N + t / a.tpv, static_cast<size_t>(b) * a.T + t,
                         a.tc_ptr ? __ldg(a.tc_ptr) : 1.0f, ld);
    reinterpret_cast<TOut*>(a.out)[i] = from_f32<TOut>(y);
}

// ------------------------------------------------------------------------------------------------ host side
bool attn_needs_generic(const GtaAttnParams& p) {
    return p.t2 > 0 || p.euclid || ((p.triv | p.se3 | p.so3 | p.so2) & 7);
}
static int padded_dim(const GtaAttnParams& p) { return p.euclid ? p.D + 32 : p.D; }
static size_t align_up(size_t v) { return (v + 1023) & ~static_cast<size_t>(1023); }
static size_t elt(int dtype) { return dtype == GTA_DTYPE_BF16 ? 2 : 4; }

// The core kernel's flags on the generic path: staging flags make no sense here (the operands are rebuilt every call);
// the split-precision pipeline exists for head dims <= 96 only, a padded head dim of 128 multiplies in bf16.
static int core_flags(const GtaAttnParams& p) {
    int f = p.flags & ~(GTA_FLAG_SKIP_STAGE | GTA_FLAG_STAGE_ONLY);
    if (p.in_dtype == GTA_DTYPE_F32 && padded_dim(p) > 96) f |= GTA_FLAG_FAST_FP32;
    return f;
}

struct GenLayout {
    size_t q, k, v, o, core, total;
};
static GenLayout gen_layout(const GtaAttnParams& p) {
    const int Da = padded_dim(p);
    const size_t es = elt(p.in_dtype);
    GenLayout l;
    l.q = 0;
    l.k = l.q + align_up(static_cast<size_t>(p.B) * p.H * p.Tq * Da * es);
    l.v = l.k + align_up(static_cast<size_t>(p.B) * p.H * p.Tk * Da * es);
    l.o = l.v + align_up(static_cast<size_t>(p.B) * p.H * p.Tk * Da * es);
    l.core = l.o + align_up(static_cast<size_t>(p.B) * p.H * p.Tq * Da * 4);
    l.total = l.core + gta_attn_fwd_workspace_bytes_ex(p.B, p.H, p.Tk, Da, p.in_dtype, core_flags(p));
    return l;
}
size_t generic_workspace_bytes(const GtaAttnParams& p) { return gen_layout(p).total; }

static GenArgs make_gen_args(const GtaAttnParams& p, int which /*0 q, 1 k, 2 v, 3 out*/) {
    GenArgs a;
    const bool qside = which == 0 || which == 3;
    a.x = which == 0 ? p.q : (which == 1 ? p.k : p.v);
    a.sb = which == 0 ? p.q_stride_b : (which == 1 ? p.k_stride_b : p.v_stride_b);
    a.sh = which == 0 ? p.q_stride_h : (which == 1 ? p.k_stride_h : p.v_stride_h);
    a.st = which == 0 ? p.q_stride_t : (which == 1 ? p.k_stride_t : p.v_stride_t);
    a.out = nullptr;
    a.B = p.B; a.H = p.H; a.D = p.D; a.Da = padded_dim(p);
    a.T = qside ? p.Tq : p.Tk; a.N = qside ? p.Nq : p.Nk; a.tpv = a.T / a.N;
    a.C = p.so2 >> 1;
    a.gd = GenDims{p.triv, p.se3, p.so3, p.so2, p.t2, p.euclid};
    a.mode = which == 0 ? kModeQ : (which == 3 ? kModeOut : kModeKV);
    a.se3m = which == 0 ? (p.euclid ? p.reps.se3_qi : p.reps.se3_q) : (which == 3 ? p.reps.se3_q : p.reps.se3_k);
    a.so3m = qside ? p.reps.so3_q : p.reps.so3_k;
    a.so2cs = qside ? p.reps.so2_q : p.reps.so2_k;
    a.xy = qside ? p.reps.t2_q : p.reps.t2_k;
    a.tc_ptr = p.trans_coeff;
    a.rotate = (which < 2) || p.v_transform;
    a.ones = (which == 0 && p.euclid) ? 1 : 0;
    a.out_bthd = 0;
    a.se3_lin_t = 0; a.sub_trans = 0; a.bias_k = nullptr;
    return a;
}

template <typename TIn, typename TStore>
static void launch_in(const GenArgs& a, cudaStream_t st) {
    const int64_t total = static_cast<int64_t>(a.B) * a.H * a.T * a.Da;
    gen_rotate_in_kernel<TIn, TStore><<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(a);
}

int launch_attn_fwd_generic(const GtaAttnParams& p, cudaStream_t st) {
    const int Da = padded_dim(p);
    if (Da > 128) return set_error(GTA_ERR_UNSUPPORTED, "euclid_sim needs head dim + 32 <= 128");
    const GenLayout l = gen_layout(p);
    uint8_t* ws = static_cast<uint8_t*>(p.workspace);
    const bool bf = p.in_dtype == GTA_DTYPE_BF16;
    for (int which = 0; which < 3; ++which) {
        GenArgs a = make_gen_args(p, which);
        a.out = ws + (which == 0 ? l.q : (which == 1 ? l.k : l.v));
        if (bf) launch_in<__nv_bfloat16, __nv_bfloat16>(a, st); else launch_in<float, float>(a, st);
    }
    if (p.euclid) {
        const int64_t rows = static_cast<int64_t>(p.B) * p.H * p.Tk;
        const unsigned blocks = static_cast<unsigned>((rows + 127) / 128);
        if (bf) gen_key_bias_kernel<__nv_bfloat16><<<blocks, 128, 0, st>>>(reinterpret_cast<__nv_bfloat16*>(ws + l.k), rows, p.D, Da);
        else gen_key_bias_kernel<float><<<blocks, 128, 0, st>>>(reinterpret_cast<float*>(ws + l.k), rows, p.D, Da);
    }
    int rc = check_launch("gta_attn_fwd (generic rep application)");
    if (rc) return rc;

    GtaAttnParams c = p;                       // the tensor-core attention on the prepared operands
    c.q = ws + l.q; c.k = ws + l.k; c.v = ws + l.v;
    c.q_stride_t = c.k_stride_t = c.v_stride_t = Da;
    c.q_stride_h = static_cast<int64_t>(p.Tq) * Da; c.k_stride_h = c.v_stride_h = static_cast<int64_t>(p.Tk) * Da;
    c.q_stride_b = c.q_stride_h * p.H; c.k_stride_b = c.v_stride_b = c.k_stride_h * p.H;
    c.out = ws + l.o; c.out_dtype = GTA_DTYPE_F32;
    c.D = Da; c.triv = Da; c.se3 = c.so3 = c.so2 = c.t2 = 0; c.euclid = 0;
    c.reps = GtaReps{};
    c.trans_coeff = nullptr;
    c.v_transform = 0;
    c.flags = core_flags(p);
    c.workspace = ws + l.core;
    c.workspace_bytes = p.workspace_bytes - l.core;
    rc = gta_attn_fwd(&c, st);
    if (rc) return rc;

    GenArgs o = make_gen_args(p, 3);
    o.x = ws + l.o;
    o.out = p.out;
    const int64_t total = static_cast<int64_t>(p.B) * p.Tq * p.H * p.D;
    const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
    if (p.out_dtype == GTA_DTYPE_BF16) gen_rotate_out_kernel<float, __nv_bfloat16><<<blocks, 256, 0, st>>>(o);
    else gen_rotate_out_kernel<float, float><<<blocks, 256, 0, st>>>(o);
    return check_launch("gta_attn_fwd (generic output rep)");
}

// ------------------------------------------------------------------------------------------------ generic backward
// The reps are constant linear maps, so the backward of the generic path is the forward's scheme run around the
// tensor-core backward (gta_attn_bwd.cu / gta_attn_bwd2.cu) on dense operands with an all-trivial head layout:
//   q', k', v' (as in the forward), dO' = rho_q^{-T} dO  ->  dQ', dK', dV'  ->  dq = rho_q^{-1} dQ', dk = rho_k^T dK', dv = rho_k^T dV'.
// delta = rowsum(dO * O) is invariant under the output rep and is taken from the un-rotated pair.  d(trans_coeff): the four
// per-4-vector terms of the SE(3) block (query, key, value, output side), one element-wise reduction kernel.
// Not for euclid_sim (its forward has no log-sum-exp output).
struct GenDtcArgs {
    const void* raw; int64_t sb, sh, st;     // raw input rows (q, k or v), strided [B,H,T,D]; part 3: O [B,T,H,D]
    const void* g;                           // un-rotated gradient [B,T,H,D] contiguous (dQ', dK', dV'); part 3: dO
    const float* se3m;                       // [B,N,16]
    float* dtc;
    int B, H, T, D, Dg, N, tpv, triv, se3, part; // part 0: query side, 1/2: key / value side, 3: output side; Dg: row pitch of g
    int euclid;                              // homogenised 3-vectors: every term is the gradient dotted with the translation column
    const void* bias_k; int Da;              // euclid, key side: dense k' [B,H,T,Da]; the gradient is g[e] - g[D] k'[e] (bias -|k'|^2/2)
};

template <typename T>
__global__ void gen_dtc_kernel(const GenDtcArgs a) {
    const int vlen = a.euclid ? 3 : 4;
    const int nv = a.se3 / vlen;
    const int64_t total = static_cast<int64_t>(a.B) * a.T * a.H * nv;
    float part = 0.f;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int v4 = static_cast<int>(i % nv);
        int64_t r = i / nv;
        const int h = static_cast<int>(r % a.H); r /= a.H;
        const int t = static_cast<int>(r % a.T);
        const int b = static_cast<int>(r / a.T);
        const float* M = a.se3m + (static_cast<size_t>(b) * a.N + t / a.tpv) * 16;
        const T* g = reinterpret_cast<const T*>(a.g) + ((static_cast<int64_t>(b) * a.T + t) * a.H + h) * a.Dg + a.triv + vlen * v4;
        float g0 = to_f32<T>(g[0]), g1 = to_f32<T>(g[1]), g2 = to_f32<T>(g[2]);
        if (a.euclid) {                 // y = A x + t * tc  =>  d/dtc = g . t
            if (a.bias_k) {
                const T* kr = reinterpret_cast<const T*>(a.bias_k) + ((static_cast<int64_t>(b) * a.H + h) * a.T + t) * a.Da + a.triv + vlen * v4;
                const float gb = to_f32<T>(g[a.D - a.triv - vlen * v4]);        // column D of the row
                g0 = fmaf(-gb, to_f32<T>(kr[0]), g0); g1 = fmaf(-gb, to_f32<T>(kr[1]), g1); g2 = fmaf(-gb, to_f32<T>(kr[2]), g2);
            }
            part += g0 * M[3] + g1 * M[7] + g2 * M[11];
            continue;
        }
        const float g3 = to_f32<T>(g[3]);
        const T* y = a.part == 3 ? reinterpret_cast<const T*>(a.raw) + ((static_cast<int64_t>(b) * a.T + t) * a.H + h) * a.D + a.triv + 4 * v4
                                 : reinterpret_cast<const T*>(a.raw) + b * a.sb + h * a.sh + t * a.st + a.triv + 4 * v4;
        if (a.part == 0) part += g3 * (M[3] * to_f32<T>(y[0]) + M[7] * to_f32<T>(y[1]) + M[11] * to_f32<T>(y[2]));
        else if (a.part == 3) part += (g0 * M[3] + g1 * M[7] + g2 * M[11]) * to_f32<T>(y[3]) / M[15];
        else part += (g0 * M[3] + g1 * M[7] + g2 * M[11]) * to_f32<T>(y[3]);
    }
    __shared__ float red[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += red[w];
        if (s != 0.f) atomicAdd(a.dtc, s);
    }
}

struct GenBwdLayout {
    size_t q, k, v, dout, dq, dk, dv, oeff, dpad, core, total;
};
static GenBwdLayout gen_bwd_layout(const GtaAttnParams& p) {
    const int Da = padded_dim(p);
    const size_t es = elt(p.in_dtype);
    const size_t nq = align_up(static_cast<size_t>(p.B) * p.H * p.Tq * Da * es), nk = align_up(static_cast<size_t>(p.B) * p.H * p.Tk * Da * es);
    GenBwdLayout l;
    l.q = 0; l.k = l.q + nq; l.v = l.k + nk; l.dout = l.v + nk;
    l.dq = l.dout + nq; l.dk = l.dq + nq; l.dv = l.dk + nk;
    l.oeff = l.dv + nk;                                   // euclid only: O - tc * translation, padded; dO padded
    l.dpad = l.oeff + (p.euclid ? nq : 0);
    l.core = l.dpad + (p.euclid ? nq : 0);
    l.total = l.core + attn_bwd_workspace_bytes(p.B, p.H, p.Tq, p.Tk, Da);
    return l;
}
size_t generic_bwd_workspace_bytes(const GtaAttnParams& p) { return gen_bwd_layout(p).total; }

int launch_attn_bwd_generic(const GtaAttnBwdParams& bp, cudaStream_t st) {
    const GtaAttnParams& p = bp.fwd;
    const int Da = padded_dim(p);
    if (Da > 128) return set_error(GTA_ERR_UNSUPPORTED, "euclid_sim needs head dim + 32 <= 128");
    if (p.out_dtype != p.in_dtype) return set_error(GTA_ERR_UNSUPPORTED, "gta_attn_bwd: out/dout must have the dtype of q/k/v");
    if (!p.lse || !bp.dout || !bp.dq || !bp.dk || !bp.dv) return set_error(GTA_ERR_INVALID, "gta_attn_bwd: null lse/dout/dq/dk/dv");
    const GenBwdLayout l = gen_bwd_layout(p);
    if (!bp.workspace || bp.workspace_bytes < l.total) return set_error(GTA_ERR_INVALID, "gta_attn_bwd: workspace too small (need %zu bytes)", l.total);
    uint8_t* ws = static_cast<uint8_t*>(bp.workspace);
    const bool bf = p.in_dtype == GTA_DTYPE_BF16;
    const int64_t do_sb = static_cast<int64_t>(p.Tq) * p.H * p.D, do_sh = p.D, do_st = static_cast<int64_t>(p.H) * p.D;   // [B,Tq,H,D] as [B,H,T,D]
    for (int which = 0; which < 4; ++which) {          // q', k', v' as in the forward, and dO' = (output map)^T dO
        GenArgs a = make_gen_args(p, which == 3 ? 0 : which);
        if (which == 3) {
            a.x = bp.dout; a.sb = do_sb; a.sh = do_sh; a.st = do_st;
            a.rotate = p.v_transform;
            a.out_bthd = 1;
            a.ones = 0;
            if (p.euclid) { a.se3m = p.reps.se3_q; a.se3_lin_t = 1; }      // output: y = A_o x + t_o tc with A_o | t_o from E_q (gta.py:251-253)
        }
        a.out = ws + (which == 0 ? l.q : (which == 1 ? l.k : (which == 2 ? l.v : l.dout)));
        if (bf) launch_in<__nv_bfloat16, __nv_bfloat16>(a, st); else launch_in<float, float>(a, st);
    }
    if (p.euclid) {
        const int64_t rows = static_cast<int64_t>(p.B) * p.H * p.Tk;
        const unsigned blocks = static_cast<unsigned>((rows + 127) / 128);
        if (bf) gen_key_bias_kernel<__nv_bfloat16><<<blocks, 128, 0, st>>>(reinterpret_cast<__nv_bfloat16*>(ws + l.k), rows, p.D, Da);
        else gen_key_bias_kernel<float><<<blocks, 128, 0, st>>>(reinterpret_cast<float*>(ws + l.k), rows, p.D, Da);
        // delta = rowsum(dO' * O') with O = A_o O' + t_o tc: rowsum(dO * (O - t_o tc)) -> the pair (O - t_o tc, dO), padded to Da
        for (int which = 0; which < 2; ++which) {
            GenArgs a = make_gen_args(p, 0);
            a.x = which == 0 ? p.out : bp.dout; a.sb = do_sb; a.sh = do_sh; a.st = do_st;
            a.ones = 0; a.out_bthd = 1;
            a.rotate = which == 0 && p.v_transform;
            a.se3m = p.reps.se3_q; a.sub_trans = 1;
            a.out = ws + (which == 0 ? l.oeff : l.dpad);
            if (bf) launch_in<__nv_bfloat16, __nv_bfloat16>(a, st); else launch_in<float, float>(a, st);
        }
    }
    int rc = check_launch("gta_attn_bwd (generic rep application)");
    if (rc) return rc;

    GtaAttnBwdParams c = bp;                           // the tensor-core backward on the prepared operands
    GtaAttnParams& f = c.fwd;
    f.q = ws + l.q; f.k = ws + l.k; f.v = ws + l.v;
    f.q_stride_t = f.k_stride_t = f.v_stride_t = Da;
    f.q_stride_h = static_cast<int64_t>(p.Tq) * Da; f.k_stride_h = f.v_stride_h = static_cast<int64_t>(p.Tk) * Da;
    f.q_stride_b = f.q_stride_h * p.H; f.k_stride_b = f.v_stride_b = f.k_stride_h * p.H;
    f.D = Da; f.triv = Da; f.se3 = f.so3 = f.so2 = f.t2 = 0; f.euclid = 0;
    f.reps = GtaReps{};
    f.trans_coeff = nullptr;
    f.v_transform = 0;
    if (p.euclid) f.out = ws + l.oeff;
    c.dout = ws + l.dout;
    c.dq = ws + l.dq; c.dk = ws + l.dk; c.dv = ws + l.dv;
    c.dtrans_coeff = nullptr;
    c.workspace = ws + l.core;
    c.workspace_bytes = bp.workspace_bytes - l.core;
    rc = launch_attn_bwd(c, st, /*delta_dout=*/p.euclid ? static_cast<const void*>(ws + l.dpad) : bp.dout);   // delta from the un-rotated pair
    if (rc) return rc;

    if (bp.dtrans_coeff && p.se3 > 0) {
        for (int part = 0; part < 4; ++part) {
            if (part >= 2 && !p.v_transform) continue;
            GenDtcArgs d;
            const bool qside = part == 0 || part == 3;
            d.raw = part == 0 ? p.q : (part == 1 ? p.k : (part == 2 ? p.v : p.out));
            d.sb = part == 0 ? p.q_stride_b : (part == 1 ? p.k_stride_b : p.v_stride_b);
            d.sh = part == 0 ? p.q_stride_h : (part == 1 ? p.k_stride_h : p.v_stride_h);
            d.st = part == 0 ? p.q_stride_t : (part == 1 ? p.k_stride_t : p.v_stride_t);
            d.g = part == 0 ? static_cast<const void*>(ws + l.dq) : (part == 1 ? ws + l.dk : (part == 2 ? ws + l.dv : bp.dout));
            d.Dg = part == 3 ? p.D : Da;
            // euclid: the query side uses c_q = inv(E_q) (se3_qi), the output side E_q (gta.py:140,153,251-253)
            d.se3m = qside ? ((p.euclid && part == 0) ? p.reps.se3_qi : p.reps.se3_q) : p.reps.se3_k;
            d.dtc = bp.dtrans_coeff;
            d.B = p.B; d.H = p.H; d.D = p.D; d.T = qside ? p.Tq : p.Tk; d.N = qside ? p.Nq : p.Nk; d.tpv = d.T / d.N;
            d.triv = p.triv; d.se3 = p.se3; d.part = part; d.euclid = p.euclid;
            d.bias_k = (p.euclid && part == 1) ? ws + l.k : nullptr; d.Da = Da;
            const int64_t total = static_cast<int64_t>(d.B) * d.T * d.H * (d.se3 / (p.euclid ? 3 : 4));
            const unsigned blocks = static_cast<unsigned>(std::min<int64_t>((total + 255) / 256, 148 * 8));
            if (bf) gen_dtc_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(d); else gen_dtc_kernel<float><<<blocks, 256, 0, st>>>(d);
        }
    }
    for (int which = 0; which < 3; ++which) {          // dq = (query map)^T dQ', dk = rho_k^T dK', dv = rho_k^T dV'
        GenArgs o = make_gen_args(p, which == 0 ? 3 : which);
        if (which > 0) o.mode = kModeKVT;
        o.ones = 0;
        o.x = ws + (which == 0 ? l.dq : (which == 1 ? l.dk : l.dv));
        o.out = which == 0 ? bp.dq : (which == 1 ? bp.dk : bp.dv);
        o.rotate = (which < 2) || p.v_transform;
        if (p.euclid) {
            o.se3_lin_t = 1;
            if (which == 0) o.se3m = p.reps.se3_qi;     // q' = A_q q + t_q tc with A_q | t_q from inv(E_q)
            if (which == 1) o.bias_k = ws + l.k;        // the -|k'|^2/2 columns: dk' = g[:D] - g[D] k'
        }
        const int64_t total = static_cast<int64_t>(o.B) * o.T * o.H * o.D;
        const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
        if (bf) gen_rotate_out_kernel<__nv_bfloat16, __nv_bfloat16><<<blocks, 256, 0, st>>>(o);
        else gen_rotate_out_kernel<float, float><<<blocks, 256, 0, st>>>(o);
    }
    return check_launch("gta_attn_bwd (generic gradient reps)");
}

// attn[b,h,i,j] = exp(q'_i . k'_j * scale - lse[b,h,i]): 16x16 output tile per block, operands staged through shared memory.
__global__ void attn_probs_kernel(const float* __restrict__ qt, const float* __restrict__ kt, const float* __restrict__ lse,
                                  float* __restrict__ attn, int Tq, int Tk, int D, float scale) {
    __shared__ float sq[16][129], sk[16][129];
    const int bh = blockIdx.z, i0 = blockIdx.y * 16, j0 = blockIdx.x * 16;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const float* qb = qt + static_cast<size_t>(bh) * Tq * D;
    const float* kb = kt + static_cast<size_t>(bh) * Tk * D;
    for (int idx = threadIdx.x; idx < 16 * D; idx += 256) {
        const int r = idx / D, c = idx - r * D;
        sq[r][c] = (i0 + r < Tq) ? qb[static_cast<size_t>(i0 + r) * D + c] : 0.f;
        sk[r][c] = (j0 + r < Tk) ? kb[static_cast<size_t>(j0 + r) * D + c] : 0.f;
    }
    __syncthreads();
    const int i = i0 + ty, j = j0 + tx;
    if (i >= Tq || j >= Tk) return;
    float dot = 0.f;
    for (int c = 0; c < D; ++c) dot = fmaf(sq[ty][c], sk[tx][c], dot);
    attn[(static_cast<size_t>(bh) * Tq + i) * Tk + j] = expf(dot * scale - lse[static_cast<size_t>(bh) * Tq + i]);
}

size_t attn_probs_workspace_bytes(int B, int H, int Tq, int Tk, int D) {
    if (B <= 0 || H <= 0 || Tq <= 0 || Tk <= 0 || D <= 0) return 0;
    return align_up(static_cast<size_t>(B) * H * Tq * D * 4) + align_up(static_cast<size_t>(B) * H * Tk * D * 4);
}

int launch_attn_probs(const GtaAttnParams& p, float* attn, cudaStream_t st) {
    if (p.euclid) return set_error(GTA_ERR_UNSUPPORTED, "gta_attn_probs: not defined for euclid_sim");
    if (!p.lse || !attn) return set_error(GTA_ERR_INVALID, "gta_attn_probs: null lse / attn");
    if (!p.workspace || p.workspace_bytes < attn_probs_workspace_bytes(p.B, p.H, p.Tq, p.Tk, p.D))
        return set_error(GTA_ERR_INVALID, "gta_attn_probs: workspace too small");
    float* qt = static_cast<float*>(p.workspace);
    float* kt = reinterpret_cast<float*>(static_cast<uint8_t*>(p.workspace) + align_up(static_cast<size_t>(p.B) * p.H * p.Tq * p.D * 4));
    int rc = launch_rotate_debug_generic(p, qt, kt, nullptr, st);      // dense fp32 q' = rho_q^{-T} q, k' = rho_k k
    if (rc) return rc;
    dim3 grid((p.Tk + 15) / 16, (p.Tq + 15) / 16, p.B * p.H);
    attn_probs_kernel<<<grid, 256, 0, st>>>(qt, kt, p.lse, attn, p.Tq, p.Tk, p.D, p.scale);
    return check_launch("gta_attn_probs");
}

// fp32 rotated operands [B,H,T,D] (no padding) for tests.
int launch_rotate_debug_generic(const GtaAttnParams& p, float* qt, float* kt, float* vt, cudaStream_t st) {
    for (int which = 0; which < 3; ++which) {
        float* out = which == 0 ? qt : (which == 1 ? kt : vt);
        if (!out) continue;
        GenArgs a = make_gen_args(p, which);
        a.Da = p.D; a.ones = 0;
        a.out = out;
        if (p.in_dtype == GTA_DTYPE_BF16) launch_in<__nv_bfloat16, float>(a, st); else launch_in<float, float>(a, st);
    }
    return check_launch("gta_rotate_debug (generic)");
}

__global__ void t2_mats_kernel(const float* __restrict__ coord, int64_t n, float* __restrict__ mats, float* __restrict__ inv) {
    int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = coord[i * 2], y = coord[i * 2 + 1];
    const float m[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, x, y, 1.f};
    for (int j = 0; j < 9; ++j) mats[i * 9 + j] = m[j];
    if (inv) {
        const float v[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, -x, -y, 1.f};
        for (int j = 0; j < 9; ++j) inv[i * 9 + j] = v[j];
    }
}

int launch_t2_mats(const float* coord, int64_t n, float* mats, float* inv, cudaStream_t st) {
    if (n <= 0 || !coord || !mats) return set_error(GTA_ERR_INVALID, "gta_t2_mats: empty input");
    t2_mats_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(coord, n, mats, inv);
    return check_launch("gta_t2_mats");
}

}  // namespace gta
