// Rep construction on the device — one launch for the per-view matrices and one for the per-token
// SO(2) tables.  Replaces the ~60 tiny ATen launches of pre_compute_reps
// (reference: source/encoder.py:183-265, source/decoder.py:247-353).
#include "common.cuh"

namespace gta {

// General 4x4 inverse (Gauss-Jordan, partial pivoting) in fp64 — torch.linalg.inv at encoder.py:219.
__device__ bool inv4(const double* a, double* out) {
    double m[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { m[i][j] = a[i * 4 + j]; m[i][4 + j] = (i == j) ? 1.0 : 0.0; }
    for (int c = 0; c < 4; ++c) {
        int p = c;
        for (int r = c + 1; r < 4; ++r) if (fabs(m[r][c]) > fabs(m[p][c])) p = r;
        if (m[p][c] == 0.0) return false;
        if (p != c) for (int j = 0; j < 8; ++j) { double t = m[c][j]; m[c][j] = m[p][j]; m[p][j] = t; }
        double d = 1.0 / m[c][c];
        for (int j = 0; j < 8; ++j) m[c][j] *= d;
        for (int r = 0; r < 4; ++r) if (r != c) {
            double f = m[r][c];
            for (int j = 0; j < 8; ++j) m[r][j] -= f * m[c][j];
        }
    }
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) out[i * 4 + j] = m[i][4 + j];
    return true;
}

// ZYZ Euler angles incl. the two gimbal branches (wigner_d.py:39-49).
__device__ void zyz_euler(const double* R, double* g) {
    const double eps = 1e-5;
    double g1 = atan2(R[7], -R[6]);
    double g2 = atan2(sqrt(R[2] * R[2] + R[5] * R[5]), R[8]);
    double g3 = atan2(R[5], R[2]);
    if (fabs(R[8] - 1.0) < eps) { g1 = atan2(R[3], R[0]); g3 = 0.0; }
    else if (fabs(R[8] + 1.0) < eps) { g1 = atan2(-R[3], -R[0]); g3 = 0.0; }
    g[0] = g1; g[1] = g2; g[2] = g3;
}

// out = Z(a) * in for the (2l+1)-dim z-rotation: row i mixes rows i and n-1-i (wigner_d.py:16-25).
template <int N>
__device__ void zrot_left(double a, const double* in, double* out) {
    constexpr int l = (N - 1) / 2;
    for (int i = 0; i < N; ++i) {
        double f = static_cast<double>(l - i), c = cos(f * a), s = sin(f * a);
        for (int j = 0; j < N; ++j)
            out[i * N + j] = (i == N - 1 - i) ? in[i * N + j] : c * in[i * N + j] + s * in[(N - 1 - i) * N + j];
    }
}
template <int N>
__device__ void mat_left(const double* J, const double* in, double* out) {
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            double s = 0.0;
            for (int k = 0; k < N; ++k) s += J[i * N + k] * in[k * N + j];
            out[i * N + j] = s;
        }
}
// D_l = Z(g3) J Z(g2) J Z(g1)  (wigner_d.py:28-35), J = J_dense.pt[l].
template <int N>
__device__ void wigner(const double* g, const double* J, float* out) {
    double a[N * N], b[N * N];
    for (int i = 0; i < N * N; ++i) a[i] = (i / N == i % N) ? 1.0 : 0.0;
    zrot_left<N>(g[0], a, b);
    mat_left<N>(J, b, a);
    zrot_left<N>(g[1], a, b);
    mat_left<N>(J, b, a);
    zrot_left<N>(g[2], a, b);
    for (int i = 0; i < N * N; ++i) out[i] = static_cast<float>(b[i]);
}

__device__ void wigner_d12(const double* R, float* d1, float* d2) {
    const double J1[9] = {0, -1, 0, -1, 0, 0, 0, 0, 1};
    const double s3 = 0.86602540378443864676;
    const double J2[25] = {0, 0, 0, -1, 0, 0, 1, 0, 0, 0, 0, 0, -0.5, 0, -s3, -1, 0, 0, 0, 0, 0, 0, -s3, 0, 0.5};
    double g[3];
    zyz_euler(R, g);
    wigner<3>(g, J1, d1);
    wigner<5>(g, J2, d2);
}

// One thread per view.  side 0: query views (se3 = E), side 1: key views (se3 = inv(E)).
__device__ __forceinline__ void build_view_reps(int i, const float* __restrict__ extr_q, const float* __restrict__ extr_k, int nq,
                                                int nk, int want_so3, float* __restrict__ se3_q, float* __restrict__ se3_k,
                                                float* __restrict__ so3_q, float* __restrict__ so3_k) {
    if (i >= nq + nk) return;
    const bool key = i >= nq;
    const int v = key ? i - nq : i;
    const float* E = (key ? extr_k : extr_q) + static_cast<size_t>(v) * 16;
    double e[16], ie[16];
    for (int j = 0; j < 16; ++j) e[j] = static_cast<double>(E[j]);
    if (!inv4(e, ie)) for (int j = 0; j < 16; ++j) ie[j] = nan("");
    float* se3 = key ? se3_k : se3_q;
    if (se3) for (int j = 0; j < 16; ++j) se3[static_cast<size_t>(v) * 16 + j] = static_cast<float>(key ? ie[j] : e[j]);
    float* so3 = key ? so3_k : so3_q;
    if (want_so3 && so3) {
        double R[9];
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R[r * 3 + c] = ie[r * 4 + c];
        wigner_d12(R, so3 + static_cast<size_t>(v) * 34, so3 + static_cast<size_t>(v) * 34 + 9);
    }
}

// theta = fp32(max_freq * 2*pi) * fp32(coord * freq_j), freq_j = 2^(j+1)/2^n (gta.py:57-63); pair index
// j*2 + axis (gta.py:68 + encoder.py:195).  One thread per (token, pair).
__device__ __forceinline__ void so2_table(int64_t i, const float* __restrict__ coord, int64_t ntok, int nfreqs, float wh, float ww,
                                          int shared, float* __restrict__ cs /* [ntok, 2*nfreqs, 2] */,
                                          float* __restrict__ mats /* [ntok, 2*nfreqs, 2, 2] or null */) {
    const int C = 2 * nfreqs;
    if (i >= ntok * C) return;
    const int64_t t = i / C;
    const int p = static_cast<int>(i % C), j = p >> 1, axis = p & 1;
    const float freq = shared ? 1.0f : exp2f(static_cast<float>(j + 1 - nfreqs));
    const float th = (axis ? ww : wh) * (coord[t * 2 + axis] * freq);
    float s, c;
    sincosf(th, &s, &c);
    if (cs) { cs[i * 2] = c; cs[i * 2 + 1] = s; }
    if (mats) { mats[i * 4] = c; mats[i * 4 + 1] = -s; mats[i * 4 + 2] = s; mats[i * 4 + 3] = c; }
}

__global__ void so2_table_kernel(const float* __restrict__ coord, int64_t ntok, int nfreqs, float wh, float ww, int shared,
                                 float* __restrict__ cs, float* __restrict__ mats) {
    so2_table(static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x, coord, ntok, nfreqs, wh, ww, shared, cs, mats);
}

// pre_compute_reps in ONE launch: blocks [0, nb_view) build the per-view tables (one thread per view), the next nb_q blocks
// the query-side SO(2) table, the rest the key-side one (three back-to-back launches of ~7-12 us each were 3.6 % of the
// headline step; the parts are independent).
struct BuildRepsArgs {
    const float* extr_q; const float* extr_k; const float* coord_q; const float* coord_k;
    int nq, nk, want_so3;
    float* se3_q; float* se3_k; float* so3_q; float* so3_k; float* so2_q; float* so2_k;
    int64_t ntok_q, ntok_k;
    int nfreqs, shared;
    float wh, ww;
    unsigned nb_view, nb_q;
};
__global__ void __launch_bounds__(256) build_reps_kernel(const BuildRepsArgs a) {
    if (blockIdx.x < a.nb_view) {
        // ONE warp of views per block (the other warps leave): the fp64 inverse / Euler / Wigner chain is bound by the SM's few
        // fp64 lanes, so the views are spread over as many SMs as possible (with 256 views per block the merged launch was
        // ~6 us slower than the three separate launches it replaced at the decoder shapes)
        if (threadIdx.x < 32)
            build_view_reps(blockIdx.x * 32 + threadIdx.x, a.extr_q, a.extr_k, a.nq, a.nk, a.want_so3, a.se3_q, a.se3_k, a.so3_q, a.so3_k);
    } else if (blockIdx.x < a.nb_view + a.nb_q) {
        so2_table(static_cast<int64_t>(blockIdx.x - a.nb_view) * 256 + threadIdx.x, a.coord_q, a.ntok_q, a.nfreqs, a.wh, a.ww, a.shared, a.so2_q, nullptr);
    } else {
        so2_table(static_cast<int64_t>(blockIdx.x - a.nb_view - a.nb_q) * 256 + threadIdx.x, a.coord_k, a.ntok_k, a.nfreqs, a.wh, a.ww, a.shared, a.so2_k, nullptr);
    }
}

// torch.linalg.inv of a batch of 4x4 extrinsics (source/encoder.py:219).
__global__ void se3_inverse_kernel(const float* __restrict__ E, int64_t n, float* __restrict__ out) {
    int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double e[16], ie[16];
    for (int j = 0; j < 16; ++j) e[j] = static_cast<double>(E[i * 16 + j]);
    if (!inv4(e, ie)) for (int j = 0; j < 16; ++j) ie[j] = nan("");
    for (int j = 0; j < 16; ++j) out[i * 16 + j] = static_cast<float>(ie[j]);
}

__global__ void wigner_kernel(const float* __restrict__ R, int64_t n, float* __restrict__ d1, float* __restrict__ d2) {
    int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double r[9];
    for (int j = 0; j < 9; ++j) r[j] = static_cast<double>(R[i * 9 + j]);
    float a[9], b[25];
    wigner_d12(r, a, b);
    for (int j = 0; j < 9; ++j) d1[i * 9 + j] = a[j];
    for (int j = 0; j < 25; ++j) d2[i * 25 + j] = b[j];
}

static inline float two_pi_times(float f) { return static_cast<float>(static_cast<double>(f) * 2.0 * 3.14159265358979323846); }

int launch_build_reps(const float* extr_q, const float* extr_k, const float* coord_q, const float* coord_k, int B,
                      int Nq, int Nk, int Tq, int Tk, int so2_nfreqs, float mfh, float mfw, int shared,
                      int so3_maxdeg, float* se3_q, float* se3_k, float* so3_q, float* so3_k, float* so2_q,
                      float* so2_k, cudaStream_t st) {
    if (B <= 0 || Nq <= 0 || Nk <= 0) return set_error(GTA_ERR_INVALID, "gta_build_reps: empty batch/views");
    if (so3_maxdeg != 0 && so3_maxdeg != 2)
        return set_error(GTA_ERR_UNSUPPORTED, "gta_build_reps: only so3 max degree 2 is implemented");
    BuildRepsArgs a{};
    if (extr_q && extr_k && (se3_q || se3_k || so3_q || so3_k)) {
        a.extr_q = extr_q; a.extr_k = extr_k; a.nq = B * Nq; a.nk = B * Nk; a.want_so3 = so3_maxdeg == 2;
        a.se3_q = se3_q; a.se3_k = se3_k; a.so3_q = so3_q; a.so3_k = so3_k;
        a.nb_view = static_cast<unsigned>((a.nq + a.nk + 31) / 32);
    }
    unsigned nb_k = 0;
    if (so2_nfreqs > 0) {
        a.wh = two_pi_times(mfh); a.ww = two_pi_times(mfw); a.nfreqs = so2_nfreqs; a.shared = shared;
        const int C = 2 * so2_nfreqs;
        if (so2_q && coord_q) {
            a.coord_q = coord_q; a.so2_q = so2_q; a.ntok_q = static_cast<int64_t>(B) * Tq;
            a.nb_q = static_cast<unsigned>((a.ntok_q * C + 255) / 256);
        }
        if (so2_k && coord_k && so2_k != so2_q) {
            a.coord_k = coord_k; a.so2_k = so2_k; a.ntok_k = static_cast<int64_t>(B) * Tk;
            nb_k = static_cast<unsigned>((a.ntok_k * C + 255) / 256);
        }
    }
    const unsigned nb = a.nb_view + a.nb_q + nb_k;
    if (nb) build_reps_kernel<<<nb, 256, 0, st>>>(a);
    return check_launch("gta_build_reps");
}

int launch_so2_mats(const float* coord, int64_t n, int nfreqs, float mfh, float mfw, int shared, float* mats,
                    cudaStream_t st) {
    if (n <= 0 || nfreqs <= 0) return set_error(GTA_ERR_INVALID, "gta_so2_mats: empty input");
    int64_t tot = n * 2 * nfreqs;
    so2_table_kernel<<<static_cast<unsigned>((tot + 255) / 256), 256, 0, st>>>(coord, n, nfreqs, two_pi_times(mfh),
                                                                               two_pi_times(mfw), shared, nullptr, mats);
    return check_launch("gta_so2_mats");
}

int launch_se3_inverse(const float* extr, int64_t n, float* inv, cudaStream_t st) {
    if (n <= 0 || !extr || !inv) return set_error(GTA_ERR_INVALID, "gta_se3_inverse: empty input");
    se3_inverse_kernel<<<static_cast<unsigned>((n + 63) / 64), 64, 0, st>>>(extr, n, inv);
    return check_launch("gta_se3_inverse");
}

int launch_wigner(const float* R, int64_t n, float* d1, float* d2, cudaStream_t st) {
    if (n <= 0) return set_error(GTA_ERR_INVALID, "gta_wigner_d: empty input");
    wigner_kernel<<<static_cast<unsigned>((n + 63) / 64), 64, 0, st>>>(R, n, d1, d2);
    return check_launch("gta_wigner_d");
}

}  // namespace gta
