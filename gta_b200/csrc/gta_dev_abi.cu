// extern "C" entry points of libgta_b200_dev.so (include/gta_b200_dev.h): probes, micro-benchmarks and the first-generation
// attention kernel.  A separate library: none of this ships in the product ABI.
#include "common.cuh"
#include "../../include/gta_b200_dev.h"

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

}  // namespace gta

using namespace gta;

extern "C" {

const char* gta_dev_last_error(void) { return g_err; }

int gta_dev_umma_probe(const void* A, const void* Bm, const void* P, const void* V, int D, int p_in_tmem, float* outS,
                       float* outO, void* stream) {
    return launch_umma_probe(A, Bm, P, V, D, p_in_tmem, outS, outO, static_cast<cudaStream_t>(stream));
}

int gta_dev_umma_bench(int D, int mode, int reps, int grid, long long* out, void* stream) {
    return launch_umma_bench(D, mode, reps, grid, out, static_cast<cudaStream_t>(stream));
}

int gta_dev_softmax_bench(int num, int den, int warps, int reps, int grid, const float* in, float* out, long long* clk,
                          void* stream) {
    return launch_softmax_bench(num, den, warps, reps, grid, in, out, clk, static_cast<cudaStream_t>(stream));
}

int gta_dev_attn_fwd_v0(const GtaAttnParams* p, void* stream) {
    if (!p || !p->q || !p->out || !p->workspace) return set_error(GTA_ERR_INVALID, "null params");
    return launch_attn_fwd_v0(*p, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
