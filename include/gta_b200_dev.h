/*
 * gta_b200 — development / measurement hooks, built into a SEPARATE library (gta_b200/libgta_b200_dev.so).  Nothing here is
 * part of the product ABI (include/gta_b200.h); these entry points exist for tests/test_gpu_parity.py::test_umma_probe_exact
 * and the micro-benchmarks under tools/.
 */
#ifndef GTA_B200_DEV_H_
#define GTA_B200_DEV_H_

#include "gta_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* tcgen05 self-test: S = A B^T (A,B [128,D] bf16 row-major) and O = P V (P [128,128] bf16, V [128,D] bf16) through the same
 * descriptor helpers and tile images as gta_attn_fwd.  outS [128,128], outO [128,D] fp32.  p_in_tmem selects the TS form for
 * the PV product. */
int gta_dev_umma_probe(const void* A, const void* Bm, const void* P, const void* V, int D, int p_in_tmem,
                       float* outS, float* outO, void* stream);

/* tcgen05.mma issue/throughput micro-benchmark (tools/umma_bench.py): out[grid][2] = clocks (issue, issue+drain). */
int gta_dev_umma_bench(int D, int mode, int reps, int grid, long long* out, void* stream);

/* exp2/pack phase micro-benchmark (tools/softmax_bench.py): clk[grid] = clocks of `reps` 128-column rows per thread. */
int gta_dev_softmax_bench(int num, int den, int warps, int reps, int grid, const float* in, float* out, long long* clk,
                          void* stream);

/* First-generation attention kernel (one query tile per CTA, one softmax warpgroup; round 1's measured starting point,
 * 22.7 % of peak) on a workspace that gta_attn_fwd(GTA_FLAG_STAGE_ONLY) of the product library has staged.
 * p->flags & 1: P operand of the PV MMA read from tensor memory instead of shared memory. */
int gta_dev_attn_fwd_v0(const GtaAttnParams* p, void* stream);

const char* gta_dev_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* GTA_B200_DEV_H_ */
