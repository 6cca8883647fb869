// Shared host/device helpers for libmamdr_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/mamdr_b200.h"

struct mamdr_ctx {
    int   device;
    int   sm_count;
    int   max_smem_optin;
    int   pass_ctas;   // CTAs of the persistent pass kernel (0 = one per SM); mamdr_ctx_set_pass_ctas
    void* tmap_cache;  // tensor-map cache of the pass kernel (tc_tmap.cuh)
    void* prog;        // program being recorded (mamdr_program_begin .. mamdr_program_end), else NULL
    void* dbg_timing;  // debug: phase time stamps of the pass kernel (mamdr_debug_pass_timing)
    long long dbg_timing_cap;
    char  err[512];
};

// optimizer / step state, device resident (see mamdr_opt_state_* in the header)
struct OptState {
    long long    step;    // number of optimizer applies so far (dropout global step)
    float        b1pow;   // beta1^(step+1), fp32 running product like TF's beta1_power variable
    float        b2pow;
    unsigned int ticket;  // last-block-done counter of the optimizer sweep
    unsigned int pad[3];
};

// one id list to de-duplicate (scatter.cu); two jobs (user + item table of a mini-batch) share the launches
struct DedupJob {
    const int32_t* ids;        // [n]
    const float*   grad_rows;  // [n, grad_stride]
    int64_t        grad_stride;
    int            dim;
    int32_t*       uniq_ids;   // [n]
    float*         uniq_rows;  // [n, dim]
    int32_t*       n_uniq;     // [1]
    int32_t*       perm;       // [n]   workspace (mamdr_scatter_job_ws)
    int32_t*       seg_start;  // [n+1] workspace
};
struct DedupArgs { DedupJob job[2]; int n, npow2; };
int  mamdr_scatter_dedup_jobs(mamdr_ctx* ctx, const DedupJob* jobs, int n_jobs, int n, cudaStream_t st);
void mamdr_scatter_job_ws(DedupJob* job, void* ws, int64_t n);

extern char g_mamdr_create_err[512];

#define MAMDR_SET_ERR(ctx, ...)                                         \
    do {                                                                \
        if (ctx) snprintf((ctx)->err, sizeof((ctx)->err), __VA_ARGS__); \
    } while (0)

#define MAMDR_REQUIRE(ctx, cond, code, ...)  \
    do {                                     \
        if (!(cond)) {                       \
            MAMDR_SET_ERR(ctx, __VA_ARGS__); \
            return (code);                   \
        }                                    \
    } while (0)

#define MAMDR_CUDA_OK(ctx, expr)                                                              \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            MAMDR_SET_ERR(ctx, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                  \
                          cudaGetErrorString(e__));                                           \
            return MAMDR_E_CUDA;                                                              \
        }                                                                                     \
    } while (0)

#define MAMDR_LAUNCH_OK(ctx) MAMDR_CUDA_OK(ctx, cudaGetLastError())

// Keras clips p to [1e-7, 1 - 1e-7] before the BCE, which zeroes the gradient of saturated rows.  In exact arithmetic
// that is |logit| <= ln((1 - 1e-7) / 1e-7); deciding it on the fp32 LOGIT (instead of on p, whose spacing next to 1 is
// 6e-8, i.e. a 0.7-wide band of logits) makes the indicator robust to the last ulp of expf.  oracle/mlp.py: LOGIT_CLIP.
#define MAMDR_LOGIT_CLIP 16.118095f

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// streaming (evict-first) 128-bit accesses for one-touch sweeps
__device__ __forceinline__ float4 ld_stream_f4(const float* p) {
    return __ldcs(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ void st_stream_f4(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }
