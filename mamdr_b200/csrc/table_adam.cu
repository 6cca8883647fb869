// K6 + K7 for trainable embedding tables: ONE fused sweep per table and mini-batch that merges the de-duplicated
// sparse gradient rows with the dense L2 term and applies the (non-lazy) Adam update to every row.
//
// Replaces what TF does for an Embedding variable with `embeddings_regularizer=l2(1e-5)` under
// train.AdamOptimizer (/root/reference/model_zoo/DeepCTR/deepctr.py:54-55,104-126): the
// IndexedSlices gradient of tf.gather is de-duplicated (tf.unique + unsorted_segment_sum), aggregated with the
// dense 2*l2*E regulariser gradient, and the Adam update runs over the WHOLE tensor (m, v decay and parameter
// step on every row, every step).  Also accumulates l2 * sum(E^2) (the table's term of the Keras loss) from the
// pre-update values, so the loss needs no second pass over the table.
//
// HBM-bound: 24 B per element (read p, m, v; write p, m, v) + 4 B per row of slot map; one warp per row
// (dim 128 = one float4 per lane per array), streaming loads / stores, ONE resident wave: 3 CTAs x 256 threads per SM
// (79 registers), grid-stride over the rows -- measured 5 983 GB/s vs 5 631 GB/s with 8 CTAs per SM (2.67 waves).
// Deterministic: fixed grid, fixed-order reductions, no float atomics.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxBlocks = 148 * 8 * 2;
constexpr int kTableCtasPerSm = 3, kTableWaves = 1;   // grid of the table sweep = SMs x this (A/B: profiles/r1_v5_table_sweep_grid.txt)

__global__ void __launch_bounds__(256)
slot_scatter_kernel(const int32_t* __restrict__ uniq_ids, const int32_t* __restrict__ n_uniq, int32_t* __restrict__ slot) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n_uniq[0]) slot[uniq_ids[k]] = k;
}

__device__ __forceinline__ void adam1(float& p, float& m, float& v, float g, float alpha, float omb1, float omb2, float eps) {
    m = __fadd_rn(m, __fmul_rn(__fsub_rn(g, m), omb1));
    v = __fadd_rn(v, __fmul_rn(__fsub_rn(__fmul_rn(g, g), v), omb2));
    p = __fsub_rn(p, __fdiv_rn(__fmul_rn(m, alpha), __fadd_rn(__fsqrt_rn(v), eps)));
}

struct TableArgs {
    float *p, *m, *v;
    long long rows;
    int dim;
    const float* uniq_rows;
    int32_t* slot;
    const OptState* st;
    float l2, lr, beta1, beta2, eps;
    double* sq_part;          // [gridDim.x]
    unsigned int* ticket;
    float* loss;              // optional: += l2 * sum(E^2)
};

// SGD = true: the finetune stage's GradientDescentOptimizer (specific_base_model.py:120, base_model.py:69) on a trainable table:
// p -= (2*l2*p + sparse) * lr on every row; m / v are not touched (8 B per element instead of 24).
template <bool SGD>
__global__ void __launch_bounds__(kThreads) adam_table_kernel(TableArgs a) {
    const float b1p = SGD ? 0.f : a.st->b1pow, b2p = SGD ? 0.f : a.st->b2pow;
    const float alpha = __fdiv_rn(__fmul_rn(a.lr, __fsqrt_rn(__fsub_rn(1.0f, b2p))), __fsub_rn(1.0f, b1p));
    const float omb1 = __fsub_rn(1.0f, a.beta1), omb2 = __fsub_rn(1.0f, a.beta2);
    const float two_l2 = 2.0f * a.l2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long gwarp = (long long)blockIdx.x * (kThreads / 32) + warp;
    const long long nwarps = (long long)gridDim.x * (kThreads / 32);
    double sq = 0.0;
    // two rows per warp and trip: six 512-byte loads in flight per warp before the first dependent instruction
    for (long long row0 = gwarp; row0 < a.rows; row0 += 2 * nwarps) {
        const long long row1 = row0 + nwarps;
        const bool has1 = row1 < a.rows;
        const int slot0 = a.slot[row0], slot1 = has1 ? a.slot[row1] : -1;
        for (int c = lane * 4; c < a.dim; c += 128) {
            const long long o0 = row0 * a.dim + c, o1 = row1 * a.dim + c;
            float4 P[2], M[2], V[2], S[2];
            P[0] = ld_stream_f4(a.p + o0);
            if (!SGD) { M[0] = ld_stream_f4(a.m + o0); V[0] = ld_stream_f4(a.v + o0); }
            if (has1) {
                P[1] = ld_stream_f4(a.p + o1);
                if (!SGD) { M[1] = ld_stream_f4(a.m + o1); V[1] = ld_stream_f4(a.v + o1); }
            }
            S[0] = slot0 >= 0 ? ldg_f4(a.uniq_rows + (long long)slot0 * a.dim + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            S[1] = slot1 >= 0 ? ldg_f4(a.uniq_rows + (long long)slot1 * a.dim + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                if (r == 1 && !has1) break;
                const int slot = r == 0 ? slot0 : slot1;
                float4 G = make_float4(__fmul_rn(two_l2, P[r].x), __fmul_rn(two_l2, P[r].y), __fmul_rn(two_l2, P[r].z), __fmul_rn(two_l2, P[r].w));
                if (slot >= 0) {
                    G.x = __fadd_rn(G.x, S[r].x); G.y = __fadd_rn(G.y, S[r].y); G.z = __fadd_rn(G.z, S[r].z); G.w = __fadd_rn(G.w, S[r].w);
                }
                sq += (double)P[r].x * P[r].x + (double)P[r].y * P[r].y + (double)P[r].z * P[r].z + (double)P[r].w * P[r].w;
                const long long o = r == 0 ? o0 : o1;
                if (SGD) {
                    P[r].x = __fsub_rn(P[r].x, __fmul_rn(G.x, a.lr)); P[r].y = __fsub_rn(P[r].y, __fmul_rn(G.y, a.lr));
                    P[r].z = __fsub_rn(P[r].z, __fmul_rn(G.z, a.lr)); P[r].w = __fsub_rn(P[r].w, __fmul_rn(G.w, a.lr));
                    st_stream_f4(a.p + o, P[r]);
                } else {
                    adam1(P[r].x, M[r].x, V[r].x, G.x, alpha, omb1, omb2, a.eps);
                    adam1(P[r].y, M[r].y, V[r].y, G.y, alpha, omb1, omb2, a.eps);
                    adam1(P[r].z, M[r].z, V[r].z, G.z, alpha, omb1, omb2, a.eps);
                    adam1(P[r].w, M[r].w, V[r].w, G.w, alpha, omb1, omb2, a.eps);
                    st_stream_f4(a.p + o, P[r]);
                    st_stream_f4(a.m + o, M[r]);
                    st_stream_f4(a.v + o, V[r]);
                }
            }
        }
        if (lane == 0) {   // leave the map clean for the next mini-batch
            if (slot0 >= 0) a.slot[row0] = -1;
            if (slot1 >= 0) a.slot[row1] = -1;
        }
    }
    // ---- l2 * sum(E^2): warp shuffle -> block (fixed order) -> last block sums the per-block partials in order
    __shared__ double wsum[kThreads / 32];
    __shared__ bool last;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if (lane == 0) wsum[warp] = sq;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < kThreads / 32; ++w) s += wsum[w];
        a.sq_part[blockIdx.x] = s;
        __threadfence();
        last = (atomicAdd(a.ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        double s = 0.0;
        for (unsigned int b = 0; b < gridDim.x; ++b) s += __ldcg(a.sq_part + b);
        if (a.loss) a.loss[0] = (float)((double)a.loss[0] + (double)a.l2 * s);
        a.sq_part[gridDim.x] = s;   // kept for callers that want the raw sum
        *a.ticket = 0;
    }
}

// sum of squares in double, fixed order (per-thread strided, warp shuffle, per-block partial, last block in order)
__global__ void __launch_bounds__(kThreads)
sumsq_kernel(const float* __restrict__ x, long long n4, double* __restrict__ part, unsigned int* __restrict__ ticket, double* __restrict__ out) {
    double sq = 0.0;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n4; i += (long long)gridDim.x * kThreads) {
        const float4 v = ld_stream_f4(x + 4 * i);
        sq += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    }
    __shared__ double wsum[kThreads / 32];
    __shared__ bool last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if (lane == 0) wsum[warp] = sq;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < kThreads / 32; ++w) s += wsum[w];
        part[blockIdx.x] = s;
        __threadfence();
        last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        double s = 0.0;
        for (unsigned int b = 0; b < gridDim.x; ++b) s += __ldcg(part + b);
        out[0] = s;
        *ticket = 0;
    }
}

}  // namespace

extern "C" int mamdr_sum_squares_f64(mamdr_ctx* ctx, const float* x_dev, int64_t n, double* out_dev, void* ws_dev, size_t ws_bytes,
                                     mamdr_stream stream) {
    MAMDR_REQUIRE(ctx, ctx && x_dev && out_dev && ws_dev, MAMDR_E_INVALID, "NULL pointer");
    MAMDR_REQUIRE(ctx, n > 0 && n % 4 == 0 && aligned16(x_dev) && aligned16(ws_dev), MAMDR_E_INVALID, "n must be a positive multiple of 4, pointers 16-byte aligned");
    MAMDR_REQUIRE(ctx, ws_bytes >= (size_t)(kMaxBlocks + 2) * sizeof(double) + 64, MAMDR_E_WORKSPACE, "workspace too small");
    long long want = ((n >> 2) + kThreads - 1) / kThreads;
    long long cap = (long long)ctx->sm_count * 8;
    if (cap > kMaxBlocks) cap = kMaxBlocks;
    const int grid = (int)(want < cap ? want : cap);
    sumsq_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(x_dev, n >> 2, (double*)ws_dev,
                                                            (unsigned int*)((unsigned char*)ws_dev + (size_t)(kMaxBlocks + 2) * sizeof(double)), out_dev);
    MAMDR_LAUNCH_OK(ctx);
    return MAMDR_OK;
}

extern "C" size_t mamdr_adam_table_workspace_bytes(void) { return (size_t)(kMaxBlocks + 2) * sizeof(double) + 64; }

static int table_step(mamdr_ctx* ctx, bool sgd, float* table_dev, float* m_dev, float* v_dev, int64_t rows, int32_t dim,
                      const int32_t* uniq_ids_dev, const float* uniq_rows_dev, const int32_t* n_uniq_dev,
                      int64_t max_uniq, int32_t* slot_map_dev, float l2, const void* opt_state_dev, float lr,
                      float beta1, float beta2, float eps, float* loss_dev, void* ws_dev, size_t ws_bytes,
                      mamdr_stream stream) {
    MAMDR_REQUIRE(ctx, ctx != nullptr, MAMDR_E_INVALID, "ctx is NULL");
    MAMDR_REQUIRE(ctx, table_dev && slot_map_dev && ws_dev && (sgd || (m_dev && v_dev && opt_state_dev)), MAMDR_E_INVALID, "NULL pointer");
    MAMDR_REQUIRE(ctx, rows > 0 && dim > 0 && dim % 4 == 0, MAMDR_E_INVALID, "bad table shape");
    MAMDR_REQUIRE(ctx, aligned16(table_dev) && aligned16(m_dev) && aligned16(v_dev) && aligned16(ws_dev), MAMDR_E_INVALID, "misaligned pointer");
    MAMDR_REQUIRE(ctx, ws_bytes >= mamdr_adam_table_workspace_bytes(), MAMDR_E_WORKSPACE, "workspace too small");
    MAMDR_REQUIRE(ctx, max_uniq == 0 || (uniq_ids_dev && uniq_rows_dev && n_uniq_dev && aligned16(uniq_rows_dev)), MAMDR_E_INVALID,
                  "sparse gradient pointers NULL or misaligned");
    cudaStream_t st = (cudaStream_t)stream;
    if (max_uniq > 0) {
        slot_scatter_kernel<<<(unsigned)((max_uniq + 255) / 256), 256, 0, st>>>(uniq_ids_dev, n_uniq_dev, slot_map_dev);
        MAMDR_LAUNCH_OK(ctx);
    }
    TableArgs a;
    a.p = table_dev; a.m = m_dev; a.v = v_dev; a.rows = rows; a.dim = dim;
    a.uniq_rows = uniq_rows_dev; a.slot = slot_map_dev; a.st = (const OptState*)opt_state_dev;
    a.l2 = l2; a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps;
    a.sq_part = (double*)ws_dev;
    a.ticket = (unsigned int*)((unsigned char*)ws_dev + (size_t)(kMaxBlocks + 2) * sizeof(double));
    a.loss = loss_dev;
    long long want = (rows + (kThreads / 32) - 1) / (kThreads / 32);
    // grid = a whole number of waves: kTableCtasPerSm resident CTAs per SM (79 registers x 256 threads -> 3) x waves
    int per_sm = kTableCtasPerSm * kTableWaves;
    if (const char* e = getenv("MAMDR_TABLE_CTAS_PER_SM")) per_sm = atoi(e) > 0 ? atoi(e) : per_sm;   // tuning knob (tests/diag_table_sweep.py)
    long long cap = (long long)ctx->sm_count * per_sm;
    if (cap > kMaxBlocks) cap = kMaxBlocks;
    const int grid = (int)(want < cap ? want : cap);
    if (sgd) adam_table_kernel<true><<<grid, kThreads, 0, st>>>(a);
    else adam_table_kernel<false><<<grid, kThreads, 0, st>>>(a);
    MAMDR_LAUNCH_OK(ctx);
    return MAMDR_OK;
}

extern "C" int mamdr_adam_table_step(mamdr_ctx* ctx, float* table_dev, float* m_dev, float* v_dev, int64_t rows, int32_t dim,
                                     const int32_t* uniq_ids_dev, const float* uniq_rows_dev, const int32_t* n_uniq_dev,
                                     int64_t max_uniq, int32_t* slot_map_dev, float l2, const void* opt_state_dev, float lr,
                                     float beta1, float beta2, float eps, float* loss_dev, void* ws_dev, size_t ws_bytes,
                                     mamdr_stream stream) {
    return table_step(ctx, false, table_dev, m_dev, v_dev, rows, dim, uniq_ids_dev, uniq_rows_dev, n_uniq_dev, max_uniq, slot_map_dev, l2,
                      opt_state_dev, lr, beta1, beta2, eps, loss_dev, ws_dev, ws_bytes, stream);
}

extern "C" int mamdr_sgd_table_step(mamdr_ctx* ctx, float* table_dev, int64_t rows, int32_t dim, const int32_t* uniq_ids_dev,
                                    const float* uniq_rows_dev, const int32_t* n_uniq_dev, int64_t max_uniq, int32_t* slot_map_dev,
                                    float l2, float lr, float* loss_dev, void* ws_dev, size_t ws_bytes, mamdr_stream stream) {
    return table_step(ctx, true, table_dev, nullptr, nullptr, rows, dim, uniq_ids_dev, uniq_rows_dev, n_uniq_dev, max_uniq, slot_map_dev, l2,
                      nullptr, lr, 0.f, 0.f, 0.f, loss_dev, ws_dev, ws_bytes, stream);
}
