// K7, K9, K10: multi-tensor optimizer apply and DN / DR meta updates as single coalesced sweeps
// over flat parameter arenas (every trainable tensor of the model lives in ONE fp32 buffer, so
// "multi-tensor" is one 128-bit vectorised grid-stride loop; HBM-bound at Amazon sizes, L2-resident
// at Taobao sizes).
//
// Replaces: tf.train.AdamOptimizer.apply_gradients (/root/reference/model_zoo/DeepCTR/deepctr.py:54-55),
// SetVarOp / K.batch_get_value round trips (utils/tool.py:36-45, model_zoo/maml.py:181-194), and the
// host numpy algebra of model_zoo/domain_negotiation.py:118-123, model_zoo/mamdr.py:168-196,
// model_zoo/specific_base_model.py:164-172.
//
// Arithmetic is written with explicit round-to-nearest intrinsics (__fmul_rn/__fadd_rn/...) in the
// operation order of TF's ApplyAdam kernel and of the reference's numpy expressions, so each op is
// bit-exact against the fp32 numpy oracle for identical inputs (no FMA contraction).
#include "common.cuh"
#include "meta_ops.cuh"
#include "program.cuh"

namespace {

constexpr int kThreads = 256;

inline int sweep_grid(const mamdr_ctx* ctx, int64_t n_vec) {
    const int64_t want = ceil_div64(n_vec, kThreads);
    const int64_t cap = (int64_t)ctx->sm_count * 8;
    return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

__global__ void opt_state_init_kernel(OptState* s, float b1, float b2) {
    s->step = 0;
    s->b1pow = b1;
    s->b2pow = b2;
    s->ticket = 0;
    s->pad[0] = s->pad[1] = s->pad[2] = 0;
}

// last-block-done: every block reads the beta powers before it arrives; the block that draws the
// final ticket advances them (and the step) after all others have read.
__device__ __forceinline__ void finish_step(OptState* st, float beta1, float beta2, bool adam) {
    __shared__ bool last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(&st->ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (last && threadIdx.x == 0) {
        if (adam) {
            st->b1pow = __fmul_rn(st->b1pow, beta1);
            st->b2pow = __fmul_rn(st->b2pow, beta2);
        }
        st->step += 1;
        st->ticket = 0;
    }
}

__device__ __forceinline__ void adam1(float& p, float& m, float& v, float g, float alpha, float omb1, float omb2,
                                      float eps) {
    m = __fadd_rn(m, __fmul_rn(__fsub_rn(g, m), omb1));
    v = __fadd_rn(v, __fmul_rn(__fsub_rn(__fmul_rn(g, g), v), omb2));
    p = __fsub_rn(p, __fdiv_rn(__fmul_rn(m, alpha), __fadd_rn(__fsqrt_rn(v), eps)));
}

__global__ void __launch_bounds__(kThreads)
adam_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ g,
            int64_t n, OptState* st, float lr, float beta1, float beta2, float eps) {
    const float b1p = st->b1pow, b2p = st->b2pow;
    const float alpha = __fdiv_rn(__fmul_rn(lr, __fsqrt_rn(__fsub_rn(1.0f, b2p))), __fsub_rn(1.0f, b1p));
    const float omb1 = __fsub_rn(1.0f, beta1), omb2 = __fsub_rn(1.0f, beta2);
    const int64_t nv = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < nv; i += stride) {
        float4 P = *reinterpret_cast<float4*>(p + 4 * i), M = *reinterpret_cast<float4*>(m + 4 * i);
        float4 V = *reinterpret_cast<float4*>(v + 4 * i);
        const float4 G = *reinterpret_cast<const float4*>(g + 4 * i);
        adam1(P.x, M.x, V.x, G.x, alpha, omb1, omb2, eps);
        adam1(P.y, M.y, V.y, G.y, alpha, omb1, omb2, eps);
        adam1(P.z, M.z, V.z, G.z, alpha, omb1, omb2, eps);
        adam1(P.w, M.w, V.w, G.w, alpha, omb1, omb2, eps);
        *reinterpret_cast<float4*>(p + 4 * i) = P;
        *reinterpret_cast<float4*>(m + 4 * i) = M;
        *reinterpret_cast<float4*>(v + 4 * i) = V;
    }
    finish_step(st, beta1, beta2, true);
}

// ApplyAdam over a handful of arena ranges (the variables of ONE sub-model of a multi-task tower): everything outside
// the ranges keeps its value and its slots; the beta powers advance once.
constexpr int kMaxRanges = 16;
struct RangeArgs {
    int64_t begin4[kMaxRanges];   // first float4 of range q
    int64_t cum4[kMaxRanges + 1]; // float4s before range q
    int     n;
};

__global__ void __launch_bounds__(kThreads)
adam_ranges_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ g,
                   const __grid_constant__ RangeArgs r, OptState* st, float lr, float beta1, float beta2, float eps) {
    const float b1p = st->b1pow, b2p = st->b2pow;
    const float alpha = __fdiv_rn(__fmul_rn(lr, __fsqrt_rn(__fsub_rn(1.0f, b2p))), __fsub_rn(1.0f, b1p));
    const float omb1 = __fsub_rn(1.0f, beta1), omb2 = __fsub_rn(1.0f, beta2);
    const int64_t nv = r.cum4[r.n];
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x; t < nv; t += stride) {
        int q = 0;
        while (q + 1 < r.n && t >= r.cum4[q + 1]) ++q;
        const int64_t i = r.begin4[q] + (t - r.cum4[q]);
        float4 P = *reinterpret_cast<float4*>(p + 4 * i), M = *reinterpret_cast<float4*>(m + 4 * i);
        float4 V = *reinterpret_cast<float4*>(v + 4 * i);
        const float4 G = *reinterpret_cast<const float4*>(g + 4 * i);
        adam1(P.x, M.x, V.x, G.x, alpha, omb1, omb2, eps);
        adam1(P.y, M.y, V.y, G.y, alpha, omb1, omb2, eps);
        adam1(P.z, M.z, V.z, G.z, alpha, omb1, omb2, eps);
        adam1(P.w, M.w, V.w, G.w, alpha, omb1, omb2, eps);
        *reinterpret_cast<float4*>(p + 4 * i) = P;
        *reinterpret_cast<float4*>(m + 4 * i) = M;
        *reinterpret_cast<float4*>(v + 4 * i) = V;
    }
    finish_step(st, beta1, beta2, true);
}

__global__ void __launch_bounds__(kThreads)
sgd_kernel(float* __restrict__ p, const float* __restrict__ g, int64_t n, OptState* st, float lr) {
    const int64_t nv = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < nv; i += stride) {
        float4 P = *reinterpret_cast<float4*>(p + 4 * i);
        const float4 G = *reinterpret_cast<const float4*>(g + 4 * i);
        P.x = __fsub_rn(P.x, __fmul_rn(G.x, lr));
        P.y = __fsub_rn(P.y, __fmul_rn(G.y, lr));
        P.z = __fsub_rn(P.z, __fmul_rn(G.z, lr));
        P.w = __fsub_rn(P.w, __fmul_rn(G.w, lr));
        *reinterpret_cast<float4*>(p + 4 * i) = P;
    }
    finish_step(st, 0.f, 0.f, false);
}

// ---- element-wise meta sweeps (meta_ops.cuh) ---------------------------------------------------------
template <int OP>
__global__ void __launch_bounds__(kThreads) meta_kernel(MetaArgs a) {
    const int64_t nv = a.n >> 2;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < nv; i += stride) meta_float4(OP, a, i);
}

template <int OP>
int launch_meta(mamdr_ctx* ctx, MetaArgs a, mamdr_stream stream) {
    MAMDR_REQUIRE(ctx, ctx != nullptr, MAMDR_E_INVALID, "ctx is NULL");
    MAMDR_REQUIRE(ctx, a.n >= 0 && a.n % 4 == 0, MAMDR_E_INVALID, "arena length must be a multiple of 4 floats");
    if (a.n == 0) return MAMDR_OK;
    const void* ptrs[5] = {a.w0, a.w1, a.r0, a.r1, a.r2};
    for (int i = 0; i < 5; ++i) MAMDR_REQUIRE(ctx, aligned16(ptrs[i]), MAMDR_E_INVALID, "arena pointer misaligned");
    if (mamdr_prog_recording(ctx)) return mamdr_prog_push_meta(ctx, OP, a);   // deferred into the program kernel
    meta_kernel<OP><<<sweep_grid(ctx, a.n >> 2), kThreads, 0, (cudaStream_t)stream>>>(a);
    MAMDR_LAUNCH_OK(ctx);
    return MAMDR_OK;
}

}  // namespace

// run-time dispatch used when a recorded program holds no pass (program.cuh)
int mamdr_meta_launch(mamdr_ctx* ctx, int meta_op, const MetaArgs& a, mamdr_stream stream) {
    switch (meta_op) {
        case OP_COPY: return launch_meta<OP_COPY>(ctx, a, stream);
        case OP_MERGE: return launch_meta<OP_MERGE>(ctx, a, stream);
        case OP_DN: return launch_meta<OP_DN>(ctx, a, stream);
        case OP_DR: return launch_meta<OP_DR>(ctx, a, stream);
        case OP_DR_ACC: return launch_meta<OP_DR_ACC>(ctx, a, stream);
        case OP_DR_APPLY: return launch_meta<OP_DR_APPLY>(ctx, a, stream);
        case OP_SUB: return launch_meta<OP_SUB>(ctx, a, stream);
        case OP_AXPY_DIFF: return launch_meta<OP_AXPY_DIFF>(ctx, a, stream);
    }
    MAMDR_SET_ERR(ctx, "unknown meta op %d", meta_op);
    return MAMDR_E_INVALID;
}

extern "C" size_t mamdr_opt_state_bytes(void) { return sizeof(OptState); }

extern "C" int mamdr_opt_state_init(mamdr_ctx* ctx, void* state, float beta1, float beta2, mamdr_stream stream) {
    MAMDR_REQUIRE(ctx, ctx && state, MAMDR_E_INVALID, "NULL ctx/state");
    opt_state_init_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((OptState*)state, beta1, beta2);
    MAMDR_LAUNCH_OK(ctx);
    return MAMDR_OK;
}

extern "C" int mamdr_opt_state_read(mamdr_ctx* ctx, const void* state, int64_t* step, float* b1pow, float* b2pow,
                                    mamdr_stream stream) {
    MAMDR_REQUIRE(ctx, ctx && state, MAMDR_E_INVALID, "NULL ctx/state");
    OptState h;
    MAMDR_CUDA_OK(ctx, cudaMemcpyAsync(&h, state, sizeof(h), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    MAMDR_CUDA_OK(ctx, cudaStreamSynchronize((cudaStream_t)stream));
    if (step) *step = h.step;
    if (b1pow) *b1pow = h.b1pow;
    if (b2pow) *b2pow = h.b2pow;
    return MAMDR_OK;
}

extern "C" int mamdr_adam_step(mamdr_ctx* ctx, float* p, float* m, float* v, const float* g, int64_t n, void* state,
                               float lr, float beta1, float beta2, float eps, mamdr_stream stream) {
    MAMDR_REQUIRE(ctx, ctx != nullptr, MAMDR_E_INVALID, "ctx is NULL");
    MAMDR_REQUIRE(ctx, p && m && v && g && state, MAMDR_E_INVALID, "NULL pointer");
    MAMDR_REQUIRE(ctx, !mamdr_prog_recording(ctx), MAMDR_E_INVALID, "mamdr_adam_step cannot be recorded into a program");
    MAMDR_REQUIRE(ctx, n > 0 && n % 4 == 0, MAMDR_E_INVALID, "arena length must be a positive multiple of 4");
    MAMDR_REQUIRE(ctx, aligned16(p) && aligned16(m) && aligned16(v) && aligned16(g), MAMDR_E_INVALID, "misaligned arena");
    adam_kernel<<<sweep_grid(ctx, n >> 2), kThreads, 0, (cudaStream_t)stream>>>(p, m, v, g, n, (OptState*)state, lr,
                                                                               beta1, beta2, eps);
    MAMDR_LAUNCH_OK(ctx);
    return MAMDR_OK;
}

extern "C" int mamdr_adam_ranges_step(mamdr_ctx* ctx, float* p, float* m, float* v, const float* g, const int64_t* begin,
                                      const int64_t* len, int32_t n_ranges, void* state, float lr, float beta1, float beta2,
                                      float eps, mamdr_stream stream) {
    MAMDR_REQUIRE(ctx, ctx != nullptr, MAMDR_E_INVALID, "ctx is NULL");
    MAMDR_REQUIRE(ctx, p && m && v && g && state && begin && len, MAMDR_E_INVALID, "NULL pointer");
    MAMDR_REQUIRE(ctx, !mamdr_prog_recording(ctx), MAMDR_E_INVALID, "mamdr_adam_ranges_step cannot be recorded into a program");
    MAMDR_REQUIRE(ctx, n_ranges >= 1 && n_ranges <= kMaxRanges, MAMDR_E_INVALID, "n_ranges must be in 1..%d", kMaxRanges);
    MAMDR_REQUIRE(ctx, aligned16(p) && aligned16(m) && aligned16(v) && aligned16(g), MAMDR_E_INVALID, "misaligned arena");
    RangeArgs r;
    memset(&r, 0, sizeof(r));
    r.n = n_ranges;
    for (int q = 0; q < n_ranges; ++q) {
        MAMDR_REQUIRE(ctx, begin[q] >= 0 && len[q] > 0 && begin[q] % 4 == 0 && len[q] % 4 == 0, MAMDR_E_INVALID,
                      "range %d: begin / len must be non-negative multiples of 4", q);
        r.begin4[q] = begin[q] >> 2;
        r.cum4[q + 1] = r.cum4[q] + (len[q] >> 2);
    }
    adam_ranges_kernel<<<sweep_grid(ctx, r.cum4[n_ranges]), kThreads, 0, (cudaStream_t)stream>>>(p, m, v, g, r, (OptState*)state, lr, beta1,
                                                                                                beta2, eps);
    MAMDR_LAUNCH_OK(ctx);
    return MAMDR_OK;
}

extern "C" int mamdr_sgd_step(mamdr_ctx* ctx, float* p, const float* g, int64_t n, void* state, float lr,
                              mamdr_stream stream) {
    MAMDR_REQUIRE(ctx, ctx != nullptr, MAMDR_E_INVALID, "ctx is NULL");
    MAMDR_REQUIRE(ctx, p && g && state, MAMDR_E_INVALID, "NULL pointer");
    MAMDR_REQUIRE(ctx, n > 0 && n % 4 == 0, MAMDR_E_INVALID, "arena length must be a positive multiple of 4");
    MAMDR_REQUIRE(ctx, aligned16(p) && aligned16(g), MAMDR_E_INVALID, "misaligned arena");
    sgd_kernel<<<sweep_grid(ctx, n >> 2), kThreads, 0, (cudaStream_t)stream>>>(p, g, n, (OptState*)state, lr);
    MAMDR_LAUNCH_OK(ctx);
    return MAMDR_OK;
}

extern "C" int mamdr_copy(mamdr_ctx* ctx, float* dst, const float* src, int64_t n, mamdr_stream stream) {
    MAMDR_REQUIRE(ctx, ctx && dst && src, MAMDR_E_INVALID, "NULL pointer");
    MetaArgs a{dst, nullptr, src, nullptr, nullptr, 0.f, 0.f, 0, n};
    return launch_meta<OP_COPY>(ctx, a, stream);
}

extern "C" int mamdr_merge(mamdr_ctx* ctx, float* out, const float* theta, const float* theta_i, int64_t n,
                           int32_t method, mamdr_stream stream) {
    MAMDR_REQUIRE(ctx, ctx && out && theta && theta_i, MAMDR_E_INVALID, "NULL pointer");
    MAMDR_REQUIRE(ctx, method == MAMDR_MERGE_PLUS || method == MAMDR_MERGE_TIMES, MAMDR_E_INVALID, "bad merged_method");
    MetaArgs a{out, nullptr, theta, theta_i, nullptr, 0.f, 0.f, method, n};
    return launch_meta<OP_MERGE>(ctx, a, stream);
}

extern "C" int mamdr_dn_update(mamdr_ctx* ctx, float* theta, const float* model, float beta, int64_t n,
                               float* model_out, mamdr_stream stream) {
    MAMDR_REQUIRE(ctx, ctx && theta && model, MAMDR_E_INVALID, "NULL pointer");
    MetaArgs a{theta, model_out, model, nullptr, nullptr, beta, 0.f, 0, n};
    return launch_meta<OP_DN>(ctx, a, stream);
}

extern "C" int mamdr_dr_update(mamdr_ctx* ctx, float* theta_i, const float* theta, const float* model, float beta,
                               int64_t n, int32_t method, float* model_out, mamdr_stream stream) {
    MAMDR_REQUIRE(ctx, ctx && theta_i && theta && model, MAMDR_E_INVALID, "NULL pointer");
    MAMDR_REQUIRE(ctx, method == MAMDR_MERGE_PLUS || method == MAMDR_MERGE_TIMES, MAMDR_E_INVALID, "bad merged_method");
    MetaArgs a{theta_i, model_out, model, theta, nullptr, beta, 0.f, method, n};
    return launch_meta<OP_DR>(ctx, a, stream);
}

extern "C" int mamdr_dr_accumulate(mamdr_ctx* ctx, float* accum, const float* model, const float* theta,
                                   const float* theta_i, int64_t n, int32_t method, mamdr_stream stream) {
    MAMDR_REQUIRE(ctx, ctx && accum && model && theta && theta_i, MAMDR_E_INVALID, "NULL pointer");
    MAMDR_REQUIRE(ctx, method == MAMDR_MERGE_PLUS || method == MAMDR_MERGE_TIMES, MAMDR_E_INVALID, "bad merged_method");
    MetaArgs a{accum, nullptr, model, theta, theta_i, 0.f, 0.f, method, n};
    return launch_meta<OP_DR_ACC>(ctx, a, stream);
}

extern "C" int mamdr_dr_apply_accum(mamdr_ctx* ctx, float* theta_i, float* accum, float sample_num, float beta,
                                    int64_t n, mamdr_stream stream) {
    MAMDR_REQUIRE(ctx, ctx && theta_i && accum, MAMDR_E_INVALID, "NULL pointer");
    MAMDR_REQUIRE(ctx, sample_num != 0.f, MAMDR_E_INVALID, "sample_num is 0");
    MetaArgs a{theta_i, accum, nullptr, nullptr, nullptr, sample_num, beta, 0, n};
    return launch_meta<OP_DR_APPLY>(ctx, a, stream);
}

extern "C" int mamdr_sub(mamdr_ctx* ctx, float* out, const float* a_, const float* b_, int64_t n, mamdr_stream stream) {
    MAMDR_REQUIRE(ctx, ctx && out && a_ && b_, MAMDR_E_INVALID, "NULL pointer");
    MetaArgs a{out, nullptr, a_, b_, nullptr, 0.f, 0.f, 0, n};
    return launch_meta<OP_SUB>(ctx, a, stream);
}

extern "C" int mamdr_axpy_diff(mamdr_ctx* ctx, float* out, const float* a_, const float* b_, float alpha, int64_t n,
                               mamdr_stream stream) {
    MAMDR_REQUIRE(ctx, ctx && out && a_ && b_, MAMDR_E_INVALID, "NULL pointer");
    MetaArgs a{out, nullptr, a_, b_, nullptr, alpha, 0.f, 0, n};
    return launch_meta<OP_AXPY_DIFF>(ctx, a, stream);
}

// ---- PCGrad's host-side projection (model_zoo/pcgrad.py:152-160) on the device: per variable of shape [rows, cols] (a 1-D
// variable is ONE row) and per row r:  dot = sum_c cur[r,c] * aux[r,c];  if dot > 0 (the reference's test -- it projects
// the AGREEING rows):  aux' = aux - (dot / ||cur[r]||_2) * cur[r]  (divided by the norm, not its square, as the reference
// does);  cur[r] += aux'.  `final_grads` IS `current_grads` in the reference (pcgrad.py:104 aliases the list), so the sum
// is accumulated into the buffer the next support domain projects against: one in/out buffer here.  One warp per row,
// lanes stride the columns, fixed-order shuffle reductions (no atomics: deterministic).
__global__ void __launch_bounds__(256)
pcgrad_project_kernel(float* __restrict__ cur, const float* __restrict__ aux, const int64_t rows, const int cols) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = warp; r < rows; r += nwarps) {
        float* c = cur + r * cols;
        const float* a = aux + r * cols;
        float dot = 0.f, sq = 0.f;
        for (int j = lane; j < cols; j += 32) {
            const float cv = c[j], av = __ldg(a + j);
            dot = __fadd_rn(dot, __fmul_rn(cv, av));
            sq = __fadd_rn(sq, __fmul_rn(cv, cv));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            dot = __fadd_rn(dot, __shfl_xor_sync(0xffffffffu, dot, o));
            sq = __fadd_rn(sq, __shfl_xor_sync(0xffffffffu, sq, o));
        }
        const bool project = dot > 0.f;
        const float coef = project ? __fdiv_rn(dot, __fsqrt_rn(sq)) : 0.f;
        for (int j = lane; j < cols; j += 32) {
            const float cv = c[j];
            float av = __ldg(a + j);
            if (project) av = __fsub_rn(av, __fmul_rn(coef, cv));
            c[j] = __fadd_rn(cv, av);
        }
    }
}

extern "C" int mamdr_pcgrad_project(mamdr_ctx* ctx, float* final_grads, const float* aux_grads, int64_t rows, int32_t cols,
                                    mamdr_stream stream) {
    MAMDR_REQUIRE(ctx, ctx && final_grads && aux_grads, MAMDR_E_INVALID, "NULL pointer");
    MAMDR_REQUIRE(ctx, rows > 0 && cols > 0, MAMDR_E_INVALID, "rows and cols must be positive");
    const int64_t want = (rows + 7) / 8;
    const int64_t cap = (int64_t)ctx->sm_count * 8;
    pcgrad_project_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(final_grads, aux_grads, rows, cols);
    MAMDR_LAUNCH_OK(ctx);
    return MAMDR_OK;
}
