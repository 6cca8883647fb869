// K1: 128-bit vectorised embedding gather, and the fused mini-batch assembly
// (gather user/item/domain rows + concat + label gather) that feeds the tower.
//
// Replaces tf.gather under the three Embedding lookups + Concatenate + Flatten of
// /root/reference/model_zoo/DeepCTR/deepctr.py:125-128 (DeepCTR input_from_feature_columns /
// combined_dnn_input).  HBM/L2-bound copy: one 16-byte load and one 16-byte store per thread per
// element group, consecutive lanes on consecutive 16-byte chunks of the same row (a dim-128 row is
// exactly one 512-byte warp access), 4 independent row loads in flight per thread.
#include "common.cuh"

namespace {

constexpr int kGatherThreads = 256;
constexpr int kGatherUnroll  = 4;

__global__ void __launch_bounds__(kGatherThreads)
gather_rows_f32_kernel(const float* __restrict__ table, const int32_t* __restrict__ ids, int64_t n,
                       int dv /* dim / 4 */, float* __restrict__ out, int64_t out_stride) {
    const int64_t total  = n * dv;
    const int64_t stride = (int64_t)gridDim.x * kGatherThreads;
    int64_t v = (int64_t)blockIdx.x * kGatherThreads + threadIdx.x;
    const int64_t row_floats = (int64_t)dv * 4;
    for (; v + (kGatherUnroll - 1) * stride < total; v += kGatherUnroll * stride) {
        float4  val[kGatherUnroll];
        int64_t r[kGatherUnroll];
        int     c[kGatherUnroll];
#pragma unroll
        for (int u = 0; u < kGatherUnroll; ++u) {
            const int64_t vv = v + u * stride;
            r[u] = vv / dv;
            c[u] = (int)(vv - r[u] * dv);
        }
#pragma unroll
        for (int u = 0; u < kGatherUnroll; ++u) {
            const int64_t id = __ldg(ids + r[u]);
            val[u] = ldg_f4(table + id * row_floats + c[u] * 4);
        }
#pragma unroll
        for (int u = 0; u < kGatherUnroll; ++u) st_stream_f4(out + r[u] * out_stride + c[u] * 4, val[u]);
    }
    for (; v < total; v += stride) {
        const int64_t r  = v / dv;
        const int     c  = (int)(v - r * dv);
        const int64_t id = __ldg(ids + r);
        st_stream_f4(out + r * out_stride + c * 4, ldg_f4(table + id * row_floats + c * 4));
    }
}

// X[r, :] = [E_u[uid[o]] | E_i[pid[o]] | E_d[domain]], y[r] = label[o], o = order ? order[off+r] : off+r.
// One warp per row; also emits the gathered ids (needed by the sparse embedding backward).
__global__ void __launch_bounds__(256)
assemble_batch_kernel(const float* __restrict__ Eu, const float* __restrict__ Ei,
                      const float* __restrict__ Ed, const int32_t* __restrict__ uid,
                      const int32_t* __restrict__ pid, const float* __restrict__ label,
                      const int32_t* __restrict__ order, int64_t offset, int rows, int domain, int du,
                      int di, int dd, float* __restrict__ X, float* __restrict__ y,
                      int32_t* __restrict__ uid_b, int32_t* __restrict__ pid_b) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int in_dim = du + di + dd;
    for (int r = warp; r < rows; r += nwarps) {
        const int64_t o = order ? (int64_t)__ldg(order + offset + r) : offset + r;
        const int64_t u = __ldg(uid + o), p = __ldg(pid + o);
        float* xr = X + (int64_t)r * in_dim;
        const float* su = Eu + u * du;
        const float* si = Ei + p * di;
        const float* sd = Ed + (int64_t)domain * dd;
        auto put = [&](int dst, const float* src) { *reinterpret_cast<float4*>(xr + dst) = ldg_f4(src); };
        for (int c = lane * 4; c < du; c += 128) put(c, su + c);
        for (int c = lane * 4; c < di; c += 128) put(du + c, si + c);
        for (int c = lane * 4; c < dd; c += 128) put(du + di + c, sd + c);
        if (lane == 0) {
            y[r] = __ldg(label + o);
            uid_b[r] = (int32_t)u;
            pid_b[r] = (int32_t)p;
        }
    }
}

}  // namespace

extern "C" int mamdr_gather_f32(mamdr_ctx* ctx, const float* table, int64_t rows, int32_t dim,
                                const int32_t* ids, int64_t n, float* out, int64_t out_stride,
                                mamdr_stream stream) {
    MAMDR_REQUIRE(ctx, ctx != nullptr, MAMDR_E_INVALID, "ctx is NULL");
    MAMDR_REQUIRE(ctx, n >= 0 && rows >= 0, MAMDR_E_INVALID, "negative size");
    if (n == 0) return MAMDR_OK;
    MAMDR_REQUIRE(ctx, table && ids && out, MAMDR_E_INVALID, "NULL pointer");
    MAMDR_REQUIRE(ctx, dim > 0 && dim % 4 == 0, MAMDR_E_INVALID, "dim must be a positive multiple of 4");
    MAMDR_REQUIRE(ctx, out_stride >= dim && out_stride % 4 == 0, MAMDR_E_INVALID, "bad out_stride");
    MAMDR_REQUIRE(ctx, aligned16(table) && aligned16(out), MAMDR_E_INVALID, "table/out must be 16-byte aligned");
    const int     dv    = dim / 4;
    const int64_t total = n * dv;
    const int64_t want  = ceil_div64(total, (int64_t)kGatherThreads * kGatherUnroll);
    const int64_t cap   = (int64_t)ctx->sm_count * 8;  // 8 CTAs x 256 threads = full occupancy
    const int     grid  = (int)(want < cap ? (want < 1 ? 1 : want) : cap);
    gather_rows_f32_kernel<<<grid, kGatherThreads, 0, (cudaStream_t)stream>>>(table, ids, n, dv, out, out_stride);
    MAMDR_LAUNCH_OK(ctx);
    return MAMDR_OK;
}

// internal (used by mlp.cu)
int mamdr_assemble_batch(mamdr_ctx* ctx, const float* Eu, const float* Ei, const float* Ed,
                         const mamdr_batch* b, int du, int di, int dd, float* X, float* y,
                         int32_t* uid_b, int32_t* pid_b, cudaStream_t stream) {
    const int warps_per_block = 8;
    const int grid = (b->rows + warps_per_block - 1) / warps_per_block;
    assemble_batch_kernel<<<grid, warps_per_block * 32, 0, stream>>>(
        Eu, Ei, Ed, b->uid_dev, b->pid_dev, b->label_dev, b->order_dev, b->offset, b->rows, b->domain, du,
        di, dd, X, y, uid_b, pid_b);
    MAMDR_LAUNCH_OK(ctx);
    return MAMDR_OK;
}
