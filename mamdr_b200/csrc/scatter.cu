// K6: sparse embedding-gradient de-duplication: sort-by-id + warp-per-segment sum.
//
// Replaces TF's _deduplicate_indexed_slices (tf.unique + unsorted_segment_sum) that runs inside
// AdamOptimizer.apply_gradients for the IndexedSlices gradient of tf.gather
// (/root/reference/model_zoo/DeepCTR/deepctr.py:54-55,125-126).
//
//   pass 1 (one CTA): 64-bit keys (id << 32 | batch position) -> in-smem bitonic sort (unique keys
//           => the order is the stable order), head flags, exclusive scan -> sorted unique ids,
//           segment starts, permutation.
//   pass 2 (warp per unique id): rows of one id are added sequentially in batch order, each lane
//           owning 4-float column groups -> deterministic, bit-identical to numpy add.at.
// No atomics.  n <= 8192 per call (one mini-batch).
#include "common.cuh"

namespace {

constexpr int kSortThreads = 1024;
constexpr int kMaxN = 8192;

__global__ void __launch_bounds__(kSortThreads) sort_unique_kernel(const __grid_constant__ DedupArgs a) {
    extern __shared__ __align__(16) unsigned long long keys[];  // [npow2]
    __shared__ int scan_part[kSortThreads];
    const DedupJob& J = a.job[blockIdx.x];
    const int32_t* __restrict__ ids = J.ids;
    int32_t* __restrict__ uniq_ids = J.uniq_ids;
    int32_t* __restrict__ perm = J.perm;
    int32_t* __restrict__ seg_start = J.seg_start;
    const int n = a.n, npow2 = a.npow2;
    const int tid = threadIdx.x;
    __shared__ int n_valid;
    if (tid == 0) n_valid = 0;
    // negative ids are padding (fixed-capacity all-to-all buffers of the row-sharded tables): they sort behind every real id
    for (int i = tid; i < npow2; i += kSortThreads)
        keys[i] = (i < n && ids[i] >= 0) ? (((unsigned long long)(uint32_t)ids[i]) << 32) | (uint32_t)i : ~0ull;
    __syncthreads();
    for (int k = 2; k <= npow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < npow2; i += kSortThreads) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long x = keys[i], y = keys[ixj];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { keys[i] = y; keys[ixj] = x; }
                }
            }
            __syncthreads();
        }
    }
    // head flags + exclusive scan (each thread owns a contiguous chunk)
    const int per = (n + kSortThreads - 1) / kSortThreads;
    const int beg = tid * per, end = min(n, beg + per);
    int cnt = 0;
    for (int i = beg; i < end; ++i) {
        if (keys[i] == ~0ull) break;
        const uint32_t id = (uint32_t)(keys[i] >> 32);
        const bool head = (i == 0) || ((uint32_t)(keys[i - 1] >> 32) != id);
        cnt += head ? 1 : 0;
        if (i + 1 == n || keys[i + 1] == ~0ull) n_valid = i + 1;   // exactly one thread sees the last real entry
    }
    scan_part[tid] = cnt;
    __syncthreads();
    // inclusive Hillis-Steele over 1024 partials
    for (int off = 1; off < kSortThreads; off <<= 1) {
        const int add = tid >= off ? scan_part[tid - off] : 0;
        __syncthreads();
        scan_part[tid] += add;
        __syncthreads();
    }
    int seg = scan_part[tid] - cnt;  // exclusive prefix
    for (int i = beg; i < end; ++i) {
        if (keys[i] == ~0ull) break;
        const uint32_t id = (uint32_t)(keys[i] >> 32);
        const bool head = (i == 0) || ((uint32_t)(keys[i - 1] >> 32) != id);
        if (head) {
            uniq_ids[seg] = (int32_t)id;
            seg_start[seg] = i;
            ++seg;
        }
        perm[i] = (int32_t)(keys[i] & 0xffffffffull);
    }
    if (tid == kSortThreads - 1) {
        const int total = scan_part[tid];
        J.n_uniq[0] = total;
        seg_start[total] = n_valid;
    }
}

// warp per unique id; the rows of one id are added sequentially in batch order (bit-identical to numpy add.at).  The
// permutation entries of a segment are fetched 32 at a time and the row loads run 4 deep ahead of the ordered adds.
__global__ void __launch_bounds__(256) segment_sum_kernel(const __grid_constant__ DedupArgs a) {
    const DedupJob& J = a.job[blockIdx.y];
    const float* __restrict__ grad_rows = J.grad_rows;
    const int64_t grad_stride = J.grad_stride;
    const int dim = J.dim;
    const int32_t* __restrict__ perm = J.perm;
    const int32_t* __restrict__ seg_start = J.seg_start;
    float* __restrict__ uniq_rows = J.uniq_rows;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int nu = J.n_uniq[0];
    for (int k = warp; k < nu; k += nwarps) {
        const int s = seg_start[k], e = seg_start[k + 1];
        for (int base = s; base < e; base += 32) {
            const int cnt = min(32, e - base);
            const int myp = lane < cnt ? perm[base + lane] : 0;
            for (int cc = 0; cc < dim; cc += 128) {
                const int c = cc + lane * 4;
                const bool active = c < dim;
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                if (base != s && active) acc = *reinterpret_cast<const float4*>(uniq_rows + (int64_t)k * dim + c);
                for (int i = 0; i < cnt; i += 4) {
                    int p[4];
                    float4 v[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) p[q] = __shfl_sync(0xffffffffu, myp, min(i + q, cnt - 1));
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        v[q] = (active && i + q < cnt) ? *reinterpret_cast<const float4*>(grad_rows + (int64_t)p[q] * grad_stride + c)
                                                       : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (i + q >= cnt) break;
                        if (base == s && i + q == 0) {
                            acc = v[q];
                        } else {
                            acc.x = __fadd_rn(acc.x, v[q].x); acc.y = __fadd_rn(acc.y, v[q].y);
                            acc.z = __fadd_rn(acc.z, v[q].z); acc.w = __fadd_rn(acc.w, v[q].w);
                        }
                    }
                }
                if (active) *reinterpret_cast<float4*>(uniq_rows + (int64_t)k * dim + c) = acc;
            }
        }
    }
}

inline size_t al(size_t x) { return (x + 255) / 256 * 256; }

}  // namespace

int mamdr_scatter_init_kernels(mamdr_ctx* ctx) {
    MAMDR_CUDA_OK(ctx, cudaFuncSetAttribute(sort_unique_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            kMaxN * (int)sizeof(unsigned long long)));
    return MAMDR_OK;
}

extern "C" int64_t mamdr_scatter_max_n(void) { return kMaxN; }

extern "C" size_t mamdr_scatter_workspace_bytes(int64_t n) {
    if (n < 0) return 0;
    return al((size_t)n * 4) + al((size_t)(n + 1) * 4);
}

// de-duplicate `n_jobs` (<= 2: the user and the item table of one mini-batch) id lists of n entries in two launches
int mamdr_scatter_dedup_jobs(mamdr_ctx* ctx, const DedupJob* jobs, int n_jobs, int n, cudaStream_t st) {
    MAMDR_REQUIRE(ctx, ctx != nullptr && jobs != nullptr, MAMDR_E_INVALID, "ctx / jobs is NULL");
    MAMDR_REQUIRE(ctx, n_jobs >= 1 && n_jobs <= 2, MAMDR_E_INVALID, "n_jobs must be 1 or 2");
    MAMDR_REQUIRE(ctx, n >= 1 && n <= kMaxN, MAMDR_E_UNSUPPORTED, "n=%d outside 1..%d", n, kMaxN);
    DedupArgs a;
    memset(&a, 0, sizeof(a));
    for (int q = 0; q < n_jobs; ++q) {
        const DedupJob& J = jobs[q];
        MAMDR_REQUIRE(ctx, J.ids && J.grad_rows && J.uniq_ids && J.uniq_rows && J.n_uniq && J.perm && J.seg_start, MAMDR_E_INVALID, "NULL pointer");
        MAMDR_REQUIRE(ctx, J.dim > 0 && J.dim % 4 == 0 && J.grad_stride >= J.dim && J.grad_stride % 4 == 0, MAMDR_E_INVALID, "bad dim/stride");
        MAMDR_REQUIRE(ctx, aligned16(J.grad_rows) && aligned16(J.uniq_rows), MAMDR_E_INVALID, "misaligned pointer");
        a.job[q] = J;
    }
    a.n = n;
    a.npow2 = 1;
    while (a.npow2 < n) a.npow2 <<= 1;
    sort_unique_kernel<<<n_jobs, kSortThreads, (size_t)a.npow2 * sizeof(unsigned long long), st>>>(a);
    MAMDR_LAUNCH_OK(ctx);
    segment_sum_kernel<<<dim3((n + 7) / 8, n_jobs), 256, 0, st>>>(a);   // n bounds the number of unique ids
    MAMDR_LAUNCH_OK(ctx);
    return MAMDR_OK;
}

void mamdr_scatter_job_ws(DedupJob* job, void* ws, int64_t n) {
    job->perm = (int32_t*)ws;
    job->seg_start = (int32_t*)((unsigned char*)ws + al((size_t)n * 4));
}

extern "C" int mamdr_scatter_dedup_f32(mamdr_ctx* ctx, const int32_t* ids, const float* grad_rows, int64_t grad_stride,
                                       int64_t n, int32_t dim, int32_t* uniq_ids, float* uniq_rows, int32_t* n_uniq,
                                       void* ws, size_t ws_bytes, mamdr_stream stream) {
    MAMDR_REQUIRE(ctx, ctx != nullptr, MAMDR_E_INVALID, "ctx is NULL");
    MAMDR_REQUIRE(ctx, n >= 0 && n <= kMaxN, MAMDR_E_UNSUPPORTED, "n=%lld exceeds %d", (long long)n, kMaxN);
    MAMDR_REQUIRE(ctx, n_uniq != nullptr, MAMDR_E_INVALID, "n_uniq_dev is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        MAMDR_CUDA_OK(ctx, cudaMemsetAsync(n_uniq, 0, 4, st));
        return MAMDR_OK;
    }
    MAMDR_REQUIRE(ctx, ws && aligned16(ws), MAMDR_E_INVALID, "workspace NULL or misaligned");
    MAMDR_REQUIRE(ctx, ws_bytes >= mamdr_scatter_workspace_bytes(n), MAMDR_E_WORKSPACE, "workspace too small");
    DedupJob J{ids, grad_rows, grad_stride, dim, uniq_ids, uniq_rows, n_uniq, nullptr, nullptr};
    mamdr_scatter_job_ws(&J, ws, n);
    return mamdr_scatter_dedup_jobs(ctx, &J, 1, (int)n, st);
}
