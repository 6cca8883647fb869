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

struct ScatterWs {
    int32_t* perm;       // [n]   batch positions in sorted order
    int32_t* seg_start;  // [n+1] first sorted index of each unique id
};

__global__ void __launch_bounds__(kSortThreads)
sort_unique_kernel(const int32_t* __restrict__ ids, int n, int npow2, int32_t* __restrict__ uniq_ids,
                   int32_t* __restrict__ perm, int32_t* __restrict__ seg_start, int32_t* __restrict__ n_uniq) {
    extern __shared__ __align__(16) unsigned long long keys[];  // [npow2]
    __shared__ int scan_part[kSortThreads];
    const int tid = threadIdx.x;
    for (int i = tid; i < npow2; i += kSortThreads)
        keys[i] = i < n ? (((unsigned long long)(uint32_t)ids[i]) << 32) | (uint32_t)i : ~0ull;
    __syncthreads();
    for (int k = 2; k <= npow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < npow2; i += kSortThreads) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = keys[i], b = keys[ixj];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
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
        const uint32_t id = (uint32_t)(keys[i] >> 32);
        const bool head = (i == 0) || ((uint32_t)(keys[i - 1] >> 32) != id);
        cnt += head ? 1 : 0;
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
        n_uniq[0] = total;
        seg_start[total] = n;
    }
}

__global__ void __launch_bounds__(256)
segment_sum_kernel(const float* __restrict__ grad_rows, int64_t grad_stride, int n, int dim,
                   const int32_t* __restrict__ perm, const int32_t* __restrict__ seg_start,
                   const int32_t* __restrict__ n_uniq, float* __restrict__ uniq_rows) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int nu = n_uniq[0];
    for (int k = warp; k < nu; k += nwarps) {
        const int s = seg_start[k], e = seg_start[k + 1];
        for (int c = lane * 4; c < dim; c += 128) {
            float4 acc = *reinterpret_cast<const float4*>(grad_rows + (int64_t)perm[s] * grad_stride + c);
            for (int i = s + 1; i < e; ++i) {
                const float4 v = *reinterpret_cast<const float4*>(grad_rows + (int64_t)perm[i] * grad_stride + c);
                acc.x = __fadd_rn(acc.x, v.x); acc.y = __fadd_rn(acc.y, v.y);
                acc.z = __fadd_rn(acc.z, v.z); acc.w = __fadd_rn(acc.w, v.w);
            }
            *reinterpret_cast<float4*>(uniq_rows + (int64_t)k * dim + c) = acc;
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
    MAMDR_REQUIRE(ctx, ids && grad_rows && uniq_ids && uniq_rows && ws, MAMDR_E_INVALID, "NULL pointer");
    MAMDR_REQUIRE(ctx, dim > 0 && dim % 4 == 0 && grad_stride >= dim && grad_stride % 4 == 0, MAMDR_E_INVALID, "bad dim/stride");
    MAMDR_REQUIRE(ctx, aligned16(grad_rows) && aligned16(uniq_rows) && aligned16(ws), MAMDR_E_INVALID, "misaligned pointer");
    MAMDR_REQUIRE(ctx, ws_bytes >= mamdr_scatter_workspace_bytes(n), MAMDR_E_WORKSPACE, "workspace too small");
    int32_t* perm = (int32_t*)ws;
    int32_t* seg_start = (int32_t*)((unsigned char*)ws + al((size_t)n * 4));
    int npow2 = 1;
    while (npow2 < n) npow2 <<= 1;
    sort_unique_kernel<<<1, kSortThreads, (size_t)npow2 * sizeof(unsigned long long), st>>>(ids, (int)n, npow2, uniq_ids, perm,
                                                                                          seg_start, n_uniq);
    MAMDR_LAUNCH_OK(ctx);
    const int warps = (int)n;  // upper bound on the number of unique ids
    const int grid = (warps + 7) / 8;
    segment_sum_kernel<<<grid, 256, 0, st>>>(grad_rows, grad_stride, (int)n, dim, perm, seg_start, n_uniq, uniq_rows);
    MAMDR_LAUNCH_OK(ctx);
    return MAMDR_OK;
}
