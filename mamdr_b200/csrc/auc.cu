// K8: streaming 500-threshold ROC-AUC as a threshold-bin histogram + suffix sums.
//
// Replaces AUC.update_state / result (/root/reference/utils/auc.py:159-177,248-281) and
// update_confusion_matrix_variables (utils/metrics_utils.py:297-354), which materialise four
// [T, b] boolean tiles per batch.  Here each prediction is binned once by binary search on the
// fp32 threshold table (bit-exact with the strict fp32 `pred > threshold` compare), integer
// histograms are suffix-summed, and the four fp32 [T] accumulators get one add each.
#include "common.cuh"

namespace {

constexpr int kThreads = 1024;

__global__ void __launch_bounds__(kThreads)
auc_update_kernel(const float* __restrict__ p, const float* __restrict__ y, int64_t n, float* __restrict__ acc,
                  const float* __restrict__ thr, int T) {
    extern __shared__ int hist[];  // [2][T+1]
    const int T1 = T + 1, tid = threadIdx.x;
    for (int i = tid; i < 2 * T1; i += kThreads) hist[i] = 0;
    __syncthreads();
    for (int64_t i = tid; i < n; i += kThreads) {
        const float pv = p[i];
        int lo = 0, hi = T;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(thr + mid) < pv) lo = mid + 1; else hi = mid;
        }
        atomicAdd(&hist[(y[i] != 0.f ? T1 : 0) + lo], 1);
    }
    __syncthreads();
    int vneg = tid < T1 ? hist[tid] : 0;
    int vpos = tid < T1 ? hist[T1 + tid] : 0;
    for (int off = 1; off < T1; off <<= 1) {
        __syncthreads();
        const int aneg = (tid + off < T1) ? hist[tid + off] : 0;
        const int apos = (tid + off < T1) ? hist[T1 + tid + off] : 0;
        __syncthreads();
        vneg += aneg; vpos += apos;
        if (tid < T1) { hist[tid] = vneg; hist[T1 + tid] = vpos; }
    }
    __syncthreads();
    if (tid < T) {
        const int npos = hist[T1], nneg = hist[0];
        const int tp = hist[T1 + tid + 1], fp = hist[tid + 1];
        acc[0 * T + tid] += (float)tp;
        acc[1 * T + tid] += (float)fp;
        acc[2 * T + tid] += (float)(npos - tp);
        acc[3 * T + tid] += (float)(nneg - fp);
    }
}

__device__ __forceinline__ float div_no_nan(float a, float b) { return b != 0.f ? a / b : 0.f; }

__global__ void __launch_bounds__(kThreads) auc_result_kernel(const float* __restrict__ acc, int T, float* out) {
    __shared__ float recall[kThreads], fpr[kThreads], term[kThreads];
    const int tid = threadIdx.x;
    if (tid < T) {
        const float tp = acc[tid], fp = acc[T + tid], fn = acc[2 * T + tid], tn = acc[3 * T + tid];
        recall[tid] = div_no_nan(tp, tp + fn);
        fpr[tid] = div_no_nan(fp, fp + tn);
    }
    __syncthreads();
    term[tid] = (tid < T - 1) ? __fmul_rn(__fsub_rn(fpr[tid], fpr[tid + 1]), __fdiv_rn(__fadd_rn(recall[tid], recall[tid + 1]), 2.0f))
                              : 0.f;
    __syncthreads();
    for (int s = kThreads / 2; s > 0; s >>= 1) {
        if (tid < s) term[tid] += term[tid + s];
        __syncthreads();
    }
    if (tid == 0) out[0] = term[0];
}

}  // namespace

extern "C" int mamdr_auc_update(mamdr_ctx* ctx, const float* probs, const float* labels, int64_t n, float* acc,
                                const float* thr, int32_t T, mamdr_stream stream) {
    MAMDR_REQUIRE(ctx, ctx != nullptr, MAMDR_E_INVALID, "ctx is NULL");
    MAMDR_REQUIRE(ctx, n >= 0, MAMDR_E_INVALID, "negative n");
    MAMDR_REQUIRE(ctx, acc && thr && T >= 2 && T + 1 <= kThreads, MAMDR_E_INVALID, "bad accumulators/thresholds (2 <= T <= 1023)");
    if (n == 0) return MAMDR_OK;
    MAMDR_REQUIRE(ctx, probs && labels, MAMDR_E_INVALID, "NULL probs/labels");
    auc_update_kernel<<<1, kThreads, 2 * (T + 1) * sizeof(int), (cudaStream_t)stream>>>(probs, labels, n, acc, thr, T);
    MAMDR_LAUNCH_OK(ctx);
    return MAMDR_OK;
}

extern "C" int mamdr_auc_result(mamdr_ctx* ctx, const float* acc, int32_t T, float* auc, mamdr_stream stream) {
    MAMDR_REQUIRE(ctx, ctx && acc && auc, MAMDR_E_INVALID, "NULL pointer");
    MAMDR_REQUIRE(ctx, T >= 2 && T <= kThreads, MAMDR_E_INVALID, "2 <= T <= 1024");
    auc_result_kernel<<<1, kThreads, 0, (cudaStream_t)stream>>>(acc, T, auc);
    MAMDR_LAUNCH_OK(ctx);
    return MAMDR_OK;
}
