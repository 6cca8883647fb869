// K2-K5, K8: the MLP tower mini-batch (forward, sigmoid-BCE head, backward) -- fp32 SIMT path (MAMDR_PREC_FP32,
// the parity mode).  The tensor-core modes live in mlp_pass.cu (one persistent kernel per domain pass).
//
// Replaces the Keras train / test functions that the reference drives once per mini-batch
// (/root/reference/model_zoo/mamdr.py:54,85-97; model_zoo/domain_negotiation.py:71-72;
// model_zoo/specific_base_model.py:82-85) for the topology built in
// model_zoo/DeepCTR/deepctr.py:118-136.  Numerical contract: SURVEY.md Appendix A-2, A-3, A-10
// (restated on the CPU in oracle/mlp.py).
//
// Launch sequence of one training mini-batch (all on the caller's stream, graph-capturable):
//   assemble_batch -> L x fwd GEMM(+bias+ReLU+dropout) -> head(sigmoid, BCE, ds, dg, dw, dZ_last, AUC)
//   -> (L-1) x dH GEMM(+mask) -> L x dW GEMM (deterministic split-K) -> colsum(db) -> domain-emb grad
#include "common.cuh"
#include "gemm_simt.cuh"
#include "mlp_ws.cuh"
#include "philox.cuh"

int mamdr_assemble_batch(mamdr_ctx* ctx, const float* Eu, const float* Ei, const float* Ed,
                         const mamdr_batch* b, int du, int di, int dd, float* X, float* y,
                         int32_t* uid_b, int32_t* pid_b, cudaStream_t stream);

namespace {

using namespace mlpws;

// ---- epilogues ---------------------------------------------------------------------------------
struct FwdEpilogue {  // H_out = dropout(relu(acc + bias))
    const float* bias;
    float*       out;
    int          N;
    const OptState* state;
    DropoutParams dp;  // dp.step filled from state on device
    __device__ __forceinline__ void operator()(int m, int n, float4 a) const {
        const float4 b = ldg_f4(bias + n);
        float v[4] = {fmaxf(a.x + b.x, 0.f), fmaxf(a.y + b.y, 0.f), fmaxf(a.z + b.z, 0.f), fmaxf(a.w + b.w, 0.f)};
        if (dp.enabled) {
            DropoutParams q = dp;
            q.step = (uint32_t)(state->step & 0xffffffffll);
            const uint4 w = dropout_words4(q, ((uint32_t)m + q.row0) * (uint32_t)N + (uint32_t)n);   // element index of the GLOBAL batch row
            v[0] = dropout_apply(q, w.x, v[0]);
            v[1] = dropout_apply(q, w.y, v[1]);
            v[2] = dropout_apply(q, w.z, v[2]);
            v[3] = dropout_apply(q, w.w, v[3]);
        }
        *reinterpret_cast<float4*>(out + (int64_t)m * N + n) = make_float4(v[0], v[1], v[2], v[3]);
    }
};

struct DhEpilogue {  // dZ_prev = acc * (H > 0 ? inv_keep : 0)    (H = relu(Z) * M, so H>0 <=> Z>0 and kept)
    const float* H;
    float*       out;
    int          N;
    float        inv_keep;
    __device__ __forceinline__ void operator()(int m, int n, float4 a) const {
        const float4 h = *reinterpret_cast<const float4*>(H + (int64_t)m * N + n);
        float4 r;
        r.x = h.x > 0.f ? a.x * inv_keep : 0.f;
        r.y = h.y > 0.f ? a.y * inv_keep : 0.f;
        r.z = h.z > 0.f ? a.z * inv_keep : 0.f;
        r.w = h.w > 0.f ? a.w * inv_keep : 0.f;
        *reinterpret_cast<float4*>(out + (int64_t)m * N + n) = r;
    }
};

struct StoreEpilogue {  // plain store with leading dimension
    float* out;
    int    ld;
    __device__ __forceinline__ void operator()(int m, int n, float4 a) const {
        *reinterpret_cast<float4*>(out + (int64_t)m * ld + n) = a;
    }
};

// ---- head: z = H_L.w, p = sigmoid(z+g), BCE, ds, dg, dw, dZ_{L-1}, AUC bins -------------------------
// One CTA of 1024 threads = 32 warps, warp per row (coalesced row reads, shuffle reduction);
// all cross-row reductions are fixed-order => deterministic.
constexpr int kHeadThreads = 1024;
// kHeadMaxN, kHeadCtas and the layout of the per-CTA partial records: mlp_ws.cuh

struct HeadArgs {
    const float* HL;      // [b, n]
    const float* w;       // [n]
    const float* g;       // [1]
    const float* y;       // [b]
    const float* Ed;      // domain table [n_domain*dd] for the L2 term
    int          ed_elems;
    int          b, n;
    int          train;
    float        inv_keep, l2_emb, frozen_reg;
    float*       p_out;   // [b] workspace
    float*       probs;   // optional user copy
    float*       ds;      // [b]
    float*       dZ;      // [b, n]   (train)
    float*       g_w;     // grads of dense_kernel [n] (train)
    float*       g_g;     // grad of global_bias [1] (train)
    float*       loss;    // [1]
    float*       auc_acc; // optional [4, T]
    const float* thr;
    int          T;
    float*       part;    // [gridDim.x][kHeadPartStride] per-CTA partials (dw, dg, bce, AUC bins), workspace
    unsigned int* ticket; // last-CTA-done counter (zero between launches)
    int          rows_per_cta;
};

__global__ void __launch_bounds__(kHeadThreads) head_kernel(HeadArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: w[n] | dw_part[32][n] | red_d[32] (double) | red_f[32] | hist[2][T+1] (int) | thr[T] | sz[1024]
    float*  sw      = reinterpret_cast<float*>(smem_raw);
    float*  dw_part = sw + a.n;
    double* red_d   = reinterpret_cast<double*>(dw_part + 32 * a.n + ((32 * a.n + a.n) & 1));
    float*  red_f   = reinterpret_cast<float*>(red_d + 32);
    int*    hist    = reinterpret_cast<int*>(red_f + 32);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T1 = a.T + 1;
    float*  sthr    = reinterpret_cast<float*>(hist + 2 * T1);   // threshold table: the binary search runs in smem
    float*  sz      = sthr + a.T;                                // per-row logit, then per-row ds, of a 1024-row chunk

    for (int c = tid; c < a.n; c += kHeadThreads) sw[c] = a.w[c];
    if (a.auc_acc) {
        for (int i = tid; i < 2 * T1; i += kHeadThreads) hist[i] = 0;
        for (int i = tid; i < a.T; i += kHeadThreads) sthr[i] = a.thr[i];
    }
    __syncthreads();

    const float gbias = a.g[0];
    const float fb = (float)a.b;
    const float lo = 1e-7f, hi = 1.0f - 1e-7f;
    double bce_sum = 0.0;   // per thread (thread t owns rows t, t + 1024, ...)
    float  dg_sum  = 0.f;
    // per-lane partial dw for columns c = lane + 32*j, j < n/32 (n <= kHeadMaxN => <= 16 regs)
    float dwacc[kHeadMaxN / 32];
#pragma unroll
    for (int j = 0; j < kHeadMaxN / 32; ++j) dwacc[j] = 0.f;

    const int row_beg = blockIdx.x * a.rows_per_cta, row_end = min(a.b, row_beg + a.rows_per_cta);
    for (int r0 = row_beg; r0 < row_end; r0 += kHeadThreads) {
        // ---- phase A: logits, one warp per row, four rows in flight (coalesced row reads, shuffle reduction)
        for (int i0 = 0; i0 < 32; i0 += 4) {
            float z[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int r = r0 + warp + 32 * (i0 + q);
                if (r < row_end) {
                    const float* h = a.HL + (int64_t)r * a.n;
#pragma unroll
                    for (int j = 0; j < kHeadMaxN / 32; ++j) {
                        const int c = lane + 32 * j;
                        if (c < a.n) z[q] = fmaf(h[c], sw[c], z[q]);
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) z[q] += __shfl_xor_sync(0xffffffffu, z[q], o);
                if (lane == 0) sz[warp + 32 * (i0 + q)] = z[q];
            }
        }
        __syncthreads();
        // ---- phase B: the scalar chain of a row (sigmoid, BCE, ds, AUC bin) -- one THREAD per row
        {
            const int r = r0 + tid;
            float dsv = 0.f;
            if (r < row_end) {
                const float s = sz[tid] + gbias;
                const float p = 1.0f / (1.0f + expf(-s));
                const float yv = a.y[r];
                if (a.train) dsv = (fabsf(s) <= MAMDR_LOGIT_CLIP) ? __fdiv_rn(__fsub_rn(p, yv), fb) : 0.f;
                const float ph = fminf(fmaxf(p, lo), hi);
                const float lg = logf(ph / (1.0f - ph));
                const float bce = fmaxf(lg, 0.f) - lg * yv + log1pf(expf(-fabsf(lg)));
                bce_sum += (double)bce;
                dg_sum += dsv;
                a.p_out[r] = p;
                if (a.probs) a.probs[r] = p;
                if (a.train) a.ds[r] = dsv;
                if (a.auc_acc) {
                    // k = number of thresholds strictly below p  (pred_is_pos[j] = p > thr[j]  <=>  j < k)
                    int lo_i = 0, hi_i = a.T;
                    while (lo_i < hi_i) {
                        const int mid = (lo_i + hi_i) >> 1;
                        if (sthr[mid] < p) lo_i = mid + 1; else hi_i = mid;
                    }
                    atomicAdd(&hist[(yv != 0.f ? T1 : 0) + lo_i], 1);
                }
            }
            sz[tid] = dsv;
        }
        __syncthreads();
        // ---- phase C: dZ of the last hidden layer and the per-lane partials of dw, one warp per row
        if (a.train) {
#pragma unroll 4
            for (int i = 0; i < 32; ++i) {
                const int r = r0 + warp + 32 * i;
                if (r < row_end) {
                    const float dsv = sz[warp + 32 * i];
                    const float* h = a.HL + (int64_t)r * a.n;
#pragma unroll
                    for (int j = 0; j < kHeadMaxN / 32; ++j) {
                        const int c = lane + 32 * j;
                        if (c < a.n) {
                            const float hv = h[c];
                            dwacc[j] = fmaf(hv, dsv, dwacc[j]);
                            const float dh = __fmul_rn(dsv, sw[c]);
                            a.dZ[(int64_t)r * a.n + c] = hv > 0.f ? __fmul_rn(dh, a.inv_keep) : 0.f;
                        }
                    }
                }
            }
        }
        __syncthreads();
    }
    // warp totals in a fixed (butterfly) order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        bce_sum += __shfl_xor_sync(0xffffffffu, bce_sum, o);
        dg_sum += __shfl_xor_sync(0xffffffffu, dg_sum, o);
    }
    // ---- fixed-order reductions inside the CTA, parked as this CTA's partial record
    if (lane == 0) { red_d[warp] = bce_sum; red_f[warp] = dg_sum; }
    if (a.train) {
#pragma unroll
        for (int j = 0; j < kHeadMaxN / 32; ++j) {
            const int c = lane + 32 * j;
            if (c < a.n) dw_part[warp * a.n + c] = dwacc[j];
        }
    }
    __syncthreads();
    float* mine = a.part + (size_t)blockIdx.x * kHeadPartStride;
    if (a.train) {
        for (int c = tid; c < a.n; c += kHeadThreads) {
            float s = 0.f;
            for (int wv = 0; wv < 32; ++wv) s += dw_part[wv * a.n + c];
            mine[c] = s;
        }
    }
    if (tid == 0) {
        double bs = 0.0;
        float dg = 0.f;
        for (int wv = 0; wv < 32; ++wv) { bs += red_d[wv]; dg += red_f[wv]; }
        mine[kHeadPartDg] = dg;
        *reinterpret_cast<double*>(mine + kHeadPartBce) = bs;
    }
    if (a.auc_acc)
        for (int i = tid; i < 2 * T1; i += kHeadThreads) reinterpret_cast<int*>(mine + kHeadPartHist)[i] = hist[i];
    // ---- the last CTA to arrive combines the records in CTA order
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) is_last = (atomicAdd(a.ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const int G = gridDim.x;
    if (a.train) {
        for (int c = tid; c < a.n; c += kHeadThreads) {
            float s = 0.f;
            for (int g = 0; g < G; ++g) s += __ldcg(a.part + (size_t)g * kHeadPartStride + c);
            a.g_w[c] = s;
        }
    }
    if (a.auc_acc) {
        for (int i = tid; i < 2 * T1; i += kHeadThreads) {
            int v = 0;
            for (int g = 0; g < G; ++g) v += __ldcg(reinterpret_cast<const int*>(a.part + (size_t)g * kHeadPartStride + kHeadPartHist) + i);
            hist[i] = v;
        }
    }
    // L2 penalty of the (always trainable) domain table: sum of squares in double, fixed order
    double sq = 0.0;
    for (int i = tid; i < a.ed_elems; i += kHeadThreads) { const double e = a.Ed[i]; sq += e * e; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    __shared__ double sq_part[32];
    if (lane == 0) sq_part[warp] = sq;
    __syncthreads();
    if (tid == 0) {
        double bs = 0.0, sqs = 0.0;
        float dg = 0.f;
        for (int g = 0; g < G; ++g) {
            bs += __ldcg(reinterpret_cast<const double*>(a.part + (size_t)g * kHeadPartStride + kHeadPartBce));
            dg += __ldcg(a.part + (size_t)g * kHeadPartStride + kHeadPartDg);
        }
        for (int wv = 0; wv < 32; ++wv) sqs += sq_part[wv];
        a.loss[0] = (float)(bs / (double)a.b + (double)a.frozen_reg + (double)a.l2_emb * sqs);
        if (a.train) a.g_g[0] = dg;
        *a.ticket = 0;   // re-arm for the next launch
    }
    // ---- AUC: suffix sums of the two histograms -> tp/fp/fn/tn increments
    if (a.auc_acc) {
        __syncthreads();
        // inclusive suffix scan (Hillis-Steele) over T1 bins, both classes; T1 <= 1024 threads
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
        if (tid < a.T) {
            const int npos = hist[T1], nneg = hist[0];     // suffix sum at bin 0 = totals
            const int tp = hist[T1 + tid + 1], fp = hist[tid + 1];   // bins k > j
            a.auc_acc[0 * a.T + tid] += (float)tp;
            a.auc_acc[1 * a.T + tid] += (float)fp;
            a.auc_acc[2 * a.T + tid] += (float)(npos - tp);
            a.auc_acc[3 * a.T + tid] += (float)(nneg - fp);
        }
    }
}

// rows are split over up to kHeadCtas CTAs (>= 128 rows each); `part_ws` = head_part_bytes() of workspace, zero between launches
inline void launch_head(HeadArgs a, void* part_ws, size_t smem, cudaStream_t st) {
    int G = (a.b + 127) / 128;
    if (G > kHeadCtas) G = kHeadCtas;
    if (G < 1) G = 1;
    a.rows_per_cta = (a.b + G - 1) / G;
    G = (a.b + a.rows_per_cta - 1) / a.rows_per_cta;
    a.part = (float*)part_ws;
    a.ticket = (unsigned int*)((unsigned char*)part_ws + (size_t)kHeadCtas * kHeadPartStride * 4);
    head_kernel<<<G, kHeadThreads, smem, st>>>(a);
}

size_t head_smem_bytes(int n, int T) {
    size_t floats = (size_t)n + 32 * (size_t)n;
    floats += (floats & 1);
    return floats * 4 + 32 * 8 + 32 * 4 + 2 * (size_t)(T + 1) * 4 + (size_t)T * 4 + (size_t)kHeadThreads * 4 + 16;
}

// ---- column sums: db_l[c] = sum_r dZ_l[r, c], fixed order ------------------------------------------
struct ColsumJob { const float* src; float* dst; int n; };
struct ColsumArgs { ColsumJob job[MAMDR_MAX_LAYERS]; int rows; };

// 32 columns x 32 row groups per CTA; every thread keeps 4 independent loads in flight and adds them in a fixed order
constexpr int kColsumThreads = 1024;
__device__ __forceinline__ float colsum_thread(const float* __restrict__ src, int rows, int ld, int c, int ty) {
    float s = 0.f;
    int r = ty;
    for (; r + 96 < rows; r += 128) {
        const float v0 = src[(int64_t)r * ld + c], v1 = src[(int64_t)(r + 32) * ld + c];
        const float v2 = src[(int64_t)(r + 64) * ld + c], v3 = src[(int64_t)(r + 96) * ld + c];
        s += v0; s += v1; s += v2; s += v3;
    }
    for (; r < rows; r += 32) s += src[(int64_t)r * ld + c];
    return s;
}

__global__ void __launch_bounds__(kColsumThreads) colsum_kernel(ColsumArgs a) {
    const ColsumJob j = a.job[blockIdx.y];
    if (blockIdx.x * 32 >= j.n) return;
    const int lx = threadIdx.x & 31, c = blockIdx.x * 32 + lx;
    const int ty = threadIdx.x >> 5;  // 32 row groups
    __shared__ float part[32][33];
    part[ty][lx] = c < j.n ? colsum_thread(j.src, a.rows, j.n, c, ty) : 0.f;
    __syncthreads();
    if (ty == 0 && c < j.n) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k) t += part[k][lx];
        j.dst[c] = t;
    }
}

// ---- domain-embedding gradient:
//   gEd = 2*l2*Ed ; gEd[dom, c] += sum_k db0[k] * W0[du+di+c, k]     (dom uniform per batch)
__global__ void __launch_bounds__(1024)
domain_emb_grad_kernel(const float* __restrict__ Ed, const float* __restrict__ W0dom /* [dd, n1] */,
                       const float* __restrict__ db0, int n_domain, int dd, int n1, int dom,
                       float two_l2, float* __restrict__ gEd) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < n_domain * dd; i += 1024)
        if (i / dd != dom) gEd[i] = two_l2 * Ed[i];
    for (int c = warp; c < dd; c += 32) {
        const float* wr = W0dom + (int64_t)c * n1;
        float s = 0.f;
        for (int k = lane; k < n1; k += 32) s = fmaf(db0[k], wr[k], s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) {
            const int i = dom * dd + c;
            gEd[i] = __fadd_rn(__fmul_rn(two_l2, Ed[i]), s);
        }
    }
}

int validate(mamdr_ctx* ctx, const mamdr_mlp_desc* d, const mamdr_batch* b, const void* ws, size_t ws_bytes,
             int precision_mode, const float* ut, const float* it) {
    MAMDR_REQUIRE(ctx, ctx && d && b, MAMDR_E_INVALID, "NULL ctx/desc/batch");
    MAMDR_REQUIRE(ctx, ctx->prog == nullptr, MAMDR_E_INVALID, "per-mini-batch calls cannot be recorded into a program");
    MAMDR_REQUIRE(ctx, d->n_layers >= 1 && d->n_layers <= MAMDR_MAX_LAYERS, MAMDR_E_INVALID, "n_layers out of range");
    for (int i = 0; i < 3; ++i)
        MAMDR_REQUIRE(ctx, d->emb_dim[i] > 0 && d->emb_dim[i] % 4 == 0, MAMDR_E_INVALID, "emb_dim must be a multiple of 4");
    for (int l = 0; l < d->n_layers; ++l)
        MAMDR_REQUIRE(ctx, d->hidden[l] > 0 && d->hidden[l] % 4 == 0, MAMDR_E_INVALID, "hidden widths must be multiples of 4");
    MAMDR_REQUIRE(ctx, d->hidden[d->n_layers - 1] <= kHeadMaxN, MAMDR_E_UNSUPPORTED, "last hidden layer wider than %d", kHeadMaxN);
    MAMDR_REQUIRE(ctx, d->dropout_rate >= 0.f && d->dropout_rate < 1.f, MAMDR_E_INVALID, "dropout_rate must be in [0,1)");
    MAMDR_REQUIRE(ctx, b->rows >= 1, MAMDR_E_INVALID, "empty batch");
    MAMDR_REQUIRE(ctx, b->domain >= 0 && b->domain < d->n_domain, MAMDR_E_INVALID, "domain id out of range");
    MAMDR_REQUIRE(ctx, b->uid_dev && b->pid_dev && b->label_dev, MAMDR_E_INVALID, "NULL batch column");
    MAMDR_REQUIRE(ctx, ws && aligned16(ws), MAMDR_E_INVALID, "workspace NULL or misaligned");
    MAMDR_REQUIRE(ctx, ws_bytes >= ws_layout(*d, b->rows).total, MAMDR_E_WORKSPACE, "workspace too small: %zu < %zu",
                  ws_bytes, ws_layout(*d, b->rows).total);
    MAMDR_REQUIRE(ctx, precision_mode == MAMDR_PREC_FP32 || precision_mode == MAMDR_PREC_TF32 || precision_mode == MAMDR_PREC_TF32X3,
                  MAMDR_E_INVALID, "unknown precision_mode %d", precision_mode);
    MAMDR_REQUIRE(ctx, precision_mode == MAMDR_PREC_FP32, MAMDR_E_UNSUPPORTED,
                  "the per-mini-batch entry points run the fp32 SIMT tower; the tcgen05 modes are served by "
                  "mamdr_mlp_train_pass / mamdr_mlp_eval_pass");
    if (!d->emb_trainable) MAMDR_REQUIRE(ctx, ut && it, MAMDR_E_INVALID, "frozen tables are NULL");
    if (d->emb_trainable) MAMDR_REQUIRE(ctx, b->rows <= mamdr_scatter_max_n(), MAMDR_E_UNSUPPORTED, "batch too large for the sparse-gradient dedup");
    return MAMDR_OK;
}

int run_forward(mamdr_ctx* ctx, const mamdr_mlp_desc* d, const mamdr_batch* b, const float* ut, const float* it,
                const float* params, unsigned char* ws, const WsLayout& w, const OptState* state, bool train,
                cudaStream_t st) {
    const int du = d->emb_dim[0], di = d->emb_dim[1], dd = d->emb_dim[2];
    const float* Eu = d->emb_trainable ? params + d->off_user_emb : ut;
    const float* Ei = d->emb_trainable ? params + d->off_item_emb : it;
    const float* Ed = params + d->off_domain_emb;
    int rc = mamdr_assemble_batch(ctx, Eu, Ei, Ed, b, du, di, dd, (float*)(ws + w.H[0]), (float*)(ws + w.y),
                                  (int32_t*)(ws + w.uid_b), (int32_t*)(ws + w.pid_b), st);
    if (rc) return rc;
    int K = du + di + dd;
    const float keep = 1.0f - d->dropout_rate;
    for (int l = 0; l < d->n_layers; ++l) {
        const int N = d->hidden[l];
        FwdEpilogue epi;
        epi.bias = params + d->off_bias[l];
        epi.out = (float*)(ws + w.H[l + 1]);
        epi.N = N;
        epi.state = state;
        epi.dp.enabled = (train && d->dropout_rate > 0.f) ? 1 : 0;
        epi.dp.seed = d->dropout_seed + (uint32_t)l;
        epi.dp.step = 0;
        epi.dp.row0 = (uint32_t)b->row0;
        double thr = floor((double)keep * 4294967296.0);
        epi.dp.threshold = thr >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)thr;
        epi.dp.scale = 1.0f / keep;
        simt::GemmShape s{b->rows, N, K, K, N};
        simt::LaunchPlan p = simt::plan(b->rows, N, K, 0, 1);
        simt::gemm_kernel<true, true, FwdEpilogue><<<p.grid, simt::THREADS, 0, st>>>(
            (const float*)(ws + w.H[l]), params + d->off_kernel[l], s, p.k_chunk, nullptr, nullptr, epi);
        MAMDR_LAUNCH_OK(ctx);
        K = N;
    }
    return MAMDR_OK;
}

int run_head(mamdr_ctx* ctx, const mamdr_mlp_desc* d, const mamdr_batch* b, const float* params, float* grads,
             unsigned char* ws, const WsLayout& w, bool train, float* loss, float* probs, float* auc_acc,
             const float* thr, int T, cudaStream_t st) {
    const int L = d->n_layers, n = d->hidden[L - 1];
    MAMDR_REQUIRE(ctx, loss != nullptr, MAMDR_E_INVALID, "loss_dev is NULL");
    if (auc_acc) MAMDR_REQUIRE(ctx, thr && T >= 2 && T + 1 <= kHeadThreads, MAMDR_E_INVALID, "bad AUC thresholds (2 <= T <= 1023)");
    HeadArgs a;
    a.HL = (const float*)(ws + w.H[L]);
    a.w = params + d->off_dense_kernel;
    a.g = params + d->off_global_bias;
    a.y = (const float*)(ws + w.y);
    a.Ed = params + d->off_domain_emb;
    a.ed_elems = d->n_domain * d->emb_dim[2];
    a.b = b->rows;
    a.n = n;
    a.train = train ? 1 : 0;
    a.inv_keep = (train && d->dropout_rate > 0.f) ? 1.0f / (1.0f - d->dropout_rate) : 1.0f;
    a.l2_emb = d->l2_emb;
    a.frozen_reg = d->frozen_reg;
    a.p_out = (float*)(ws + w.p);
    a.probs = probs;
    a.ds = (float*)(ws + w.ds);
    a.dZ = (float*)(ws + w.dZ[L - 1]);
    a.g_w = train ? grads + d->off_dense_kernel : nullptr;
    a.g_g = train ? grads + d->off_global_bias : nullptr;
    a.loss = loss;
    a.auc_acc = auc_acc;
    a.thr = thr;
    a.T = auc_acc ? T : 0;
    const size_t smem = head_smem_bytes(n, a.T);
    MAMDR_REQUIRE(ctx, smem <= 100 * 1024, MAMDR_E_UNSUPPORTED, "head smem %zu too large", smem);
    launch_head(a, ws + w.hist, smem, st);
    MAMDR_LAUNCH_OK(ctx);
    return MAMDR_OK;
}


}  // namespace

#include "star.cuh"
#include "mtl.cuh"

int mamdr_mlp_init_kernels(mamdr_ctx* ctx) {
    MAMDR_CUDA_OK(ctx, cudaFuncSetAttribute(head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    return MAMDR_OK;
}

extern "C" size_t mamdr_mlp_workspace_bytes(const mamdr_mlp_desc* desc, int32_t max_batch) {
    if (!desc || max_batch < 1 || desc->n_layers < 1 || desc->n_layers > MAMDR_MAX_LAYERS) return 0;
    return ws_layout(*desc, max_batch).total;
}

extern "C" int mamdr_mlp_sparse_grads(const mamdr_mlp_desc* desc, int32_t rows, void* ws_dev, int32_t table,
                                      const int32_t** uniq_ids_dev, const float** uniq_rows_dev, const int32_t** n_uniq_dev) {
    if (!desc || !ws_dev || !desc->emb_trainable || table < 0 || table > 1 || rows < 1) return MAMDR_E_INVALID;
    const WsLayout w = ws_layout(*desc, rows);
    unsigned char* ws = (unsigned char*)ws_dev;
    if (uniq_ids_dev) *uniq_ids_dev = (const int32_t*)(ws + w.sp_ids[table]);
    if (uniq_rows_dev) *uniq_rows_dev = (const float*)(ws + w.sp_rows[table]);
    if (n_uniq_dev) *n_uniq_dev = (const int32_t*)(ws + w.sp_n[table]);
    return MAMDR_OK;
}

// dX[:, 0:du+di] = dZ_0 . W_0[0:du+di, :]^T from the workspace of the LAST train step of `rows` rows: the gradient rows of
// the gathered user / item embeddings.  Used when the tables live outside the arena (row-sharded across GPUs).
extern "C" int mamdr_mlp_input_grads(mamdr_ctx* ctx, const mamdr_mlp_desc* d, int32_t rows, const float* params, void* ws_,
                                     size_t ws_bytes, float* dX_out, mamdr_stream stream) {
    MAMDR_REQUIRE(ctx, ctx && d && params && ws_ && dX_out, MAMDR_E_INVALID, "NULL pointer");
    MAMDR_REQUIRE(ctx, rows >= 1 && aligned16(dX_out) && aligned16(ws_) && aligned16(params), MAMDR_E_INVALID, "bad rows / misaligned pointer");
    MAMDR_REQUIRE(ctx, ctx->prog == nullptr, MAMDR_E_INVALID, "per-mini-batch calls cannot be recorded into a program");
    const WsLayout w = ws_layout(*d, rows);
    MAMDR_REQUIRE(ctx, ws_bytes >= w.total, MAMDR_E_WORKSPACE, "workspace too small");
    unsigned char* ws = (unsigned char*)ws_;
    const int dui = d->emb_dim[0] + d->emb_dim[1], n1 = d->hidden[0];
    StoreEpilogue epi{dX_out, dui};
    simt::GemmShape s{rows, dui, n1, n1, n1};
    simt::LaunchPlan p = simt::plan(rows, dui, n1, 0, 1);
    simt::gemm_kernel<true, false, StoreEpilogue><<<p.grid, simt::THREADS, 0, (cudaStream_t)stream>>>(
        (const float*)(ws + w.dZ[0]), params + d->off_kernel[0], s, p.k_chunk, nullptr, nullptr, epi);
    MAMDR_LAUNCH_OK(ctx);
    return MAMDR_OK;
}

extern "C" int mamdr_mlp_eval_step(mamdr_ctx* ctx, const mamdr_mlp_desc* d, const mamdr_batch* b,
                                   const float* ut, const float* it, const float* params, void* ws_,
                                   size_t ws_bytes, float* loss, float* probs, float* auc_acc,
                                   const float* thr, int32_t T, int32_t precision_mode, mamdr_stream stream) {
    int rc = validate(ctx, d, b, ws_, ws_bytes, precision_mode, ut, it);
    if (rc) return rc;
    MAMDR_REQUIRE(ctx, params && aligned16(params), MAMDR_E_INVALID, "params NULL or misaligned");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* ws = (unsigned char*)ws_;
    const WsLayout w = ws_layout(*d, b->rows);
    rc = run_forward(ctx, d, b, ut, it, params, ws, w, nullptr, false, st);
    if (rc) return rc;
    return run_head(ctx, d, b, params, nullptr, ws, w, false, loss, probs, auc_acc, thr, T, st);
}

extern "C" int mamdr_mlp_train_step(mamdr_ctx* ctx, const mamdr_mlp_desc* d, const mamdr_batch* b,
                                    const float* ut, const float* it, const float* params, float* grads,
                                    void* ws_, size_t ws_bytes, const void* opt_state, float* loss,
                                    float* probs, float* auc_acc, const float* thr, int32_t T,
                                    int32_t precision_mode, mamdr_stream stream) {
    int rc = validate(ctx, d, b, ws_, ws_bytes, precision_mode, ut, it);
    if (rc) return rc;
    MAMDR_REQUIRE(ctx, params && grads && aligned16(params) && aligned16(grads), MAMDR_E_INVALID,
                  "params/grads NULL or misaligned");
    MAMDR_REQUIRE(ctx, opt_state != nullptr, MAMDR_E_INVALID, "opt_state is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* ws = (unsigned char*)ws_;
    const WsLayout w = ws_layout(*d, b->rows);
    const int L = d->n_layers, rows = b->rows;
    const int in_dim = d->emb_dim[0] + d->emb_dim[1] + d->emb_dim[2];
    const float inv_keep = d->dropout_rate > 0.f ? 1.0f / (1.0f - d->dropout_rate) : 1.0f;

    MAMDR_CUDA_OK(ctx, cudaMemsetAsync(ws + w.tickets, 0, (size_t)kMaxTiles * 4, st));
    rc = run_forward(ctx, d, b, ut, it, params, ws, w, (const OptState*)opt_state, true, st);
    if (rc) return rc;
    rc = run_head(ctx, d, b, params, grads, ws, w, true, loss, probs, auc_acc, thr, T, st);
    if (rc) return rc;

    // ---- dZ_{l-1} = (dZ_l . W_l^T) * mask(H_l)   for l = L-1 .. 1
    for (int l = L - 1; l >= 1; --l) {
        const int Kd = d->hidden[l], Nd = d->hidden[l - 1];
        DhEpilogue epi{(const float*)(ws + w.H[l]), (float*)(ws + w.dZ[l - 1]), Nd, inv_keep};
        simt::GemmShape s{rows, Nd, Kd, Kd, Kd};  // A = dZ_l [rows,Kd]; B = W_l stored [Nd,Kd] (k contiguous)
        simt::LaunchPlan p = simt::plan(rows, Nd, Kd, 0, 1);
        simt::gemm_kernel<true, false, DhEpilogue><<<p.grid, simt::THREADS, 0, st>>>(
            (const float*)(ws + w.dZ[l]), params + d->off_kernel[l], s, p.k_chunk, nullptr, nullptr, epi);
        MAMDR_LAUNCH_OK(ctx);
    }
    // ---- trainable tables: dX[:, 0:du+di] = dZ_0 . W_0[0:du+di, :]^T, then per table sort + segment-sum (K6)
    if (d->emb_trainable) {
        const int dui = d->emb_dim[0] + d->emb_dim[1], n1 = d->hidden[0];
        StoreEpilogue epi{(float*)(ws + w.dX), dui};
        simt::GemmShape s{rows, dui, n1, n1, n1};   // A = dZ_0 [rows, n1]; B = W_0 rows [0, dui) stored [dui, n1]
        simt::LaunchPlan p = simt::plan(rows, dui, n1, 0, 1);
        simt::gemm_kernel<true, false, StoreEpilogue><<<p.grid, simt::THREADS, 0, st>>>(
            (const float*)(ws + w.dZ[0]), params + d->off_kernel[0], s, p.k_chunk, nullptr, nullptr, epi);
        MAMDR_LAUNCH_OK(ctx);
        DedupJob jobs[2];   // both tables: one sort launch + one segment-sum launch
        for (int t = 0; t < 2; ++t) {
            jobs[t] = DedupJob{(const int32_t*)(ws + (t == 0 ? w.uid_b : w.pid_b)), (const float*)(ws + w.dX) + (t == 0 ? 0 : d->emb_dim[0]), dui,
                               d->emb_dim[t], (int32_t*)(ws + w.sp_ids[t]), (float*)(ws + w.sp_rows[t]), (int32_t*)(ws + w.sp_n[t]), nullptr, nullptr};
            mamdr_scatter_job_ws(&jobs[t], ws + w.sp_ws + t * mamdr_scatter_workspace_bytes(rows), rows);
        }
        rc = mamdr_scatter_dedup_jobs(ctx, jobs, 2, rows, st);
        if (rc) return rc;
    }
    // ---- dW_l = H_l^T . dZ_l   (reduction over the batch rows; deterministic split-K)
    for (int l = 0; l < L; ++l) {
        const int Md = l == 0 ? in_dim : d->hidden[l - 1], Nd = d->hidden[l];
        StoreEpilogue epi{grads + d->off_kernel[l], Nd};
        simt::GemmShape s{Md, Nd, rows, Md, Nd};  // A = H_l stored [rows,Md] (m contiguous); B = dZ_l [rows,Nd]
        simt::LaunchPlan p = simt::plan(Md, Nd, rows, ctx->sm_count, kMaxSplit);
        MAMDR_REQUIRE(ctx, (int)(p.grid.x * p.grid.y) <= kMaxTiles, MAMDR_E_UNSUPPORTED, "layer too large for the ticket table");
        simt::gemm_kernel<false, true, StoreEpilogue><<<p.grid, simt::THREADS, 0, st>>>(
            (const float*)(ws + w.H[l]), (const float*)(ws + w.dZ[l]), s, p.k_chunk, (float*)(ws + w.partials),
            (unsigned int*)(ws + w.tickets), epi);
        MAMDR_LAUNCH_OK(ctx);
    }
    // ---- db_l
    {
        ColsumArgs ca;
        int maxn = 0;
        for (int l = 0; l < L; ++l) {
            ca.job[l] = ColsumJob{(const float*)(ws + w.dZ[l]), grads + d->off_bias[l], d->hidden[l]};
            if (d->hidden[l] > maxn) maxn = d->hidden[l];
        }
        ca.rows = rows;
        colsum_kernel<<<dim3((maxn + 31) / 32, L), kColsumThreads, 0, st>>>(ca);
        MAMDR_LAUNCH_OK(ctx);
    }
    // ---- domain embedding gradient (uses db_0 just written)
    {
        const int dd = d->emb_dim[2], n1 = d->hidden[0];
        const float* W0dom = params + d->off_kernel[0] + (int64_t)(d->emb_dim[0] + d->emb_dim[1]) * n1;
        domain_emb_grad_kernel<<<1, 1024, 0, st>>>(params + d->off_domain_emb, W0dom, grads + d->off_bias[0],
                                                   d->n_domain, dd, n1, b->domain, 2.0f * d->l2_emb,
                                                   grads + d->off_domain_emb);
        MAMDR_LAUNCH_OK(ctx);
    }
    return MAMDR_OK;
}
