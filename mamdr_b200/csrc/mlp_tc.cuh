// tcgen05 path of the MLP tower mini-batch (MAMDR_PREC_TF32 / MAMDR_PREC_TF32X3); included by mlp.cu.
//
// Launch sequence of one training mini-batch (11 launches, all graph-capturable):
//   assemble(X, X_lo, y) | split(params -> params_lo) | fwd L0 | fwd L1 | fwd L2 + head (sigmoid, BCE, ds, dZ_2,
//   per-CTA partials of dw / db_2 / loss / dg, AUC bins) | dH_2 -> dZ_1 (+db_1 partials) | dH_1 -> dZ_0 (+db_0
//   partials) | dW_0 | dW_1 | dW_2 (split-K, deterministic fix-up) | finalize (bias / dense / domain-emb grads,
//   loss, AUC accumulators)            [the optimizer apply is a separate C-ABI call]
// GEMM orientation (tc_gemm.cuh: C[m,n] = sum_k A(m,k) B(n,k), M tile = 128 batch rows or 128 in-features):
//   fwd  : A = H_l   [rows, K]   K-major ; B = W_l [K, N]      MN-major  -> rows x out
//   dH   : A = dZ_l  [rows, out] K-major ; B = W_l [in, out]   K-major   -> rows x in
//   dW   : A = H_l   [rows, in]  MN-major; B = dZ_l [rows, out] MN-major -> in x out, K = rows (split-K)
#pragma once
#include <unordered_map>

#include "mlp_ws.cuh"
#include "philox.cuh"
#include "tc_gemm.cuh"

namespace mlptc {

using namespace mlpws;

// ---- tensor-map cache ---------------------------------------------------------------------------------------
struct TmapKey {
    const void* p;
    uint64_t    rows, cols, ld;
    uint32_t    br, bc, swz;
    bool operator==(const TmapKey& o) const {
        return p == o.p && rows == o.rows && cols == o.cols && ld == o.ld && br == o.br && bc == o.bc && swz == o.swz;
    }
};
struct TmapHash {
    size_t operator()(const TmapKey& k) const {
        size_t h = (size_t)k.p;
        auto mix = [&](uint64_t v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
        mix(k.rows); mix(k.cols); mix(k.ld); mix(k.br); mix(k.bc); mix(k.swz);
        return h;
    }
};
typedef std::unordered_map<TmapKey, CUtensorMap, TmapHash> TmapCache;

inline bool get_tmap(mamdr_ctx* ctx, CUtensorMap* out, const float* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t br,
                     uint32_t bc, CUtensorMapSwizzle swz) {
    if (!ctx->tmap_cache) ctx->tmap_cache = new TmapCache();
    TmapCache& c = *static_cast<TmapCache*>(ctx->tmap_cache);
    TmapKey k{base, rows, cols, ld, br, bc, (uint32_t)swz};
    auto it = c.find(k);
    if (it != c.end()) {
        *out = it->second;
        return true;
    }
    if (c.size() > 65536) c.clear();
    if (!tc::make_tmap_2d_f32(out, base, rows, cols, ld, br, bc, swz)) return false;
    c.emplace(k, *out);
    return true;
}
inline void free_tmap_cache(mamdr_ctx* ctx) {
    delete static_cast<TmapCache*>(ctx->tmap_cache);
    ctx->tmap_cache = nullptr;
}

// K-major operand [rows, K]: box = [box_rows, 32];   MN-major operand [K, mn]: box = [32 k, 32 mn]
inline bool kmajor_map(mamdr_ctx* ctx, CUtensorMap* m, const float* p, uint64_t rows, uint64_t K, uint32_t box_rows) {
    return get_tmap(ctx, m, p, rows, K, K, box_rows, 32, CU_TENSOR_MAP_SWIZZLE_128B);
}
inline bool mnmajor_map(mamdr_ctx* ctx, CUtensorMap* m, const float* p, uint64_t K, uint64_t mn) {
    return get_tmap(ctx, m, p, K, mn, mn, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
}

// ---- per-step partial buffers (in the workspace "colsum" + "partials" regions are reused by name below) ---------
constexpr int kMaxMTiles = 64;   // up to 8192 batch rows

struct HeadPartials {            // written by the head epilogue, one slot per 128-row tile
    float*  dw;        // [mtiles][n_last]
    float*  db_last;   // [mtiles][n_last]
    double* loss;      // [mtiles]
    float*  dg;        // [mtiles]
    int*    hist;      // [2][T+1] global AUC bins of this batch (int atomics => deterministic)
};

// ---- epilogues --------------------------------------------------------------------------------------------------
struct FwdHiddenEpi {   // H_out = dropout(relu(acc + bias)), plus its tf32 lo part
    struct State {};
    const float* bias;
    float*       out;
    float*       out_lo;     // may be NULL (1-pass TF32)
    int          N;
    const OptState* state;   // NULL in inference
    DropoutParams dp;
    __device__ __forceinline__ void begin(State&, int, bool) const {}
    __device__ __forceinline__ void cols(State&, int row, bool valid, int col0, int, float* v) const {
        if (!valid) return;
        DropoutParams q = dp;
        if (dp.enabled) q.step = (uint32_t)(state->step & 0xffffffffll);
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
            const float4 b = ldg_f4(bias + col0 + j);
            float h[4] = {fmaxf(v[j] + b.x, 0.f), fmaxf(v[j + 1] + b.y, 0.f), fmaxf(v[j + 2] + b.z, 0.f), fmaxf(v[j + 3] + b.w, 0.f)};
            if (dp.enabled) {
                const uint4 w = dropout_words4(q, (uint32_t)row * (uint32_t)N + (uint32_t)(col0 + j));
                h[0] = dropout_apply(q, w.x, h[0]); h[1] = dropout_apply(q, w.y, h[1]);
                h[2] = dropout_apply(q, w.z, h[2]); h[3] = dropout_apply(q, w.w, h[3]);
            }
            const int64_t o = (int64_t)row * N + col0 + j;
            *reinterpret_cast<float4*>(out + o) = make_float4(h[0], h[1], h[2], h[3]);
            if (out_lo)
                *reinterpret_cast<float4*>(out_lo + o) = make_float4(tcg::tf32_lo(h[0]), tcg::tf32_lo(h[1]), tcg::tf32_lo(h[2]), tcg::tf32_lo(h[3]));
        }
    }
    __device__ __forceinline__ void end(State&, int, bool, int, int, unsigned char*) const {}
};

template <int NL>   // NL = width of the last hidden layer = BN of this GEMM (one n-tile)
struct FwdHeadEpi { // last hidden layer + Dense(1) + sigmoid + BCE + ds + dZ_last + per-CTA partials + AUC bins
    struct State {
        float h[NL];
        float z;
    };
    const float* bias;
    const float* w;        // dense_kernel [NL]
    const float* g;        // global_bias [1]
    const float* y;        // labels [rows]
    const OptState* state;
    DropoutParams dp;
    int          rows_total;
    int          train;
    float        inv_keep;
    float*       p_out;    // [rows]
    float*       probs;    // optional
    float*       ds_out;   // [rows]
    float*       dZ;       // [rows, NL]
    float*       dZ_lo;    // may be NULL
    HeadPartials part;
    const float* thr;      // AUC thresholds or NULL
    int          T;

    __device__ __forceinline__ void begin(State& st, int, bool) const { st.z = 0.f; }
    __device__ __forceinline__ void cols(State& st, int row, bool valid, int col0, int lc, float* v) const {
        DropoutParams q = dp;
        if (dp.enabled) q.step = (uint32_t)(state->step & 0xffffffffll);
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
            const float4 b = ldg_f4(bias + col0 + j);
            float h[4] = {fmaxf(v[j] + b.x, 0.f), fmaxf(v[j + 1] + b.y, 0.f), fmaxf(v[j + 2] + b.z, 0.f), fmaxf(v[j + 3] + b.w, 0.f)};
            if (dp.enabled) {
                const uint4 wd = dropout_words4(q, (uint32_t)row * (uint32_t)NL + (uint32_t)(col0 + j));
                h[0] = dropout_apply(q, wd.x, h[0]); h[1] = dropout_apply(q, wd.y, h[1]);
                h[2] = dropout_apply(q, wd.z, h[2]); h[3] = dropout_apply(q, wd.w, h[3]);
            }
            const float4 wv = ldg_f4(w + col0 + j);
#pragma unroll
            for (int t = 0; t < 4; ++t) st.h[lc + j + t] = valid ? h[t] : 0.f;
            st.z = fmaf(h[0], wv.x, st.z); st.z = fmaf(h[1], wv.y, st.z);
            st.z = fmaf(h[2], wv.z, st.z); st.z = fmaf(h[3], wv.w, st.z);
        }
    }
    __device__ __forceinline__ void end(State& st, int row, bool valid, int m_tile, int, unsigned char* scratch) const {
        const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
        const float lo = 1e-7f, hi = 1.0f - 1e-7f;
        float dsv = 0.f, pv = 0.f;
        double bce = 0.0;
        if (valid) {
            const float s = st.z + g[0];
            pv = 1.0f / (1.0f + expf(-s));
            const float yv = y[row];
            const float ph = fminf(fmaxf(pv, lo), hi);
            const float lg = logf(ph / (1.0f - ph));
            bce = (double)(fmaxf(lg, 0.f) - lg * yv + log1pf(expf(-fabsf(lg))));
            p_out[row] = pv;
            if (probs) probs[row] = pv;
            if (train) {
                dsv = (pv >= lo && pv <= hi) ? __fdiv_rn(__fsub_rn(pv, yv), (float)rows_total) : 0.f;
                ds_out[row] = dsv;
            }
            if (thr) {
                int lo_i = 0, hi_i = T;
                while (lo_i < hi_i) {
                    const int mid = (lo_i + hi_i) >> 1;
                    if (__ldg(thr + mid) < pv) lo_i = mid + 1; else hi_i = mid;
                }
                atomicAdd(&part.hist[(yv != 0.f ? (T + 1) : 0) + lo_i], 1);
            }
        }
        // ---- per-CTA reductions (fixed order): loss, dg via warp shuffles + smem; dw / db via scratch columns
        double* red_d = reinterpret_cast<double*>(scratch);                 // [4]
        float*  red_f = reinterpret_cast<float*>(scratch + 64);             // [4]
        float*  colbuf = reinterpret_cast<float*>(scratch + 1024);          // [128][NL + 1]
        double bs = bce;
        float  dgs = dsv;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            bs += __shfl_xor_sync(0xffffffffu, bs, o);
            dgs += __shfl_xor_sync(0xffffffffu, dgs, o);
        }
        if (lane == 0) { red_d[warp] = bs; red_f[warp] = dgs; }
        if (train) {
            // dZ_last = (ds * w) * inv_keep * 1[h > 0]; stage h*ds for the dense-kernel gradient
#pragma unroll
            for (int c = 0; c < NL; c += 4) {
                const float4 wv = ldg_f4(w + c);
                const float wq[4] = {wv.x, wv.y, wv.z, wv.w};
                float dz[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const float dh = __fmul_rn(dsv, wq[t]);
                    dz[t] = st.h[c + t] > 0.f ? __fmul_rn(dh, inv_keep) : 0.f;
                    colbuf[tid * (NL + 1) + c + t] = st.h[c + t] * dsv;
                }
                if (valid) {
                    const int64_t o = (int64_t)row * NL + c;
                    *reinterpret_cast<float4*>(dZ + o) = make_float4(dz[0], dz[1], dz[2], dz[3]);
                    if (dZ_lo)
                        *reinterpret_cast<float4*>(dZ_lo + o) = make_float4(tcg::tf32_lo(dz[0]), tcg::tf32_lo(dz[1]), tcg::tf32_lo(dz[2]), tcg::tf32_lo(dz[3]));
                }
#pragma unroll
                for (int t = 0; t < 4; ++t) st.h[c + t] = dz[t];   // reuse the registers for the db pass
            }
        }
        __syncthreads();
        if (tid == 0) {
            part.loss[m_tile] = red_d[0] + red_d[1] + red_d[2] + red_d[3];
            part.dg[m_tile] = red_f[0] + red_f[1] + red_f[2] + red_f[3];
        }
        if (train) {
            // column sums over the 128 rows of the tile: thread -> (column = tid % NL, row half = tid / NL)
            constexpr int GROUPS = 128 / NL;          // NL in {32, 64, 128}
            constexpr int RPG = 128 / GROUPS;
            float* gsum = reinterpret_cast<float*>(scratch + 1024 + 128 * (NL + 1) * 4);   // [GROUPS][NL]
            {
                const int c = tid % NL, gq = tid / NL;
                float s = 0.f;
                for (int r = gq * RPG; r < (gq + 1) * RPG; ++r) s += colbuf[r * (NL + 1) + c];
                gsum[gq * NL + c] = s;
            }
            __syncthreads();
            if (tid < NL) {
                float s = 0.f;
#pragma unroll
                for (int gq = 0; gq < GROUPS; ++gq) s += gsum[gq * NL + tid];
                part.dw[m_tile * NL + tid] = s;
            }
            __syncthreads();
#pragma unroll
            for (int c = 0; c < NL; ++c) colbuf[tid * (NL + 1) + c] = valid ? st.h[c] : 0.f;   // dZ_last
            __syncthreads();
            {
                const int c = tid % NL, gq = tid / NL;
                float s = 0.f;
                for (int r = gq * RPG; r < (gq + 1) * RPG; ++r) s += colbuf[r * (NL + 1) + c];
                gsum[gq * NL + c] = s;
            }
            __syncthreads();
            if (tid < NL) {
                float s = 0.f;
#pragma unroll
                for (int gq = 0; gq < GROUPS; ++gq) s += gsum[gq * NL + tid];
                part.db_last[m_tile * NL + tid] = s;
            }
        }
    }
};

template <int BN>
struct DhEpi {   // dZ_prev = acc * inv_keep * 1[H > 0]  (+ lo part) and per-tile column sums (bias gradient partials)
    struct State {
        float dz[BN];
    };
    const float* H;        // [rows, N] activations of the layer whose mask applies
    float*       out;
    float*       out_lo;   // may be NULL
    float*       db_part;  // [mtiles][N]
    int          N;
    float        inv_keep;
    __device__ __forceinline__ void begin(State&, int, bool) const {}
    __device__ __forceinline__ void cols(State& st, int row, bool valid, int col0, int lc, float* v) const {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
            float r[4] = {0.f, 0.f, 0.f, 0.f};
            if (valid) {
                const int64_t o = (int64_t)row * N + col0 + j;
                const float4 h = *reinterpret_cast<const float4*>(H + o);
                r[0] = h.x > 0.f ? v[j] * inv_keep : 0.f;
                r[1] = h.y > 0.f ? v[j + 1] * inv_keep : 0.f;
                r[2] = h.z > 0.f ? v[j + 2] * inv_keep : 0.f;
                r[3] = h.w > 0.f ? v[j + 3] * inv_keep : 0.f;
                *reinterpret_cast<float4*>(out + o) = make_float4(r[0], r[1], r[2], r[3]);
                if (out_lo)
                    *reinterpret_cast<float4*>(out_lo + o) = make_float4(tcg::tf32_lo(r[0]), tcg::tf32_lo(r[1]), tcg::tf32_lo(r[2]), tcg::tf32_lo(r[3]));
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) st.dz[lc + j + t] = r[t];
        }
    }
    __device__ __forceinline__ void end(State& st, int, bool, int m_tile, int n_tile, unsigned char* scratch) const {
        const int tid = threadIdx.x;
        float* colbuf = reinterpret_cast<float*>(scratch);                         // [128][BN + 1]
        float* gsum = reinterpret_cast<float*>(scratch + 128 * (BN + 1) * 4);      // [GROUPS][BN]
        constexpr int GROUPS = 128 / BN, RPG = 128 / GROUPS;
#pragma unroll
        for (int c = 0; c < BN; ++c) colbuf[tid * (BN + 1) + c] = st.dz[c];
        __syncthreads();
        {
            const int c = tid % BN, gq = tid / BN;
            float s = 0.f;
            for (int r = gq * RPG; r < (gq + 1) * RPG; ++r) s += colbuf[r * (BN + 1) + c];
            gsum[gq * BN + c] = s;
        }
        __syncthreads();
        if (tid < BN) {
            float s = 0.f;
#pragma unroll
            for (int gq = 0; gq < GROUPS; ++gq) s += gsum[gq * BN + tid];
            db_part[(int64_t)m_tile * N + n_tile * BN + tid] = s;
        }
    }
};

struct StoreEpi {   // dW tile -> gradient arena
    struct State {};
    float* out;
    int    ld;
    __device__ __forceinline__ void begin(State&, int, bool) const {}
    __device__ __forceinline__ void cols(State&, int row, bool valid, int col0, int, float* v) const {
        if (!valid) return;
#pragma unroll
        for (int j = 0; j < 16; j += 4)
            if (col0 + j < ld) *reinterpret_cast<float4*>(out + (int64_t)row * ld + col0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
    __device__ __forceinline__ void end(State&, int, bool, int, int, unsigned char*) const {}
};

// ---- small kernels ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) split_lo_kernel(const float* __restrict__ src, float* __restrict__ lo, int64_t n4) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
        const float4 v = reinterpret_cast<const float4*>(src)[i];
        reinterpret_cast<float4*>(lo)[i] = make_float4(tcg::tf32_lo(v.x), tcg::tf32_lo(v.y), tcg::tf32_lo(v.z), tcg::tf32_lo(v.w));
    }
}

struct FinalizeArgs {
    int          n_layers, mtiles, rows, train;
    int          hidden[MAMDR_MAX_LAYERS];
    const float* db_part[MAMDR_MAX_LAYERS];   // [mtiles][hidden[l]]
    float*       g_bias[MAMDR_MAX_LAYERS];
    const float* dw_part;                     // [mtiles][n_last]
    float*       g_w;
    const float* dg_part;                     // [mtiles]
    float*       g_g;
    const double* loss_part;                  // [mtiles]
    float*       loss;
    const float* Ed;                          // domain table (params)
    float*       g_Ed;
    const float* W0dom;                       // kernel0 rows of the domain block [dd, n1]
    int          n_domain, dd, n1, dom;
    float        l2_emb, frozen_reg;
    int*         hist;                        // [2][T+1], zeroed on exit
    float*       auc_acc;                     // [4][T] or NULL
    int          T;
};

// one CTA: fixed-order reduction of the per-tile partials -> bias / dense / global-bias / domain-emb gradients,
// the Keras loss value, and the AUC accumulators (suffix sums of the batch histogram)
__global__ void __launch_bounds__(1024) finalize_kernel(FinalizeArgs a) {
    __shared__ double sq_part[32];
    __shared__ int hist[2 * 1024];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (a.train) {
        for (int l = 0; l < a.n_layers; ++l) {
            for (int c = tid; c < a.hidden[l]; c += 1024) {
                float s = 0.f;
                for (int m = 0; m < a.mtiles; ++m) s += a.db_part[l][(int64_t)m * a.hidden[l] + c];
                a.g_bias[l][c] = s;
            }
        }
        const int nl = a.hidden[a.n_layers - 1];
        for (int c = tid; c < nl; c += 1024) {
            float s = 0.f;
            for (int m = 0; m < a.mtiles; ++m) s += a.dw_part[(int64_t)m * nl + c];
            a.g_w[c] = s;
        }
        if (tid == 0) {
            float s = 0.f;
            for (int m = 0; m < a.mtiles; ++m) s += a.dg_part[m];
            a.g_g[0] = s;
        }
    }
    // L2 penalty of the domain table (double, fixed order)
    double sq = 0.0;
    for (int i = tid; i < a.n_domain * a.dd; i += 1024) { const double e = a.Ed[i]; sq += e * e; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if (lane == 0) sq_part[warp] = sq;
    __syncthreads();
    if (tid == 0) {
        double bs = 0.0, sqs = 0.0;
        for (int m = 0; m < a.mtiles; ++m) bs += a.loss_part[m];
        for (int wv = 0; wv < 32; ++wv) sqs += sq_part[wv];
        a.loss[0] = (float)(bs / (double)a.rows + (double)a.frozen_reg + (double)a.l2_emb * sqs);
    }
    if (a.train) {
        // gEd = 2*l2*Ed ; gEd[dom, c] += sum_k db0[k] * W0[du+di+c, k]
        const float two_l2 = 2.0f * a.l2_emb;
        for (int i = tid; i < a.n_domain * a.dd; i += 1024)
            if (i / a.dd != a.dom) a.g_Ed[i] = two_l2 * a.Ed[i];
        for (int c = warp; c < a.dd; c += 32) {
            const float* wr = a.W0dom + (int64_t)c * a.n1;
            float s = 0.f;
            for (int k = lane; k < a.n1; k += 32) s = fmaf(a.g_bias[0][k], wr[k], s);   // written above by this CTA, visible after the barrier
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) {
                const int i = a.dom * a.dd + c;
                a.g_Ed[i] = __fadd_rn(__fmul_rn(two_l2, a.Ed[i]), s);
            }
        }
    }
    if (a.auc_acc) {
        const int T1 = a.T + 1;
        int vneg = tid < T1 ? a.hist[tid] : 0;
        int vpos = tid < T1 ? a.hist[T1 + tid] : 0;
        if (tid < T1) { hist[tid] = vneg; hist[T1 + tid] = vpos; a.hist[tid] = 0; a.hist[T1 + tid] = 0; }
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
            const int npos = hist[T1], nneg = hist[0];
            const int tp = hist[T1 + tid + 1], fp = hist[tid + 1];
            a.auc_acc[0 * a.T + tid] += (float)tp;
            a.auc_acc[1 * a.T + tid] += (float)fp;
            a.auc_acc[2 * a.T + tid] += (float)(npos - tp);
            a.auc_acc[3 * a.T + tid] += (float)(nneg - fp);
        }
    }
}

}  // namespace mlptc
