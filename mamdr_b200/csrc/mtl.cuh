// Multi-task towers (BASELINE config #5: MMOE / PLE with num_levels = 1 / SharedBottom) on the fp32 per-mini-batch path;
// included by mlp.cu (reuses its SIMT GEMMs, epilogues and the sigmoid-BCE head kernel).
//
// Replaces the Keras train / test function of the per-domain sub-models `Model(inputs, outputs[t])` that
// /root/reference/model_zoo/DeepMTLCTR/deep_mtl_ctr.py:57-65 compiles over deepctr.models.MMOE / PLE / SharedBottom
// (:25-48).  Numerical contract: SURVEY.md A-9 as restated in oracle/mtl.py.
//
// One training mini-batch of domain t (all on the caller's stream, graph-capturable):
//   assemble X | experts: Le x grouped fwd GEMM (all k experts of a layer in ONE launch, bias + ReLU + dropout fused) |
//   gate DNN fwd | gate_mix (logits, softmax, mixture) | tower fwd | head (sigmoid, BCE, ds, dZ of the last tower layer)
//   | tower dH chain | dMix GEMM | mix_backward (da, softmax backward, masked expert / gate upstream gradients) |
//   gate_out gradient | grouped expert dH chain | gate dH chain | dX = sum_j dZ0_j . W0_j^T (+ gate) as ONE K-segmented
//   GEMM | grouped expert dW (deterministic split-K) | gate / tower dW | all bias gradients in one column-sum
//   launch | domain-embedding gradient | sparse de-duplication of the user / item gradient rows (trainable tables)
#pragma once
#include <algorithm>

namespace mtl {

constexpr int kMaxK = MAMDR_MTL_MAX_K;
static_assert(kMaxK == 8, "gate kernels load the softmax rows as two float4");
constexpr int kSmallSplit = 4;       // split-K bound of a lone narrow forward / dH layer
constexpr int kGroupSplit = 4;       // split-K bound of the grouped expert dW GEMMs
constexpr int kMaxColsumJobs = 80;   // k * Le + Lg + Lt + 1 (domain columns of dX)

struct Ws {
    size_t tickets, hist, X, y, p, ds, uid_b, pid_b;
    size_t E[kMaxK][MAMDR_MAX_LAYERS + 1];   // E[j][0] = X
    size_t dZe[kMaxK][MAMDR_MAX_LAYERS];
    size_t G[MAMDR_MAX_LAYERS + 1], dZg[MAMDR_MAX_LAYERS];
    size_t T[MAMDR_MAX_LAYERS + 1], dZt[MAMDR_MAX_LAYERS];   // T[0] = mix
    size_t a, dlogit, dMix, dX, dsum, partials;
    size_t sp_ids[2], sp_rows[2], sp_n[2], sp_ws;
    size_t total;
    int    dx_c0, dx_ld;                     // dX holds the input columns [dx_c0, in)
};

inline size_t tile_area(int M, int N) { return (size_t)((M + simt::BM - 1) / simt::BM * simt::BM) * ((N + simt::BN - 1) / simt::BN * simt::BN); }

inline Ws ws_layout(const mamdr_mtl_desc& d, int B) {
    Ws w;
    memset(&w, 0, sizeof(w));
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off = align_up(off + bytes, 1024);
        return o;
    };
    const int in = d.emb_dim[0] + d.emb_dim[1] + d.emb_dim[2];
    w.tickets = take((size_t)kMaxTiles * 4);
    w.hist = take(head_part_bytes());   // per-CTA partial records of the head + its ticket
    w.X = take((size_t)B * in * 4);
    w.y = take((size_t)B * 4);
    w.p = take((size_t)B * 4);
    w.ds = take((size_t)B * 4);
    w.uid_b = take((size_t)B * 4);
    w.pid_b = take((size_t)B * 4);
    const int Le = d.n_expert_layers, Lg = d.has_gate ? d.n_gate_layers : 0, Lt = d.n_tower_layers;
    size_t grouped = 0, single = 0;
    for (int j = 0; j < d.k; ++j) {
        w.E[j][0] = w.X;
        for (int l = 0; l < Le; ++l) {
            w.E[j][l + 1] = take((size_t)B * d.expert_hidden[l] * 4);
            w.dZe[j][l] = take((size_t)B * d.expert_hidden[l] * 4);
        }
    }
    int prev = in;
    for (int l = 0; l < Le; ++l) {
        grouped = std::max(grouped, tile_area(prev, d.expert_hidden[l]) * kGroupSplit * d.k);
        prev = d.expert_hidden[l];
    }
    const int e_last = d.expert_hidden[Le - 1];
    w.G[0] = w.X;
    prev = in;
    for (int l = 0; l < Lg; ++l) {
        w.G[l + 1] = take((size_t)B * d.gate_hidden[l] * 4);
        w.dZg[l] = take((size_t)B * d.gate_hidden[l] * 4);
        single = std::max(single, tile_area(prev, d.gate_hidden[l]) * kMaxSplit);
        prev = d.gate_hidden[l];
    }
    w.T[0] = d.has_gate ? take((size_t)B * e_last * 4) : w.E[0][Le];
    prev = e_last;
    for (int l = 0; l < Lt; ++l) {
        w.T[l + 1] = take((size_t)B * d.tower_hidden[l] * 4);
        w.dZt[l] = take((size_t)B * d.tower_hidden[l] * 4);
        single = std::max(single, tile_area(prev, d.tower_hidden[l]) * kMaxSplit);
        prev = d.tower_hidden[l];
    }
    w.a = take((size_t)B * kMaxK * 4);
    w.dlogit = take((size_t)B * kMaxK * 4);
    w.dMix = take((size_t)B * e_last * 4);
    w.dx_c0 = d.emb_trainable ? 0 : d.emb_dim[0] + d.emb_dim[1];
    w.dx_ld = in - w.dx_c0;
    w.dX = take((size_t)B * w.dx_ld * 4);
    w.dsum = take((size_t)d.emb_dim[2] * 4);
    for (int l = 0; l < Lg; ++l) single = std::max(single, tile_area(B, d.gate_hidden[l]) * kSmallSplit);
    for (int l = 0; l < Lt; ++l) single = std::max(single, tile_area(B, d.tower_hidden[l]) * kSmallSplit);
    const size_t dx_part = tile_area(B, in) * (d.k + 1);        // one partial tile set per segment of the dX GEMM
    w.partials = take(std::max(std::max(grouped, single), dx_part) * 4);
    if (d.emb_trainable) {
        for (int t = 0; t < 2; ++t) {
            w.sp_ids[t] = take((size_t)B * 4);
            w.sp_rows[t] = take((size_t)B * d.emb_dim[t] * 4);
            w.sp_n[t] = take(16);
        }
        w.sp_ws = take(2 * mamdr_scatter_workspace_bytes(B));   // one per table
    }
    w.total = off;
    return w;
}

// ---- gate: logits = G_last . Gout [g, k] ; a = softmax(logits) ; mix = sum_j a_j * E_j   (one warp per row)
struct GateMixArgs {
    const float* Gl;            // [rows, g]
    const float* Gout;          // [g, k]
    const float* E[kMaxK];      // [rows, n] each
    float *a, *mix;             // [rows, kMaxK], [rows, n]
    int rows, g, k, n;
};

__global__ void __launch_bounds__(256) gate_mix_kernel(GateMixArgs A) {
    extern __shared__ __align__(16) float sG[];   // [g][k]
    for (int i = threadIdx.x; i < A.g * A.k; i += 256) sG[i] = A.Gout[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = blockIdx.x * 8 + warp; r < A.rows; r += gridDim.x * 8) {
        float lg[kMaxK];
#pragma unroll
        for (int j = 0; j < kMaxK; ++j) lg[j] = 0.f;
        for (int c = lane; c < A.g; c += 32) {
            const float h = A.Gl[(int64_t)r * A.g + c];
#pragma unroll
            for (int j = 0; j < kMaxK; ++j)
                if (j < A.k) lg[j] = fmaf(h, sG[c * A.k + j], lg[j]);
        }
#pragma unroll
        for (int j = 0; j < kMaxK; ++j)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) lg[j] += __shfl_xor_sync(0xffffffffu, lg[j], o);
        float mx = lg[0];
#pragma unroll
        for (int j = 1; j < kMaxK; ++j)
            if (j < A.k) mx = fmaxf(mx, lg[j]);
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < kMaxK; ++j)
            if (j < A.k) { lg[j] = expf(lg[j] - mx); sum += lg[j]; }
#pragma unroll
        for (int j = 0; j < kMaxK; ++j) lg[j] = j < A.k ? __fdiv_rn(lg[j], sum) : 0.f;
        if (lane < kMaxK) {
            float v = 0.f;
#pragma unroll
            for (int j = 0; j < kMaxK; ++j)
                if (lane == j) v = lg[j];
            A.a[(int64_t)r * kMaxK + lane] = v;
        }
        for (int c = lane * 4; c < A.n; c += 128) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < kMaxK; ++j)
                if (j < A.k) {
                    const float4 e = *reinterpret_cast<const float4*>(A.E[j] + (int64_t)r * A.n + c);
                    acc.x = __fadd_rn(acc.x, __fmul_rn(lg[j], e.x));
                    acc.y = __fadd_rn(acc.y, __fmul_rn(lg[j], e.y));
                    acc.z = __fadd_rn(acc.z, __fmul_rn(lg[j], e.z));
                    acc.w = __fadd_rn(acc.w, __fmul_rn(lg[j], e.w));
                }
            *reinterpret_cast<float4*>(A.mix + (int64_t)r * A.n + c) = acc;
        }
    }
}

// ---- backward of the mixture and the softmax gate (one warp per row):
//   da_j = <dMix, E_j> ; dlogit = a * (da - <a, da>) ; dZe_j = (a_j * dMix) * mask(E_j) ;
//   dZg = (dlogit . Gout^T) * mask(G_last)
struct MixBwdArgs {
    const float* dMix;          // [rows, n]
    const float* E[kMaxK];
    const float* a;             // [rows, kMaxK]
    const float* Gl;            // [rows, g]
    const float* Gout;          // [g, k]
    float* dZe[kMaxK];          // [rows, n]
    float *dlogit, *dZg;        // [rows, kMaxK], [rows, g]
    int rows, g, k, n;
    float inv_keep;
};

__global__ void __launch_bounds__(256) mix_backward_kernel(MixBwdArgs A) {
    extern __shared__ __align__(16) float sG[];
    for (int i = threadIdx.x; i < A.g * A.k; i += 256) sG[i] = A.Gout[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = blockIdx.x * 8 + warp; r < A.rows; r += gridDim.x * 8) {
        float da[kMaxK], av[kMaxK];
#pragma unroll
        for (int j = 0; j < kMaxK; ++j) { da[j] = 0.f; av[j] = j < A.k ? A.a[(int64_t)r * kMaxK + j] : 0.f; }
        for (int c = lane * 4; c < A.n; c += 128) {
            const float4 dm = *reinterpret_cast<const float4*>(A.dMix + (int64_t)r * A.n + c);
#pragma unroll
            for (int j = 0; j < kMaxK; ++j)
                if (j < A.k) {
                    const float4 e = *reinterpret_cast<const float4*>(A.E[j] + (int64_t)r * A.n + c);
                    da[j] = fmaf(dm.x, e.x, da[j]);
                    da[j] = fmaf(dm.y, e.y, da[j]);
                    da[j] = fmaf(dm.z, e.z, da[j]);
                    da[j] = fmaf(dm.w, e.w, da[j]);
                    float4 o;
                    o.x = e.x > 0.f ? __fmul_rn(__fmul_rn(av[j], dm.x), A.inv_keep) : 0.f;
                    o.y = e.y > 0.f ? __fmul_rn(__fmul_rn(av[j], dm.y), A.inv_keep) : 0.f;
                    o.z = e.z > 0.f ? __fmul_rn(__fmul_rn(av[j], dm.z), A.inv_keep) : 0.f;
                    o.w = e.w > 0.f ? __fmul_rn(__fmul_rn(av[j], dm.w), A.inv_keep) : 0.f;
                    *reinterpret_cast<float4*>(A.dZe[j] + (int64_t)r * A.n + c) = o;
                }
        }
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < kMaxK; ++j) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) da[j] += __shfl_xor_sync(0xffffffffu, da[j], o);
            s = fmaf(av[j], da[j], s);
        }
        float dl[kMaxK];
#pragma unroll
        for (int j = 0; j < kMaxK; ++j) dl[j] = av[j] * (da[j] - s);
        if (lane < kMaxK) {
            float v = 0.f;
#pragma unroll
            for (int j = 0; j < kMaxK; ++j)
                if (lane == j) v = dl[j];
            A.dlogit[(int64_t)r * kMaxK + lane] = v;
        }
        for (int c = lane; c < A.g; c += 32) {
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < kMaxK; ++j)
                if (j < A.k) acc = fmaf(dl[j], sG[c * A.k + j], acc);
            A.dZg[(int64_t)r * A.g + c] = A.Gl[(int64_t)r * A.g + c] > 0.f ? __fmul_rn(acc, A.inv_keep) : 0.f;
        }
    }
}

// ---- gradient of the gate's output kernel: gGout[c, j] = sum_r G_last[r, c] * dlogit[r, j], fixed order
__global__ void __launch_bounds__(1024)
gate_out_grad_kernel(const float* __restrict__ Gl, const float* __restrict__ dlogit, int rows, int g, int k, float* __restrict__ gGout) {
    const int lx = threadIdx.x & 31, c = blockIdx.x * 32 + lx;
    const int ty = threadIdx.x >> 5;   // 32 row groups
    __shared__ float part[32][32][kMaxK + 1];
    float acc[kMaxK];
#pragma unroll
    for (int j = 0; j < kMaxK; ++j) acc[j] = 0.f;
    if (c < g) {
#pragma unroll 4
        for (int r = ty; r < rows; r += 32) {
            const float h = Gl[(int64_t)r * g + c];
            const float4 d0 = ldg_f4(dlogit + (int64_t)r * kMaxK), d1 = ldg_f4(dlogit + (int64_t)r * kMaxK + 4);
            acc[0] = fmaf(h, d0.x, acc[0]); acc[1] = fmaf(h, d0.y, acc[1]); acc[2] = fmaf(h, d0.z, acc[2]); acc[3] = fmaf(h, d0.w, acc[3]);
            acc[4] = fmaf(h, d1.x, acc[4]); acc[5] = fmaf(h, d1.y, acc[5]); acc[6] = fmaf(h, d1.z, acc[6]); acc[7] = fmaf(h, d1.w, acc[7]);
        }
    }
#pragma unroll
    for (int j = 0; j < kMaxK; ++j) part[ty][lx][j] = acc[j];
    __syncthreads();
    if (ty < k && c < g) {      // row group ty reduces gate column ty
        float t = 0.f;
#pragma unroll
        for (int q = 0; q < 32; ++q) t += part[q][lx][ty];
        gGout[c * k + ty] = t;
    }
}

// ---- column sums with a leading dimension (all bias gradients + the domain columns of dX in one launch)
struct ColJob { const float* src; float* dst; int n, ld; };
struct ColArgs { ColJob job[kMaxColsumJobs]; int rows; };

__global__ void __launch_bounds__(kColsumThreads) colsum_ld_kernel(const __grid_constant__ ColArgs args) {
    const ColJob j = args.job[blockIdx.y];
    if (blockIdx.x * 32 >= j.n) return;
    const int lx = threadIdx.x & 31, c = blockIdx.x * 32 + lx;
    const int ty = threadIdx.x >> 5;
    __shared__ float part[32][33];
    part[ty][lx] = c < j.n ? colsum_thread(j.src, args.rows, j.ld, c, ty) : 0.f;
    __syncthreads();
    if (ty == 0 && c < j.n) {
        float t = 0.f;
#pragma unroll
        for (int q = 0; q < 32; ++q) t += part[q][lx];
        j.dst[c] = t;
    }
}

// gEd = 2 * l2 * Ed ; gEd[dom, :] += column sums of the domain block of dX
__global__ void __launch_bounds__(256)
domain_grad_kernel(const float* __restrict__ Ed, const float* __restrict__ dsum, int n_domain, int dd, int dom, float two_l2,
                   float* __restrict__ gEd) {
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n_domain * dd; i += gridDim.x * 256) {
        const float reg = __fmul_rn(two_l2, Ed[i]);
        gEd[i] = (i / dd == dom) ? __fadd_rn(reg, dsum[i - dom * dd]) : reg;
    }
}

inline DropoutParams dropout_params(const mamdr_mtl_desc& d, bool train, uint32_t stream, uint32_t row0) {
    DropoutParams dp;
    const float keep = 1.0f - d.dropout_rate;
    dp.enabled = (train && d.dropout_rate > 0.f) ? 1 : 0;
    dp.seed = d.dropout_seed + stream;
    dp.step = 0;
    dp.row0 = row0;
    const double thr = floor((double)keep * 4294967296.0);
    dp.threshold = thr >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)thr;
    dp.scale = 1.0f / keep;
    return dp;
}

static int validate(mamdr_ctx* ctx, const mamdr_mtl_desc* d, const mamdr_mtl_domain* dom, const mamdr_batch* b, const void* ws,
                    size_t ws_bytes, const float* ut, const float* it) {
    MAMDR_REQUIRE(ctx, ctx && d && dom && b, MAMDR_E_INVALID, "NULL ctx/desc/domain/batch");
    MAMDR_REQUIRE(ctx, ctx->prog == nullptr, MAMDR_E_INVALID, "per-mini-batch calls cannot be recorded into a program");
    MAMDR_REQUIRE(ctx, d->k >= 1 && d->k <= kMaxK, MAMDR_E_INVALID, "k out of range (1..%d)", kMaxK);
    MAMDR_REQUIRE(ctx, d->has_gate || d->k == 1, MAMDR_E_INVALID, "has_gate = 0 needs k = 1");
    MAMDR_REQUIRE(ctx, d->n_expert_layers >= 1 && d->n_expert_layers <= MAMDR_MAX_LAYERS && d->n_tower_layers >= 1 &&
                           d->n_tower_layers <= MAMDR_MAX_LAYERS && (!d->has_gate || (d->n_gate_layers >= 1 && d->n_gate_layers <= MAMDR_MAX_LAYERS)),
                  MAMDR_E_INVALID, "layer counts out of range");
    for (int i = 0; i < 3; ++i) MAMDR_REQUIRE(ctx, d->emb_dim[i] > 0 && d->emb_dim[i] % 4 == 0, MAMDR_E_INVALID, "emb_dim must be a multiple of 4");
    for (int l = 0; l < d->n_expert_layers; ++l) MAMDR_REQUIRE(ctx, d->expert_hidden[l] > 0 && d->expert_hidden[l] % 4 == 0, MAMDR_E_INVALID, "widths must be multiples of 4");
    for (int l = 0; l < d->n_tower_layers; ++l) MAMDR_REQUIRE(ctx, d->tower_hidden[l] > 0 && d->tower_hidden[l] % 4 == 0, MAMDR_E_INVALID, "widths must be multiples of 4");
    if (d->has_gate) {
        for (int l = 0; l < d->n_gate_layers; ++l) MAMDR_REQUIRE(ctx, d->gate_hidden[l] > 0 && d->gate_hidden[l] % 4 == 0, MAMDR_E_INVALID, "widths must be multiples of 4");
        MAMDR_REQUIRE(ctx, (size_t)d->gate_hidden[d->n_gate_layers - 1] * d->k * 4 <= 40 * 1024, MAMDR_E_UNSUPPORTED, "gate output kernel too large for shared memory");
    }
    MAMDR_REQUIRE(ctx, d->tower_hidden[d->n_tower_layers - 1] <= kHeadMaxN, MAMDR_E_UNSUPPORTED, "last tower layer wider than %d", kHeadMaxN);
    MAMDR_REQUIRE(ctx, d->k * d->n_expert_layers + d->n_gate_layers + d->n_tower_layers + 1 <= kMaxColsumJobs, MAMDR_E_UNSUPPORTED, "too many layers");
    MAMDR_REQUIRE(ctx, d->dropout_rate >= 0.f && d->dropout_rate < 1.f, MAMDR_E_INVALID, "dropout_rate must be in [0,1)");
    MAMDR_REQUIRE(ctx, b->rows >= 1 && b->domain >= 0 && b->domain < d->n_domain && b->domain == dom->domain, MAMDR_E_INVALID,
                  "empty batch, domain id out of range or batch / sub-model domain mismatch");
    MAMDR_REQUIRE(ctx, b->uid_dev && b->pid_dev && b->label_dev, MAMDR_E_INVALID, "NULL batch column");
    MAMDR_REQUIRE(ctx, ws && aligned16(ws), MAMDR_E_INVALID, "workspace NULL or misaligned");
    MAMDR_REQUIRE(ctx, ws_bytes >= ws_layout(*d, b->rows).total, MAMDR_E_WORKSPACE, "workspace too small");
    if (!d->emb_trainable) MAMDR_REQUIRE(ctx, ut && it, MAMDR_E_INVALID, "frozen tables are NULL");
    if (d->emb_trainable) MAMDR_REQUIRE(ctx, b->rows <= mamdr_scatter_max_n(), MAMDR_E_UNSUPPORTED, "batch too large for the sparse-gradient dedup");
    return MAMDR_OK;
}

// one dense layer forward (bias + ReLU + dropout fused) for `ng` independent inputs / kernels of one shape
static int fwd_layer(mamdr_ctx* ctx, int ng, const float* const* A, const float* const* W, const float* const* bias, float* const* out,
                     const uint32_t* streams, const mamdr_mtl_desc* d, bool train, const OptState* state, int rows, int K, int N,
                     float* partials, unsigned int* tickets, cudaStream_t st, uint32_t row0) {
    simt::GemmShape s{rows, N, K, K, N};
    // a lone narrow layer (gate, tower) has too few tiles to fill the GPU: split K, the last CTA of a tile runs the epilogue
    simt::LaunchPlan p = simt::plan(rows, N, K, ng == 1 ? ctx->sm_count : 0, ng == 1 ? kSmallSplit : 1);
    simt::GroupedArgs<FwdEpilogue> ga;
    memset(&ga, 0, sizeof(ga));
    for (int g = 0; g < ng; ++g) {
        ga.A[g] = A[g];
        ga.B[g] = W[g];
        ga.epi[g].bias = bias[g];
        ga.epi[g].out = out[g];
        ga.epi[g].N = N;
        ga.epi[g].state = state;
        ga.epi[g].dp = dropout_params(*d, train, streams[g], row0);
    }
    ga.split = (int)p.grid.z; ga.partial_stride = 0; ga.ticket_stride = 0;
    p.grid.z = ga.split * ng;
    simt::gemm_grouped_kernel<true, true, FwdEpilogue><<<p.grid, simt::THREADS, 0, st>>>(ga, s, p.k_chunk, partials, tickets);
    MAMDR_LAUNCH_OK(ctx);
    return MAMDR_OK;
}

// dZ_prev = (dZ . W^T) * mask(H_prev) for `ng` groups
static int dh_layer(mamdr_ctx* ctx, int ng, const float* const* dZ, const float* const* W, const float* const* Hprev, float* const* out,
                    float inv_keep, int rows, int Kd, int Nd, float* partials, unsigned int* tickets, cudaStream_t st) {
    simt::GemmShape s{rows, Nd, Kd, Kd, Kd};
    simt::LaunchPlan p = simt::plan(rows, Nd, Kd, ng == 1 ? ctx->sm_count : 0, ng == 1 ? kSmallSplit : 1);
    simt::GroupedArgs<DhEpilogue> ga;
    memset(&ga, 0, sizeof(ga));
    for (int g = 0; g < ng; ++g) {
        ga.A[g] = dZ[g];
        ga.B[g] = W[g];
        ga.epi[g] = DhEpilogue{Hprev[g], out[g], Nd, inv_keep};
    }
    ga.split = (int)p.grid.z;
    p.grid.z = ga.split * ng;
    simt::gemm_grouped_kernel<true, false, DhEpilogue><<<p.grid, simt::THREADS, 0, st>>>(ga, s, p.k_chunk, partials, tickets);
    MAMDR_LAUNCH_OK(ctx);
    return MAMDR_OK;
}

// dW = H^T . dZ (reduction over the batch rows, deterministic split-K) for `ng` groups
static int dw_layer(mamdr_ctx* ctx, int ng, const float* const* H, const float* const* dZ, float* const* gW, int rows, int Md, int Nd,
                    unsigned char* ws, const Ws& w, cudaStream_t st) {
    simt::GemmShape s{Md, Nd, rows, Md, Nd};
    const int max_split = ng > 1 ? kGroupSplit : kMaxSplit;
    simt::LaunchPlan p = simt::plan(Md, Nd, rows, std::max(1, 8 * ctx->sm_count / ng), max_split);
    const int tiles = (int)(p.grid.x * p.grid.y);
    MAMDR_REQUIRE(ctx, tiles * ng <= kMaxTiles, MAMDR_E_UNSUPPORTED, "layer too large for the ticket table");
    simt::GroupedArgs<StoreEpilogue> ga;
    memset(&ga, 0, sizeof(ga));
    for (int g = 0; g < ng; ++g) {
        ga.A[g] = H[g];
        ga.B[g] = dZ[g];
        ga.epi[g] = StoreEpilogue{gW[g], Nd};
    }
    ga.split = (int)p.grid.z;
    ga.partial_stride = (int64_t)ga.split * tiles * simt::BM * simt::BN;
    ga.ticket_stride = tiles;
    p.grid.z = ga.split * ng;
    simt::gemm_grouped_kernel<false, true, StoreEpilogue><<<p.grid, simt::THREADS, 0, st>>>(ga, s, p.k_chunk, (float*)(ws + w.partials),
                                                                                          (unsigned int*)(ws + w.tickets));
    MAMDR_LAUNCH_OK(ctx);
    return MAMDR_OK;
}

static int forward(mamdr_ctx* ctx, const mamdr_mtl_desc* d, const mamdr_mtl_domain* dm, const mamdr_batch* b, const float* ut, const float* it,
                   const float* params, unsigned char* ws, const Ws& w, const OptState* state, bool train, cudaStream_t st) {
    const int du = d->emb_dim[0], di = d->emb_dim[1], dd = d->emb_dim[2], in = du + di + dd, rows = b->rows, k = d->k, t = dm->domain;
    const float* Eu = d->emb_trainable ? params + d->off_user_emb : ut;
    const float* Ei = d->emb_trainable ? params + d->off_item_emb : it;
    int rc = mamdr_assemble_batch(ctx, Eu, Ei, params + d->off_domain_emb, b, du, di, dd, (float*)(ws + w.X), (float*)(ws + w.y),
                                  (int32_t*)(ws + w.uid_b), (int32_t*)(ws + w.pid_b), st);
    if (rc) return rc;
    const float *A[kMaxK], *W[kMaxK], *bs[kMaxK];
    float* out[kMaxK];
    uint32_t streams[kMaxK];
    int K = in;
    for (int l = 0; l < d->n_expert_layers; ++l) {
        const int N = d->expert_hidden[l];
        for (int j = 0; j < k; ++j) {
            A[j] = (const float*)(ws + w.E[j][l]);
            W[j] = params + dm->off_expert_kernel[j][l];
            bs[j] = params + dm->off_expert_bias[j][l];
            out[j] = (float*)(ws + w.E[j][l + 1]);
            streams[j] = 8u * (uint32_t)dm->expert_id[j] + (uint32_t)l;
        }
        rc = fwd_layer(ctx, k, A, W, bs, out, streams, d, train, state, rows, K, N, (float*)(ws + w.partials), (unsigned int*)(ws + w.tickets), st, (uint32_t)b->row0);
        if (rc) return rc;
        K = N;
    }
    const int e_last = K;
    if (d->has_gate) {
        K = in;
        for (int l = 0; l < d->n_gate_layers; ++l) {
            const int N = d->gate_hidden[l];
            A[0] = (const float*)(ws + w.G[l]);
            W[0] = params + dm->off_gate_kernel[l];
            bs[0] = params + dm->off_gate_bias[l];
            out[0] = (float*)(ws + w.G[l + 1]);
            streams[0] = 4096u + 8u * (uint32_t)t + (uint32_t)l;
            rc = fwd_layer(ctx, 1, A, W, bs, out, streams, d, train, state, rows, K, N, (float*)(ws + w.partials), (unsigned int*)(ws + w.tickets), st, (uint32_t)b->row0);
            if (rc) return rc;
            K = N;
        }
        GateMixArgs ga;
        ga.Gl = (const float*)(ws + w.G[d->n_gate_layers]);
        ga.Gout = params + dm->off_gate_out;
        for (int j = 0; j < kMaxK; ++j) ga.E[j] = (const float*)(ws + w.E[j < k ? j : 0][d->n_expert_layers]);
        ga.a = (float*)(ws + w.a);
        ga.mix = (float*)(ws + w.T[0]);
        ga.rows = rows; ga.g = K; ga.k = k; ga.n = e_last;
        gate_mix_kernel<<<(rows + 7) / 8, 256, (size_t)K * k * 4, st>>>(ga);
        MAMDR_LAUNCH_OK(ctx);
    }
    K = e_last;
    for (int l = 0; l < d->n_tower_layers; ++l) {
        const int N = d->tower_hidden[l];
        A[0] = (const float*)(ws + w.T[l]);
        W[0] = params + dm->off_tower_kernel[l];
        bs[0] = params + dm->off_tower_bias[l];
        out[0] = (float*)(ws + w.T[l + 1]);
        streams[0] = 8192u + 8u * (uint32_t)t + (uint32_t)l;
        rc = fwd_layer(ctx, 1, A, W, bs, out, streams, d, train, state, rows, K, N, (float*)(ws + w.partials), (unsigned int*)(ws + w.tickets), st, (uint32_t)b->row0);
        if (rc) return rc;
        K = N;
    }
    return MAMDR_OK;
}

static int head(mamdr_ctx* ctx, const mamdr_mtl_desc* d, const mamdr_mtl_domain* dm, const mamdr_batch* b, const float* params, float* grads,
                unsigned char* ws, const Ws& w, bool train, float* loss, float* probs, float* auc_acc, const float* thr, int T, cudaStream_t st) {
    const int Lt = d->n_tower_layers, nl = d->tower_hidden[Lt - 1];
    MAMDR_REQUIRE(ctx, loss != nullptr, MAMDR_E_INVALID, "loss_dev is NULL");
    if (auc_acc) MAMDR_REQUIRE(ctx, thr && T >= 2 && T + 1 <= kHeadThreads, MAMDR_E_INVALID, "bad AUC thresholds (2 <= T <= 1023)");
    HeadArgs a;
    a.HL = (const float*)(ws + w.T[Lt]);
    a.w = params + dm->off_tower_out;
    a.g = params + dm->off_bias;
    a.y = (const float*)(ws + w.y);
    a.Ed = params + d->off_domain_emb;
    a.ed_elems = d->n_domain * d->emb_dim[2];
    a.b = b->rows; a.n = nl; a.train = train ? 1 : 0;
    a.inv_keep = (train && d->dropout_rate > 0.f) ? 1.0f / (1.0f - d->dropout_rate) : 1.0f;
    a.l2_emb = d->l2_emb; a.frozen_reg = d->frozen_reg;
    a.p_out = (float*)(ws + w.p); a.probs = probs; a.ds = (float*)(ws + w.ds); a.dZ = (float*)(ws + w.dZt[Lt - 1]);
    a.g_w = train ? grads + dm->off_tower_out : nullptr;
    a.g_g = train ? grads + dm->off_bias : nullptr;
    a.loss = loss; a.auc_acc = auc_acc; a.thr = thr; a.T = auc_acc ? T : 0;
    const size_t smem = head_smem_bytes(nl, a.T);
    MAMDR_REQUIRE(ctx, smem <= 100 * 1024, MAMDR_E_UNSUPPORTED, "head smem %zu too large", smem);
    launch_head(a, ws + w.hist, smem, st);
    MAMDR_LAUNCH_OK(ctx);
    return MAMDR_OK;
}

}  // namespace mtl

extern "C" size_t mamdr_mtl_workspace_bytes(const mamdr_mtl_desc* d, int32_t max_batch) {
    if (!d || max_batch < 1 || d->k < 1 || d->k > mtl::kMaxK || d->n_expert_layers < 1 || d->n_expert_layers > MAMDR_MAX_LAYERS ||
        d->n_tower_layers < 1 || d->n_tower_layers > MAMDR_MAX_LAYERS || (d->has_gate && (d->n_gate_layers < 1 || d->n_gate_layers > MAMDR_MAX_LAYERS)))
        return 0;
    return mtl::ws_layout(*d, max_batch).total;
}

extern "C" int mamdr_mtl_sparse_grads(const mamdr_mtl_desc* desc, int32_t rows, void* ws_dev, int32_t table, const int32_t** uniq_ids_dev,
                                      const float** uniq_rows_dev, const int32_t** n_uniq_dev) {
    if (!desc || !ws_dev || !desc->emb_trainable || table < 0 || table > 1 || rows < 1) return MAMDR_E_INVALID;
    const mtl::Ws w = mtl::ws_layout(*desc, rows);
    unsigned char* ws = (unsigned char*)ws_dev;
    if (uniq_ids_dev) *uniq_ids_dev = (const int32_t*)(ws + w.sp_ids[table]);
    if (uniq_rows_dev) *uniq_rows_dev = (const float*)(ws + w.sp_rows[table]);
    if (n_uniq_dev) *n_uniq_dev = (const int32_t*)(ws + w.sp_n[table]);
    return MAMDR_OK;
}

// dX[:, c0:c0+width] = sum_j dZe0_j . We0_j[c0:c0+width, :]^T (+ the gate's): ONE launch, segments added in gate-column order
static int mtl_input_grad_gemm(mamdr_ctx* ctx, const mamdr_mtl_desc* d, const mamdr_mtl_domain* dm, int rows, const float* params,
                               unsigned char* ws, const mtl::Ws& w, int c0, int width, float* out, cudaStream_t st) {
    simt::SegmentArgs sa;
    memset(&sa, 0, sizeof(sa));
    sa.n = d->k + (d->has_gate ? 1 : 0);
    for (int j = 0; j < sa.n; ++j) {
        const bool gate = j == d->k;
        const int Kd = gate ? d->gate_hidden[0] : d->expert_hidden[0];
        sa.A[j] = (const float*)(ws + (gate ? w.dZg[0] : w.dZe[j][0]));
        sa.B[j] = params + (gate ? dm->off_gate_kernel[0] : dm->off_expert_kernel[j][0]) + (int64_t)c0 * Kd;
        sa.K[j] = Kd;
    }
    StoreEpilogue epi{out, width};
    const dim3 grid((width + simt::BN - 1) / simt::BN, (rows + simt::BM - 1) / simt::BM, sa.n);
    MAMDR_REQUIRE(ctx, (int)(grid.x * grid.y) <= mlpws::kMaxTiles, MAMDR_E_UNSUPPORTED, "batch too large for the ticket table");
    simt::gemm_ksegments_kernel<StoreEpilogue><<<grid, simt::THREADS, 0, st>>>(sa, rows, width, (float*)(ws + w.partials),
                                                                               (unsigned int*)(ws + w.tickets), epi);
    MAMDR_LAUNCH_OK(ctx);
    return MAMDR_OK;
}

// Gradient rows of the gathered user / item embeddings of the LAST mamdr_mtl_train_step of sub-model `dm` (`rows` rows):
// for tables that live outside the arena (row-sharded across GPUs, mamdr_b200/sharded.py).
extern "C" int mamdr_mtl_input_grads(mamdr_ctx* ctx, const mamdr_mtl_desc* d, const mamdr_mtl_domain* dm, int32_t rows, const float* params,
                                     void* ws_, size_t ws_bytes, float* dX_out, mamdr_stream stream) {
    MAMDR_REQUIRE(ctx, ctx && d && dm && params && ws_ && dX_out, MAMDR_E_INVALID, "NULL pointer");
    MAMDR_REQUIRE(ctx, rows >= 1 && aligned16(dX_out) && aligned16(ws_) && aligned16(params), MAMDR_E_INVALID, "bad rows / misaligned pointer");
    MAMDR_REQUIRE(ctx, ctx->prog == nullptr, MAMDR_E_INVALID, "per-mini-batch calls cannot be recorded into a program");
    MAMDR_REQUIRE(ctx, d->k >= 1 && d->k <= mtl::kMaxK && d->n_expert_layers >= 1 && d->n_expert_layers <= MAMDR_MAX_LAYERS, MAMDR_E_INVALID, "bad descriptor");
    const mtl::Ws w = mtl::ws_layout(*d, rows);
    MAMDR_REQUIRE(ctx, ws_bytes >= w.total, MAMDR_E_WORKSPACE, "workspace too small");
    return mtl_input_grad_gemm(ctx, d, dm, rows, params, (unsigned char*)ws_, w, 0, d->emb_dim[0] + d->emb_dim[1], dX_out, (cudaStream_t)stream);
}

extern "C" int mamdr_mtl_eval_step(mamdr_ctx* ctx, const mamdr_mtl_desc* d, const mamdr_mtl_domain* dm, const mamdr_batch* b, const float* ut,
                                   const float* it, const float* params, void* ws_, size_t ws_bytes, float* loss, float* probs,
                                   float* auc_acc, const float* thr, int32_t T, mamdr_stream stream) {
    int rc = mtl::validate(ctx, d, dm, b, ws_, ws_bytes, ut, it);
    if (rc) return rc;
    MAMDR_REQUIRE(ctx, params && aligned16(params), MAMDR_E_INVALID, "params NULL or misaligned");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* ws = (unsigned char*)ws_;
    const mtl::Ws w = mtl::ws_layout(*d, b->rows);
    rc = mtl::forward(ctx, d, dm, b, ut, it, params, ws, w, nullptr, false, st);
    if (rc) return rc;
    return mtl::head(ctx, d, dm, b, params, nullptr, ws, w, false, loss, probs, auc_acc, thr, T, st);
}

extern "C" int mamdr_mtl_train_step(mamdr_ctx* ctx, const mamdr_mtl_desc* d, const mamdr_mtl_domain* dm, const mamdr_batch* b, const float* ut,
                                    const float* it, const float* params, float* grads, void* ws_, size_t ws_bytes, const void* opt_state,
                                    float* loss, float* probs, float* auc_acc, const float* thr, int32_t T, mamdr_stream stream) {
    using namespace mtl;
    int rc = validate(ctx, d, dm, b, ws_, ws_bytes, ut, it);
    if (rc) return rc;
    MAMDR_REQUIRE(ctx, params && grads && aligned16(params) && aligned16(grads), MAMDR_E_INVALID, "params/grads NULL or misaligned");
    MAMDR_REQUIRE(ctx, opt_state != nullptr, MAMDR_E_INVALID, "opt_state is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* ws = (unsigned char*)ws_;
    const Ws w = ws_layout(*d, b->rows);
    const int rows = b->rows, k = d->k, Le = d->n_expert_layers, Lg = d->has_gate ? d->n_gate_layers : 0, Lt = d->n_tower_layers;
    const int du = d->emb_dim[0], di = d->emb_dim[1], dd = d->emb_dim[2], in = du + di + dd;
    const int e_last = d->expert_hidden[Le - 1];
    const float inv_keep = d->dropout_rate > 0.f ? 1.0f / (1.0f - d->dropout_rate) : 1.0f;

    MAMDR_CUDA_OK(ctx, cudaMemsetAsync(ws + w.tickets, 0, (size_t)kMaxTiles * 4, st));
    rc = forward(ctx, d, dm, b, ut, it, params, ws, w, (const OptState*)opt_state, true, st);
    if (rc) return rc;
    rc = head(ctx, d, dm, b, params, grads, ws, w, true, loss, probs, auc_acc, thr, T, st);
    if (rc) return rc;

    const float *A[kMaxK], *W[kMaxK], *Hp[kMaxK];
    float* out[kMaxK];
    // ---- tower: dZt_{l-1} = (dZt_l . Wt_l^T) * mask(T_l)
    for (int l = Lt - 1; l >= 1; --l) {
        A[0] = (const float*)(ws + w.dZt[l]); W[0] = params + dm->off_tower_kernel[l];
        Hp[0] = (const float*)(ws + w.T[l]); out[0] = (float*)(ws + w.dZt[l - 1]);
        rc = dh_layer(ctx, 1, A, W, Hp, out, inv_keep, rows, d->tower_hidden[l], d->tower_hidden[l - 1], (float*)(ws + w.partials), (unsigned int*)(ws + w.tickets), st);
        if (rc) return rc;
    }
    // ---- gradient w.r.t. the tower input (the mixture; for SharedBottom the bottom's output, masked like a hidden layer)
    {
        const int Kd = d->tower_hidden[0];
        simt::GemmShape s{rows, e_last, Kd, Kd, Kd};
        simt::LaunchPlan p = simt::plan(rows, e_last, Kd, 0, 1);
        if (d->has_gate) {
            StoreEpilogue epi{(float*)(ws + w.dMix), e_last};
            simt::gemm_kernel<true, false, StoreEpilogue><<<p.grid, simt::THREADS, 0, st>>>((const float*)(ws + w.dZt[0]), params + dm->off_tower_kernel[0], s,
                                                                                            p.k_chunk, nullptr, nullptr, epi);
        } else {
            DhEpilogue epi{(const float*)(ws + w.E[0][Le]), (float*)(ws + w.dZe[0][Le - 1]), e_last, inv_keep};
            simt::gemm_kernel<true, false, DhEpilogue><<<p.grid, simt::THREADS, 0, st>>>((const float*)(ws + w.dZt[0]), params + dm->off_tower_kernel[0], s,
                                                                                         p.k_chunk, nullptr, nullptr, epi);
        }
        MAMDR_LAUNCH_OK(ctx);
    }
    if (d->has_gate) {
        const int g = d->gate_hidden[Lg - 1];
        MixBwdArgs ma;
        ma.dMix = (const float*)(ws + w.dMix);
        for (int j = 0; j < kMaxK; ++j) {
            ma.E[j] = (const float*)(ws + w.E[j < k ? j : 0][Le]);
            ma.dZe[j] = (float*)(ws + w.dZe[j < k ? j : 0][Le - 1]);
        }
        ma.a = (const float*)(ws + w.a);
        ma.Gl = (const float*)(ws + w.G[Lg]);
        ma.Gout = params + dm->off_gate_out;
        ma.dlogit = (float*)(ws + w.dlogit);
        ma.dZg = (float*)(ws + w.dZg[Lg - 1]);
        ma.rows = rows; ma.g = g; ma.k = k; ma.n = e_last; ma.inv_keep = inv_keep;
        mix_backward_kernel<<<(rows + 7) / 8, 256, (size_t)g * k * 4, st>>>(ma);
        MAMDR_LAUNCH_OK(ctx);
        gate_out_grad_kernel<<<(g + 31) / 32, 1024, 0, st>>>((const float*)(ws + w.G[Lg]), (const float*)(ws + w.dlogit), rows, g, k,
                                                          grads + dm->off_gate_out);
        MAMDR_LAUNCH_OK(ctx);
    }
    // ---- experts: dZe_{l-1} = (dZe_l . We_l^T) * mask(E_l), all k experts of a layer in one launch ; gate likewise
    for (int l = Le - 1; l >= 1; --l) {
        for (int j = 0; j < k; ++j) {
            A[j] = (const float*)(ws + w.dZe[j][l]); W[j] = params + dm->off_expert_kernel[j][l];
            Hp[j] = (const float*)(ws + w.E[j][l]); out[j] = (float*)(ws + w.dZe[j][l - 1]);
        }
        rc = dh_layer(ctx, k, A, W, Hp, out, inv_keep, rows, d->expert_hidden[l], d->expert_hidden[l - 1], (float*)(ws + w.partials), (unsigned int*)(ws + w.tickets), st);
        if (rc) return rc;
    }
    for (int l = Lg - 1; l >= 1; --l) {
        A[0] = (const float*)(ws + w.dZg[l]); W[0] = params + dm->off_gate_kernel[l];
        Hp[0] = (const float*)(ws + w.G[l]); out[0] = (float*)(ws + w.dZg[l - 1]);
        rc = dh_layer(ctx, 1, A, W, Hp, out, inv_keep, rows, d->gate_hidden[l], d->gate_hidden[l - 1], (float*)(ws + w.partials), (unsigned int*)(ws + w.tickets), st);
        if (rc) return rc;
    }
    // ---- dX[:, c0:in] = sum_j dZe0_j . We0_j[c0:in, :]^T (+ the gate's): ONE launch, segments added in gate-column order
    rc = mtl_input_grad_gemm(ctx, d, dm, rows, params, ws, w, w.dx_c0, w.dx_ld, (float*)(ws + w.dX), st);
    if (rc) return rc;
    // ---- kernels: dW = H^T . dZ
    {
        float* gW[kMaxK];
        for (int l = 0; l < Le; ++l) {
            for (int j = 0; j < k; ++j) {
                A[j] = (const float*)(ws + w.E[j][l]); W[j] = (const float*)(ws + w.dZe[j][l]); gW[j] = grads + dm->off_expert_kernel[j][l];
            }
            rc = dw_layer(ctx, k, A, W, gW, rows, l == 0 ? in : d->expert_hidden[l - 1], d->expert_hidden[l], ws, w, st);
            if (rc) return rc;
        }
        for (int l = 0; l < Lg; ++l) {
            A[0] = (const float*)(ws + w.G[l]); W[0] = (const float*)(ws + w.dZg[l]); gW[0] = grads + dm->off_gate_kernel[l];
            rc = dw_layer(ctx, 1, A, W, gW, rows, l == 0 ? in : d->gate_hidden[l - 1], d->gate_hidden[l], ws, w, st);
            if (rc) return rc;
        }
        for (int l = 0; l < Lt; ++l) {
            A[0] = (const float*)(ws + w.T[l]); W[0] = (const float*)(ws + w.dZt[l]); gW[0] = grads + dm->off_tower_kernel[l];
            rc = dw_layer(ctx, 1, A, W, gW, rows, l == 0 ? e_last : d->tower_hidden[l - 1], d->tower_hidden[l], ws, w, st);
            if (rc) return rc;
        }
    }
    // ---- every bias gradient + the domain columns of dX: one column-sum launch
    {
        ColArgs ca;
        memset(&ca, 0, sizeof(ca));
        int nj = 0, maxn = dd;
        for (int j = 0; j < k; ++j)
            for (int l = 0; l < Le; ++l) ca.job[nj++] = ColJob{(const float*)(ws + w.dZe[j][l]), grads + dm->off_expert_bias[j][l], d->expert_hidden[l], d->expert_hidden[l]};
        for (int l = 0; l < Lg; ++l) ca.job[nj++] = ColJob{(const float*)(ws + w.dZg[l]), grads + dm->off_gate_bias[l], d->gate_hidden[l], d->gate_hidden[l]};
        for (int l = 0; l < Lt; ++l) ca.job[nj++] = ColJob{(const float*)(ws + w.dZt[l]), grads + dm->off_tower_bias[l], d->tower_hidden[l], d->tower_hidden[l]};
        ca.job[nj++] = ColJob{(const float*)(ws + w.dX) + (du + di - w.dx_c0), (float*)(ws + w.dsum), dd, w.dx_ld};
        for (int q = 0; q < nj; ++q) maxn = std::max(maxn, ca.job[q].n);
        ca.rows = rows;
        colsum_ld_kernel<<<dim3((maxn + 31) / 32, nj), kColsumThreads, 0, st>>>(ca);
        MAMDR_LAUNCH_OK(ctx);
        domain_grad_kernel<<<(d->n_domain * dd + 255) / 256, 256, 0, st>>>(params + d->off_domain_emb, (const float*)(ws + w.dsum), d->n_domain, dd,
                                                                          dm->domain, 2.0f * d->l2_emb, grads + d->off_domain_emb);
        MAMDR_LAUNCH_OK(ctx);
    }
    if (d->emb_trainable) {   // both tables' sparse gradients: one sort launch + one segment-sum launch
        DedupJob jobs[2];
        for (int t = 0; t < 2; ++t) {
            jobs[t] = DedupJob{(const int32_t*)(ws + (t == 0 ? w.uid_b : w.pid_b)), (const float*)(ws + w.dX) + (t == 0 ? 0 : du), w.dx_ld,
                               d->emb_dim[t], (int32_t*)(ws + w.sp_ids[t]), (float*)(ws + w.sp_rows[t]), (int32_t*)(ws + w.sp_n[t]), nullptr, nullptr};
            mamdr_scatter_job_ws(&jobs[t], ws + w.sp_ws + t * mamdr_scatter_workspace_bytes(rows), rows);
        }
        rc = mamdr_scatter_dedup_jobs(ctx, jobs, 2, rows, st);
        if (rc) return rc;
    }
    return MAMDR_OK;
}
