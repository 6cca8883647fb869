// STAR tower (BASELINE config #4) on the fp32 per-mini-batch path; included by mlp.cu (reuses its SIMT GEMMs, head and
// column-sum kernels).
//
// Replaces the Keras train / test function of the model built in /root/reference/model_zoo/Star/star.py:70-113:
// PartitionedNorm (Star/partitioned_norm.py:102-203) -> StarFCN x L (Star/star_fcn.py:105-139) -> Dense(1, sigmoid),
// BCE loss.  Numerical contract: SURVEY.md A-8 as restated in oracle/star.py (incl. the zero-debiased moving statistics
// and the exactly-zero gradient of the batch-constant domain-embedding columns).
//
// One training mini-batch (all on the caller's stream, graph-capturable):
//   memset(grads) | assemble X | pn_stats (+ moving statistics) | pn_apply -> xhat, H_0 | effective weights
//   W = W_sh * W_sp[d], b = b_sh + b_sp[d] | L x fwd GEMM(+bias+ReLU) | head | (L-1) x dH GEMM(+mask) + dY GEMM |
//   L x dW_eff GEMM (deterministic split-K) | colsum(db_eff) | star_grads (shared / specific[d] products) | pn_backward
#pragma once

struct StarWs {
    size_t tickets, hist, X, xhat, H[MAMDR_MAX_LAYERS + 1], dZ[MAMDR_MAX_LAYERS], dY, Weff[MAMDR_MAX_LAYERS], beff[MAMDR_MAX_LAYERS],
        dWeff[MAMDR_MAX_LAYERS], dbeff[MAMDR_MAX_LAYERS], mean, rstd, y, p, ds, uid_b, pid_b, partials, total;
};

inline StarWs star_ws(const mamdr_star_desc& d, int B) {
    StarWs w;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off = align_up(off + bytes, 1024);
        return o;
    };
    const int n = d.emb_dim[0] + d.emb_dim[1] + d.emb_dim[2];
    w.tickets = take((size_t)kMaxTiles * 4);
    w.hist = take(head_part_bytes());   // per-CTA partial records of the head + its ticket
    w.X = take((size_t)B * n * 4);
    w.xhat = take((size_t)B * n * 4);
    w.H[0] = take((size_t)B * n * 4);
    w.dY = take((size_t)B * n * 4);
    size_t max_mn = 0;
    int prev = n;
    for (int l = 0; l < d.n_layers; ++l) {
        w.H[l + 1] = take((size_t)B * d.hidden[l] * 4);
        w.dZ[l] = take((size_t)B * d.hidden[l] * 4);
        w.Weff[l] = take((size_t)prev * d.hidden[l] * 4);
        w.beff[l] = take((size_t)d.hidden[l] * 4);
        w.dWeff[l] = take((size_t)prev * d.hidden[l] * 4);
        w.dbeff[l] = take((size_t)d.hidden[l] * 4);
        const size_t mn = (size_t)((prev + 127) / 128 * 128) * ((d.hidden[l] + 63) / 64 * 64);
        if (mn > max_mn) max_mn = mn;
        prev = d.hidden[l];
    }
    w.mean = take((size_t)n * 4);
    w.rstd = take((size_t)n * 4);
    w.y = take((size_t)B * 4);
    w.p = take((size_t)B * 4);
    w.ds = take((size_t)B * 4);
    w.uid_b = take((size_t)B * 4);
    w.pid_b = take((size_t)B * 4);
    w.partials = take(max_mn * kMaxSplit * 4);
    w.total = off;
    return w;
}

// non-trainable PartitionedNorm state: moving_mean | moving_var | biased_mean | biased_var, each [D][n]; then int steps[D], ticket
struct PnState {
    float *moving_mean, *moving_var, *biased_mean, *biased_var;
    int* steps;
    unsigned int* ticket;
};
inline PnState pn_state_view(const mamdr_star_desc& d, void* p) {
    const size_t dn = (size_t)d.n_domain * (d.emb_dim[0] + d.emb_dim[1] + d.emb_dim[2]);
    float* f = (float*)p;
    PnState s;
    s.moving_mean = f; s.moving_var = f + dn; s.biased_mean = f + 2 * dn; s.biased_var = f + 3 * dn;
    s.steps = (int*)(f + 4 * dn);
    s.ticket = (unsigned int*)(s.steps + d.n_domain);
    return s;
}

// ---- PartitionedNorm statistics: per column mean / biased variance over the batch rows (two passes, fixed order);
// columns >= n_var (the batch-constant domain-embedding block) are centred exactly: mean = x, var = 0.
// train: also the zero-debiased moving statistics of the batch's domain.  eval: mean / rstd from the moving statistics.
__global__ void __launch_bounds__(256)
pn_stats_kernel(const float* __restrict__ X, int rows, int n, int n_var, int dom, float eps, float momentum, int train, PnState st,
                float* __restrict__ mean_out, float* __restrict__ rstd_out) {
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int ty = threadIdx.x >> 5;
    __shared__ float part[8][32];
    __shared__ float smean[32];
    __shared__ bool last;
    float mean = 0.f, var = 0.f;
    if (!train) {
        if (ty == 0 && c < n) {
            mean_out[c] = st.moving_mean[(size_t)dom * n + c];
            rstd_out[c] = 1.0f / sqrtf(st.moving_var[(size_t)dom * n + c] + eps);
        }
        return;
    }
    float s = 0.f;
    if (c < n_var)
        for (int r = ty; r < rows; r += 8) s += X[(size_t)r * n + c];
    part[ty][threadIdx.x & 31] = s;
    __syncthreads();
    if (ty == 0) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += part[k][threadIdx.x & 31];
        smean[threadIdx.x & 31] = c < n_var ? t / (float)rows : (c < n ? X[c] : 0.f);
    }
    __syncthreads();
    mean = smean[threadIdx.x & 31];
    s = 0.f;
    if (c < n_var)
        for (int r = ty; r < rows; r += 8) { const float dlt = X[(size_t)r * n + c] - mean; s += dlt * dlt; }
    __syncthreads();
    part[ty][threadIdx.x & 31] = s;
    __syncthreads();
    const int t_new = st.steps[dom] + 1;   // every block reads the old count before the last block bumps it
    if (ty == 0 && c < n) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += part[k][threadIdx.x & 31];
        var = c < n_var ? t / (float)rows : 0.f;
        mean_out[c] = mean;
        rstd_out[c] = 1.0f / sqrtf(var + eps);
        const size_t o = (size_t)dom * n + c;
        const float bm = __fadd_rn(__fmul_rn(st.biased_mean[o], momentum), __fmul_rn(mean, 1.0f - momentum));
        const float bv = __fadd_rn(__fmul_rn(st.biased_var[o], momentum), __fmul_rn(var, 1.0f - momentum));
        const float corr = (float)(1.0 - pow((double)momentum, (double)t_new));
        st.biased_mean[o] = bm;
        st.biased_var[o] = bv;
        st.moving_mean[o] = bm / corr;
        st.moving_var[o] = bv / corr;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(st.ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (last && threadIdx.x == 0) {
        st.steps[dom] = t_new;
        *st.ticket = 0;
    }
}

// xhat = (x - mean) * rstd (exactly 0 on the batch-constant columns in training) ; H_0 = xhat * gamma_sh*gamma_sp[d] + beta_sh + beta_sp[d]
__global__ void __launch_bounds__(256)
pn_apply_kernel(const float* __restrict__ X, int rows, int n, int n_var, int train, const float* __restrict__ mean,
                const float* __restrict__ rstd, const float* __restrict__ g_sh, const float* __restrict__ g_sp,
                const float* __restrict__ b_sh, const float* __restrict__ b_sp, float* __restrict__ xhat, float* __restrict__ Y) {
    const size_t total = (size_t)rows * n / 4;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
        const int c = (int)((i * 4) % n);
        const float4 x = *reinterpret_cast<const float4*>(X + i * 4);
        const float4 mu = ldg_f4(mean + c), rs = ldg_f4(rstd + c);
        const float4 gs = ldg_f4(g_sh + c), gp = ldg_f4(g_sp + c), bs = ldg_f4(b_sh + c), bp = ldg_f4(b_sp + c);
        const float xv[4] = {x.x, x.y, x.z, x.w}, m4[4] = {mu.x, mu.y, mu.z, mu.w}, r4[4] = {rs.x, rs.y, rs.z, rs.w};
        const float g4[4] = {gs.x * gp.x, gs.y * gp.y, gs.z * gp.z, gs.w * gp.w}, b4[4] = {bs.x + bp.x, bs.y + bp.y, bs.z + bp.z, bs.w + bp.w};
        float xh[4], y[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            xh[k] = (train && c + k >= n_var) ? 0.f : (xv[k] - m4[k]) * r4[k];
            y[k] = xh[k] * g4[k] + b4[k];
        }
        *reinterpret_cast<float4*>(xhat + i * 4) = make_float4(xh[0], xh[1], xh[2], xh[3]);
        *reinterpret_cast<float4*>(Y + i * 4) = make_float4(y[0], y[1], y[2], y[3]);
    }
}

struct StarLayerPtrs {
    const float *k_sh, *k_sp, *b_sh, *b_sp;   // k_sp / b_sp already point at the batch's domain slice
    float *w_eff, *b_eff;
    int kn, n;                                // kernel elements, bias elements
};
struct StarEffArgs { StarLayerPtrs l[MAMDR_MAX_LAYERS]; int n_layers; };

__global__ void __launch_bounds__(256) star_eff_kernel(StarEffArgs a) {
    const StarLayerPtrs L = a.l[blockIdx.y];
    for (int i = (blockIdx.x * 256 + threadIdx.x) * 4; i < L.kn; i += gridDim.x * 256 * 4) {
        const float4 s = ldg_f4(L.k_sh + i), p = ldg_f4(L.k_sp + i);
        *reinterpret_cast<float4*>(L.w_eff + i) = make_float4(s.x * p.x, s.y * p.y, s.z * p.z, s.w * p.w);
    }
    for (int i = (blockIdx.x * 256 + threadIdx.x) * 4; i < L.n; i += gridDim.x * 256 * 4) {
        const float4 s = ldg_f4(L.b_sh + i), p = ldg_f4(L.b_sp + i);
        *reinterpret_cast<float4*>(L.b_eff + i) = make_float4(s.x + p.x, s.y + p.y, s.z + p.z, s.w + p.w);
    }
}

struct StarGradPtrs {
    const float *dW, *db, *k_sh, *k_sp;
    float *g_k_sh, *g_k_sp, *g_b_sh, *g_b_sp;
    int kn, n;
};
struct StarGradArgs { StarGradPtrs l[MAMDR_MAX_LAYERS]; int n_layers; };

// dW_eff -> gradients of the shared kernel (x W_sp[d]) and of the batch's specific slice (x W_sh); biases alike
__global__ void __launch_bounds__(256) star_grad_kernel(StarGradArgs a) {
    const StarGradPtrs L = a.l[blockIdx.y];
    for (int i = (blockIdx.x * 256 + threadIdx.x) * 4; i < L.kn; i += gridDim.x * 256 * 4) {
        const float4 g = *reinterpret_cast<const float4*>(L.dW + i);
        const float4 s = ldg_f4(L.k_sh + i), p = ldg_f4(L.k_sp + i);
        *reinterpret_cast<float4*>(L.g_k_sh + i) = make_float4(g.x * p.x, g.y * p.y, g.z * p.z, g.w * p.w);
        *reinterpret_cast<float4*>(L.g_k_sp + i) = make_float4(g.x * s.x, g.y * s.y, g.z * s.z, g.w * s.w);
    }
    for (int i = (blockIdx.x * 256 + threadIdx.x) * 4; i < L.n; i += gridDim.x * 256 * 4) {
        const float4 g = *reinterpret_cast<const float4*>(L.db + i);
        *reinterpret_cast<float4*>(L.g_b_sh + i) = g;
        *reinterpret_cast<float4*>(L.g_b_sp + i) = g;
    }
}

// per column: s1 = sum_r dY, s2 = sum_r dY * xhat  ->  gamma / beta gradients (shared and the batch's specific slice)
__global__ void __launch_bounds__(256)
pn_backward_kernel(const float* __restrict__ dY, const float* __restrict__ xhat, int rows, int n, const float* __restrict__ g_sh,
                   const float* __restrict__ g_sp, float* __restrict__ gg_sh, float* __restrict__ gg_sp,
                   float* __restrict__ gb_sh, float* __restrict__ gb_sp) {
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int ty = threadIdx.x >> 5;
    __shared__ float p1[8][32], p2[8][32];
    float s1 = 0.f, s2 = 0.f;
    if (c < n)
        for (int r = ty; r < rows; r += 8) {
            const float d = dY[(size_t)r * n + c];
            s1 += d;
            s2 += d * xhat[(size_t)r * n + c];
        }
    p1[ty][threadIdx.x & 31] = s1;
    p2[ty][threadIdx.x & 31] = s2;
    __syncthreads();
    if (ty == 0 && c < n) {
        float t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) { t1 += p1[k][threadIdx.x & 31]; t2 += p2[k][threadIdx.x & 31]; }
        gg_sh[c] = t2 * g_sp[c];
        gg_sp[c] = t2 * g_sh[c];
        gb_sh[c] = t1;
        gb_sp[c] = t1;
    }
}

static int star_validate(mamdr_ctx* ctx, const mamdr_star_desc* d, const mamdr_batch* b, const void* ws, size_t ws_bytes,
                         const float* ut, const float* it, const void* pn_state) {
    MAMDR_REQUIRE(ctx, ctx && d && b, MAMDR_E_INVALID, "NULL ctx/desc/batch");
    MAMDR_REQUIRE(ctx, ctx->prog == nullptr, MAMDR_E_INVALID, "per-mini-batch calls cannot be recorded into a program");
    MAMDR_REQUIRE(ctx, d->n_layers >= 1 && d->n_layers <= MAMDR_MAX_LAYERS, MAMDR_E_INVALID, "n_layers out of range");
    for (int i = 0; i < 3; ++i) MAMDR_REQUIRE(ctx, d->emb_dim[i] > 0 && d->emb_dim[i] % 4 == 0, MAMDR_E_INVALID, "emb_dim must be a multiple of 4");
    for (int l = 0; l < d->n_layers; ++l) MAMDR_REQUIRE(ctx, d->hidden[l] > 0 && d->hidden[l] % 4 == 0, MAMDR_E_INVALID, "hidden widths must be multiples of 4");
    MAMDR_REQUIRE(ctx, d->hidden[d->n_layers - 1] <= kHeadMaxN, MAMDR_E_UNSUPPORTED, "last hidden layer wider than %d", kHeadMaxN);
    MAMDR_REQUIRE(ctx, b->rows >= 1 && b->domain >= 0 && b->domain < d->n_domain, MAMDR_E_INVALID, "empty batch or domain id out of range");
    MAMDR_REQUIRE(ctx, b->uid_dev && b->pid_dev && b->label_dev && ut && it && pn_state, MAMDR_E_INVALID, "NULL batch column / table / state");
    MAMDR_REQUIRE(ctx, ws && aligned16(ws), MAMDR_E_INVALID, "workspace NULL or misaligned");
    MAMDR_REQUIRE(ctx, ws_bytes >= star_ws(*d, b->rows).total, MAMDR_E_WORKSPACE, "workspace too small");
    return MAMDR_OK;
}

static int star_forward(mamdr_ctx* ctx, const mamdr_star_desc* d, const mamdr_batch* b, const float* ut, const float* it,
                        const float* params, void* pn_state, unsigned char* ws, const StarWs& w, bool train, cudaStream_t st) {
    const int du = d->emb_dim[0], di = d->emb_dim[1], dd = d->emb_dim[2], n = du + di + dd, L = d->n_layers, rows = b->rows;
    const int dom = b->domain;
    int rc = mamdr_assemble_batch(ctx, ut, it, params + d->off_domain_emb, b, du, di, dd, (float*)(ws + w.X), (float*)(ws + w.y),
                                  (int32_t*)(ws + w.uid_b), (int32_t*)(ws + w.pid_b), st);
    if (rc) return rc;
    PnState ps = pn_state_view(*d, pn_state);
    pn_stats_kernel<<<(n + 31) / 32, 256, 0, st>>>((const float*)(ws + w.X), rows, n, du + di, dom, d->pn_eps, d->pn_momentum, train ? 1 : 0,
                                                 ps, (float*)(ws + w.mean), (float*)(ws + w.rstd));
    MAMDR_LAUNCH_OK(ctx);
    const int64_t tot4 = (int64_t)rows * n / 4;
    const int grid = (int)((tot4 + 255) / 256 < 1 ? 1 : ((tot4 + 255) / 256 > 1184 ? 1184 : (tot4 + 255) / 256));
    pn_apply_kernel<<<grid, 256, 0, st>>>((const float*)(ws + w.X), rows, n, du + di, train ? 1 : 0, (const float*)(ws + w.mean),
                                          (const float*)(ws + w.rstd), params + d->off_gamma_sh, params + d->off_gamma_sp + (int64_t)dom * n,
                                          params + d->off_beta_sh, params + d->off_beta_sp + (int64_t)dom * n, (float*)(ws + w.xhat),
                                          (float*)(ws + w.H[0]));
    MAMDR_LAUNCH_OK(ctx);
    StarEffArgs ea;
    ea.n_layers = L;
    int K = n;
    for (int l = 0; l < L; ++l) {
        const int N = d->hidden[l];
        ea.l[l] = StarLayerPtrs{params + d->off_ksh[l], params + d->off_ksp[l] + (int64_t)dom * K * N, params + d->off_bsh[l],
                                params + d->off_bsp[l] + (int64_t)dom * N, (float*)(ws + w.Weff[l]), (float*)(ws + w.beff[l]), K * N, N};
        K = N;
    }
    star_eff_kernel<<<dim3(96, L), 256, 0, st>>>(ea);
    MAMDR_LAUNCH_OK(ctx);
    K = n;
    for (int l = 0; l < L; ++l) {
        const int N = d->hidden[l];
        FwdEpilogue epi;
        epi.bias = (const float*)(ws + w.beff[l]);
        epi.out = (float*)(ws + w.H[l + 1]);
        epi.N = N;
        epi.state = nullptr;
        epi.dp.enabled = 0; epi.dp.seed = 0; epi.dp.step = 0; epi.dp.threshold = 0; epi.dp.scale = 1.f; epi.dp.row0 = 0;
        simt::GemmShape s{rows, N, K, K, N};
        simt::LaunchPlan p = simt::plan(rows, N, K, 0, 1);
        simt::gemm_kernel<true, true, FwdEpilogue><<<p.grid, simt::THREADS, 0, st>>>((const float*)(ws + w.H[l]), (const float*)(ws + w.Weff[l]), s,
                                                                                     p.k_chunk, nullptr, nullptr, epi);
        MAMDR_LAUNCH_OK(ctx);
        K = N;
    }
    return MAMDR_OK;
}

static int star_head(mamdr_ctx* ctx, const mamdr_star_desc* d, const mamdr_batch* b, const float* params, float* grads, unsigned char* ws,
                     const StarWs& w, bool train, float* loss, float* probs, float* auc_acc, const float* thr, int T, cudaStream_t st) {
    const int L = d->n_layers, nl = d->hidden[L - 1];
    MAMDR_REQUIRE(ctx, loss != nullptr, MAMDR_E_INVALID, "loss_dev is NULL");
    if (auc_acc) MAMDR_REQUIRE(ctx, thr && T >= 2 && T + 1 <= kHeadThreads, MAMDR_E_INVALID, "bad AUC thresholds (2 <= T <= 1023)");
    HeadArgs a;
    a.HL = (const float*)(ws + w.H[L]);
    a.w = params + d->off_out_kernel;
    a.g = params + d->off_out_bias;
    a.y = (const float*)(ws + w.y);
    a.Ed = params + d->off_domain_emb;
    a.ed_elems = 0;                      // STAR embeddings carry no l2 regulariser
    a.b = b->rows; a.n = nl; a.train = train ? 1 : 0;
    a.inv_keep = 1.0f; a.l2_emb = 0.f; a.frozen_reg = 0.f;
    a.p_out = (float*)(ws + w.p); a.probs = probs; a.ds = (float*)(ws + w.ds); a.dZ = (float*)(ws + w.dZ[L - 1]);
    a.g_w = train ? grads + d->off_out_kernel : nullptr;
    a.g_g = train ? grads + d->off_out_bias : nullptr;
    a.loss = loss; a.auc_acc = auc_acc; a.thr = thr; a.T = auc_acc ? T : 0;
    const size_t smem = head_smem_bytes(nl, a.T);
    MAMDR_REQUIRE(ctx, smem <= 100 * 1024, MAMDR_E_UNSUPPORTED, "head smem %zu too large", smem);
    launch_head(a, ws + w.hist, smem, st);
    MAMDR_LAUNCH_OK(ctx);
    return MAMDR_OK;
}

extern "C" size_t mamdr_star_workspace_bytes(const mamdr_star_desc* d, int32_t max_batch) {
    if (!d || max_batch < 1 || d->n_layers < 1 || d->n_layers > MAMDR_MAX_LAYERS) return 0;
    return star_ws(*d, max_batch).total;
}

// debug hook (tests / diagnostics): byte offsets of the activations inside the workspace for a batch of `rows`
extern "C" int mamdr_star_debug_offsets(const mamdr_star_desc* d, int32_t rows, int64_t* out /* [3 + 2 * n_layers + 1] */) {
    if (!d || !out) return MAMDR_E_INVALID;
    const StarWs w = star_ws(*d, rows);
    int k = 0;
    out[k++] = (int64_t)w.X; out[k++] = (int64_t)w.xhat; out[k++] = (int64_t)w.dY;
    for (int l = 0; l <= d->n_layers; ++l) out[k++] = (int64_t)w.H[l];
    for (int l = 0; l < d->n_layers; ++l) out[k++] = (int64_t)w.dZ[l];
    return MAMDR_OK;
}

extern "C" size_t mamdr_star_state_bytes(const mamdr_star_desc* d) {
    if (!d) return 0;
    const size_t dn = (size_t)d->n_domain * (d->emb_dim[0] + d->emb_dim[1] + d->emb_dim[2]);
    return 4 * dn * 4 + (size_t)d->n_domain * 4 + 64;
}

extern "C" int mamdr_star_eval_step(mamdr_ctx* ctx, const mamdr_star_desc* d, const mamdr_batch* b, const float* ut, const float* it,
                                    const float* params, void* pn_state, void* ws_, size_t ws_bytes, float* loss, float* probs,
                                    float* auc_acc, const float* thr, int32_t T, mamdr_stream stream) {
    int rc = star_validate(ctx, d, b, ws_, ws_bytes, ut, it, pn_state);
    if (rc) return rc;
    MAMDR_REQUIRE(ctx, params && aligned16(params), MAMDR_E_INVALID, "params NULL or misaligned");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* ws = (unsigned char*)ws_;
    const StarWs w = star_ws(*d, b->rows);
    rc = star_forward(ctx, d, b, ut, it, params, pn_state, ws, w, false, st);
    if (rc) return rc;
    return star_head(ctx, d, b, params, nullptr, ws, w, false, loss, probs, auc_acc, thr, T, st);
}

extern "C" int mamdr_star_train_step(mamdr_ctx* ctx, const mamdr_star_desc* d, const mamdr_batch* b, const float* ut, const float* it,
                                     const float* params, float* grads, void* pn_state, void* ws_, size_t ws_bytes, float* loss,
                                     float* probs, float* auc_acc, const float* thr, int32_t T, mamdr_stream stream) {
    int rc = star_validate(ctx, d, b, ws_, ws_bytes, ut, it, pn_state);
    if (rc) return rc;
    MAMDR_REQUIRE(ctx, params && grads && aligned16(params) && aligned16(grads), MAMDR_E_INVALID, "params/grads NULL or misaligned");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* ws = (unsigned char*)ws_;
    const StarWs w = star_ws(*d, b->rows);
    const int L = d->n_layers, rows = b->rows, dom = b->domain;
    const int n = d->emb_dim[0] + d->emb_dim[1] + d->emb_dim[2];
    // every slice that this batch does not touch (other domains' specific tensors, domain_emb) has a zero gradient
    MAMDR_CUDA_OK(ctx, cudaMemsetAsync(grads, 0, (size_t)d->arena_floats * 4, st));
    MAMDR_CUDA_OK(ctx, cudaMemsetAsync(ws + w.tickets, 0, (size_t)kMaxTiles * 4, st));
    rc = star_forward(ctx, d, b, ut, it, params, pn_state, ws, w, true, st);
    if (rc) return rc;
    rc = star_head(ctx, d, b, params, grads, ws, w, true, loss, probs, auc_acc, thr, T, st);
    if (rc) return rc;
    // ---- dZ_{l-1} = (dZ_l . W_l^T) * 1[H_l > 0] ; l = 0: dY = dZ_0 . W_0^T (gradient w.r.t. the PartitionedNorm output)
    for (int l = L - 1; l >= 0; --l) {
        const int Kd = d->hidden[l], Nd = l == 0 ? n : d->hidden[l - 1];
        simt::GemmShape s{rows, Nd, Kd, Kd, Kd};
        simt::LaunchPlan p = simt::plan(rows, Nd, Kd, 0, 1);
        if (l >= 1) {
            DhEpilogue epi{(const float*)(ws + w.H[l]), (float*)(ws + w.dZ[l - 1]), Nd, 1.0f};
            simt::gemm_kernel<true, false, DhEpilogue><<<p.grid, simt::THREADS, 0, st>>>((const float*)(ws + w.dZ[l]), (const float*)(ws + w.Weff[l]), s,
                                                                                         p.k_chunk, nullptr, nullptr, epi);
        } else {
            StoreEpilogue epi{(float*)(ws + w.dY), Nd};
            simt::gemm_kernel<true, false, StoreEpilogue><<<p.grid, simt::THREADS, 0, st>>>((const float*)(ws + w.dZ[0]), (const float*)(ws + w.Weff[0]), s,
                                                                                            p.k_chunk, nullptr, nullptr, epi);
        }
        MAMDR_LAUNCH_OK(ctx);
    }
    // ---- dW_eff_l = H_l^T . dZ_l (deterministic split-K) ; db_eff_l = column sums of dZ_l
    ColsumArgs ca;
    int maxn = 0;
    for (int l = 0; l < L; ++l) {
        const int Md = l == 0 ? n : d->hidden[l - 1], Nd = d->hidden[l];
        StoreEpilogue epi{(float*)(ws + w.dWeff[l]), Nd};
        simt::GemmShape s{Md, Nd, rows, Md, Nd};
        simt::LaunchPlan p = simt::plan(Md, Nd, rows, ctx->sm_count, kMaxSplit);
        MAMDR_REQUIRE(ctx, (int)(p.grid.x * p.grid.y) <= kMaxTiles, MAMDR_E_UNSUPPORTED, "layer too large for the ticket table");
        simt::gemm_kernel<false, true, StoreEpilogue><<<p.grid, simt::THREADS, 0, st>>>((const float*)(ws + w.H[l]), (const float*)(ws + w.dZ[l]), s,
                                                                                        p.k_chunk, (float*)(ws + w.partials),
                                                                                        (unsigned int*)(ws + w.tickets), epi);
        MAMDR_LAUNCH_OK(ctx);
        ca.job[l] = ColsumJob{(const float*)(ws + w.dZ[l]), (float*)(ws + w.dbeff[l]), Nd};
        if (Nd > maxn) maxn = Nd;
    }
    ca.rows = rows;
    colsum_kernel<<<dim3((maxn + 31) / 32, L), kColsumThreads, 0, st>>>(ca);
    MAMDR_LAUNCH_OK(ctx);
    StarGradArgs ga;
    ga.n_layers = L;
    int K = n;
    for (int l = 0; l < L; ++l) {
        const int N = d->hidden[l];
        ga.l[l] = StarGradPtrs{(const float*)(ws + w.dWeff[l]), (const float*)(ws + w.dbeff[l]), params + d->off_ksh[l],
                               params + d->off_ksp[l] + (int64_t)dom * K * N, grads + d->off_ksh[l], grads + d->off_ksp[l] + (int64_t)dom * K * N,
                               grads + d->off_bsh[l], grads + d->off_bsp[l] + (int64_t)dom * N, K * N, N};
        K = N;
    }
    star_grad_kernel<<<dim3(96, L), 256, 0, st>>>(ga);
    MAMDR_LAUNCH_OK(ctx);
    pn_backward_kernel<<<(n + 31) / 32, 256, 0, st>>>((const float*)(ws + w.dY), (const float*)(ws + w.xhat), rows, n, params + d->off_gamma_sh,
                                                    params + d->off_gamma_sp + (int64_t)dom * n, grads + d->off_gamma_sh,
                                                    grads + d->off_gamma_sp + (int64_t)dom * n, grads + d->off_beta_sh,
                                                    grads + d->off_beta_sp + (int64_t)dom * n);
    MAMDR_LAUNCH_OK(ctx);
    return MAMDR_OK;
}
