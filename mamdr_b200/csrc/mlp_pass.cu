// The persistent "pass kernel": ONE cooperative launch runs a whole domain pass -- `steps` consecutive
// mini-batches of forward, sigmoid-BCE head, backward and optimizer apply -- for the frozen-table mlp tower.
//
// Replaces `model.fit(train_iter, steps_per_epoch=S)` (/root/reference/model_zoo/mamdr.py:54) and the
// `for step in range(train_step): model.train_on_batch(train_iter)` loops (mamdr.py:85-97,
// model_zoo/domain_negotiation.py:71-72), i.e. S executions of the Keras train function built in
// model_zoo/DeepCTR/deepctr.py:54-60,118-136.  With train = 0 it is `model.evaluate(dataset, steps)`
// (model_zoo/specific_base_model.py:82-85, model_zoo/base_model.py:130-133).
//
// Why one kernel: at batch 1024 a mini-batch is ~0.7 GFLOP over ~2 MB of L2-resident operands, i.e. ~1 us at
// the B200 rooflines; a stream of per-layer kernels is bound by launch + pipeline-fill latency (measured 146 us
// per mini-batch for 14 launches, profiles/r1_v2_*).  Here every SM keeps its barriers, TMEM allocation and
// pipeline alive for the whole pass and the layers are separated by grid barriers (~1.2 us) instead of launches.
//
// CTA = 10 warps: warps 0-7 = epilogue / element-wise workers (warp w <-> TMEM lanes 32 (w % 4) .. +31 <-> tile rows,
// column half w / 4 of the tile), warp 8 = TMA producer, warp 9 = tcgen05.mma issuer (owns the TMEM allocation).
// Both issue warps run their loops CONVERGED, one elected lane issuing: TMA and MMA descriptors then live in uniform
// registers (under `if (lane == 0)` every instruction pays an R2UR waterfall: 130-200 cycles per MMA instead of 50,
// profiles/r2_probe_mainloop.txt).
//
// 3xTF32 (MAMDR_PREC_TF32X3): every GEMM operand is kept in global memory as a PAIR array [2][rows][cols]: plane 0 =
// the fp32 value (the tensor core truncates it to its tf32 "hi" part), plane 1 = lo = rn_tf32(x - hi), written by the
// producing epilogue / gather / optimizer apply.  TMA stages [A | A_lo | B | B_lo]; B and B_lo are adjacent and form
// ONE operand of N = 2 bn, so a k-step is two MMAs: A.[B | B_lo] (two accumulator halves) and A_lo.B (first half);
// the epilogue adds the halves.  Dropped: A_lo.B_lo (2^-22 relative).
//
// Per mini-batch (L hidden layers) the grid walks 2L+1 phases, each a list of independent tile jobs
// (job j runs on CTA j mod grid):
//   fwd l < L-1 : H_{l+1} = dropout(relu(H_l . W_l + b_l))         tiles 128 x 32        (l = 0: + E_d[dom] . W_0dom)
//   fwd L-1     : last hidden layer + Dense(1) + sigmoid + BCE + dZ_{L-1} + AUC bins, tiles 128 x n_L; the other
//                 CTAs gather the NEXT mini-batch's embedding rows (frozen tables: no dependence on the update)
//   bwd l>=1    : dZ_{l-1} = (dZ_l . W_l^T) * mask(H_l) (+ db_{l-1} partials)  and  split-K partials of dW_l
//   bwd l = 0   : split-K partials of dW_0, and the domain-embedding job (db_0, dE_d[dom], |E_d|^2)
//   update      : fixed-order reduction of the partials fused with the Adam / SGD apply on every parameter
// The domain embedding row is the same for every sample of a batch (utils/dataset.py:73-99: per-domain
// datasets), so X is only [E_u | E_i] (K = 256) and the domain block of layer 0 is folded into its bias in fp32
// (SURVEY.md A-10); its weight gradient is the rank-1 product E_d[dom]^T (x) db_0.
// No float atomics anywhere: results are bit-reproducible run to run and rank to rank.
#include <vector>

#include "common.cuh"
#include "meta_ops.cuh"
#include "philox.cuh"
#include "program.cuh"
#include "tc_tmap.cuh"

namespace passk {

#define WSTAMP(k) do { if (tim && tid == 0) a.timing[tslot + (k)] = (unsigned long long)clock64(); } while (0)

constexpr int kWorkerWarps = 8;
constexpr int kWorkers = 32 * kWorkerWarps;   // 256
constexpr int kProdWarp = kWorkerWarps;       // warp 8
constexpr int kMmaWarp = kWorkerWarps + 1;    // warp 9
constexpr int kThreads = 32 * (kWorkerWarps + 2);
constexpr int KCH = 32;                       // floats per K chunk = one 128-byte swizzle row
constexpr int A_BYTES = 128 * KCH * 4;        // 16 KB: one A tile
constexpr int B_BYTES = 64 * KCH * 4;         // 8 KB: one B tile at BN = 64
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;   // [A | A_lo | B | B_lo] = 48 KB
constexpr int kStages = 4;
constexpr int kScratchBytes = 20 * 1024;
constexpr int kMaxThr = 1024;
constexpr int kMaxSplit = 8;
constexpr int kDomJobs = 8;                   // the domain-embedding gradient GEMV is split over this many CTAs
constexpr int kMaxMT = 64;                    // 128-row tiles per mini-batch (batch <= 8192)
constexpr uint32_t kTmemCols = 128;           // one accumulator of up to 2 x 64 columns

enum { J_NONE = 0, J_FWD, J_HEAD, J_DH, J_DW, J_DOM };
enum { SEG_ED = 0, SEG_KERNEL, SEG_BIAS, SEG_DENSE, SEG_GBIAS };

struct MapTable {   // kernel parameter (param space is a legal tensor-map address space); every map is a pair map
    CUtensorMap xk[2], xmn[2];                  // X double buffer: K-major [B, K0] / MN-major view
    CUtensorMap hk[MAMDR_MAX_LAYERS];           // H_l  K-major  (A of fwd l),      l = 1..L-1
    CUtensorMap hmn[MAMDR_MAX_LAYERS];          // H_l  MN-major (A of dW_l)
    CUtensorMap dzk[MAMDR_MAX_LAYERS];          // dZ_l K-major  (A of dH_l),       l = 1..L-1
    CUtensorMap dzmn[MAMDR_MAX_LAYERS];         // dZ_l MN-major (B of dW_l),       l = 0..L-1
    CUtensorMap wf[MAMDR_MAX_LAYERS];           // W_l  MN-major [K, N] (B of fwd l)
    CUtensorMap wb[MAMDR_MAX_LAYERS];           // W_l  K-major  [N = in, K = out] (B of dH_l), l = 1..L-1
};

struct Seg { long long off; int numel; int kind; int layer; };

struct PassDyn {   // the per-pass fields, loaded from the current ProgOp
    int dom, steps;
    long long n_data;
    const int32_t *uid, *pid, *order;
    const float* label;
    float *losses, *probs;
};

struct PassArgs {
    // ---- program: ops == NULL -> the single inline op (one pass per launch)
    const ProgOp* ops;
    int n_ops;
    ProgOp inline_op;
    // ---- model
    int L, n[MAMDR_MAX_LAYERS + 1];   // n[0] = K0 = du + di (the domain block is folded), n[l+1] = hidden[l]
    int du, di, dd, n_domain;
    long long off_Ed, off_W[MAMDR_MAX_LAYERS], off_b[MAMDR_MAX_LAYERS], off_w, off_g, arena;
    int nseg;
    Seg seg[2 * MAMDR_MAX_LAYERS + 3];
    float *params, *m, *v, *grads;    // grads may be NULL
    float* wpair;                     // arena-indexed pair shadow of the kernels: plane 0 at wpair, plane 1 at wpair + wz
    long long wz;
    const float *Eu, *Ei;
    // ---- data
    int bs, max_rows;
    // ---- workspace (pair arrays: the lo plane lies max_rows * width floats behind the hi plane)
    float *X[2], *y[2], *H[MAMDR_MAX_LAYERS], *dZ[MAMDR_MAX_LAYERS], *partials[MAMDR_MAX_LAYERS], *db_part[MAMDR_MAX_LAYERS];
    float *dw_part, *dg_part, *db0_red, *gEd_row, *ed_row;
    double *loss_part, *ed_sq;
    int* hist;
    unsigned int* bar;
    // ---- optimizer / loss
    OptState* state;
    int opt_kind;   // 0 = Adam, 1 = SGD
    float lr, beta1, beta2, eps;
    int dropout_enabled;
    uint32_t dropout_seed, dropout_threshold;
    float dropout_scale, l2_emb, frozen_reg;
    float* auc_acc;
    const float* thr;
    int T, train, passes;
    unsigned long long* timing;   // debug: [step][phase][cta][16] time stamps (NULL in production)
    long long timing_cap;
};

struct Job {
    int type, layer, m_tile, n_tile, z, bn, nch, c_beg, tiles, NT;
};

__host__ __device__ inline int cdiv(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline int dw_bn(const PassArgs& a, int l) { return a.n[l + 1] < 64 ? a.n[l + 1] : 64; }
__host__ __device__ inline void split_plan(int rows, int& chunks, int& S, int& cps) {
    chunks = cdiv(rows, KCH);
    S = chunks < kMaxSplit ? chunks : kMaxSplit;
    cps = cdiv(chunks, S);
    S = cdiv(chunks, cps);
}
__host__ __device__ inline int dw_tiles(const PassArgs& a, int l) { return cdiv(a.n[l], 128) * (a.n[l + 1] / dw_bn(a, l)); }

__host__ __device__ inline int phase_jobs(const PassArgs& a, int phase, int rows) {
    const int L = a.L, mt = cdiv(rows, 128);
    if (phase < L) return phase < L - 1 ? mt * (a.n[phase + 1] / 32) : mt;
    const int l = L - 1 - (phase - L);
    int chunks, S, cps;
    split_plan(rows, chunks, S, cps);
    return (l >= 1 ? mt * (a.n[l] / 32) : kDomJobs) + dw_tiles(a, l) * S;
}

__device__ __forceinline__ Job decode_job(const PassArgs& a, int phase, int rows, int j) {
    Job J;
    J.z = 0; J.c_beg = 0; J.tiles = 0; J.NT = 1; J.n_tile = 0;
    const int L = a.L;
    if (phase < L) {
        const int l = phase;
        J.layer = l;
        J.nch = a.n[l] / KCH;
        if (l < L - 1) {
            const int nt = a.n[l + 1] / 32;
            J.type = J_FWD; J.m_tile = j / nt; J.n_tile = j - J.m_tile * nt; J.bn = 32;
        } else {
            J.type = J_HEAD; J.m_tile = j; J.bn = a.n[L];
        }
        return J;
    }
    const int l = L - 1 - (phase - L);
    J.layer = l;
    const int mt = cdiv(rows, 128);
    const int nlead = l >= 1 ? mt * (a.n[l] / 32) : kDomJobs;
    if (j < nlead) {
        if (l >= 1) {
            const int nt = a.n[l] / 32;
            J.type = J_DH; J.m_tile = j / nt; J.n_tile = j - J.m_tile * nt; J.bn = 32; J.nch = a.n[l + 1] / KCH;
        } else {
            J.type = J_DOM; J.m_tile = j; J.bn = 0; J.nch = 0;
        }
        return J;
    }
    const int jj = j - nlead;
    int chunks, S, cps;
    split_plan(rows, chunks, S, cps);
    J.type = J_DW;
    J.bn = dw_bn(a, l);
    J.NT = a.n[l + 1] / J.bn;
    J.tiles = cdiv(a.n[l], 128) * J.NT;
    J.z = jj / J.tiles;
    const int t = jj - J.z * J.tiles;
    J.m_tile = t / J.NT;
    J.n_tile = t - J.m_tile * J.NT;
    J.c_beg = J.z * cps;
    const int c_end = chunks < J.c_beg + cps ? chunks : J.c_beg + cps;
    J.nch = c_end - J.c_beg;
    return J;
}

// ---- small device helpers ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void worker_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ float ldcg_f(const float* p) { return __ldcg(p); }
__device__ __forceinline__ float4 ldcg_f4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add_u32(unsigned int* p, unsigned int v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// grid-wide barrier over a monotonic counter (zeroed by the host before the launch; cooperative launch
// guarantees co-residency).  Generic-proxy writes made before the barrier are read after it through TMA (async
// proxy) by other SMs, hence the proxy fences on both sides.
__device__ __forceinline__ void grid_barrier(unsigned int* ctr, unsigned int& target) {
    fence_proxy_async_all();     // this thread's generic-proxy writes -> visible to async-proxy (TMA) readers
    __syncthreads();
    target += gridDim.x;
    if (threadIdx.x == 0) {
        // release is cumulative: it covers the writes of every thread of the CTA ordered before it by bar.sync
        red_release_add_u32(ctr, 1u);
        while (ld_acquire_u32(ctr) < target) {}
        fence_proxy_async_all();
    }
    __syncthreads();
}

// round-to-nearest fp32 -> tf32 (low 13 mantissa bits cleared).  The tensor core TRUNCATES raw fp32 operands, a
// biased error that does not cancel in long sums; the 1-pass TF32 mode therefore feeds it pre-rounded operands.
__device__ __forceinline__ float rn_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ float4 rn_tf32_4(float4 v) { return make_float4(rn_tf32(v.x), rn_tf32(v.y), rn_tf32(v.z), rn_tf32(v.w)); }
__device__ __forceinline__ float4 tf32_lo_4(float4 v) { return make_float4(tc::tf32_lo(v.x), tc::tf32_lo(v.y), tc::tf32_lo(v.z), tc::tf32_lo(v.w)); }

// store 4 consecutive elements of a pair array: plane 0 = v (pre-rounded in the 1-pass mode), plane 1 = lo (3-pass mode)
__device__ __forceinline__ void store_pair4(float* hi, long long z, float4 v, bool rnd, bool x3) {
    *reinterpret_cast<float4*>(hi) = rnd ? rn_tf32_4(v) : v;
    if (x3) *reinterpret_cast<float4*>(hi + z) = tf32_lo_4(v);
}

__device__ __forceinline__ void adam1(float& p, float& m, float& v, float g, float alpha, float omb1, float omb2, float eps) {
    m = __fadd_rn(m, __fmul_rn(__fsub_rn(g, m), omb1));
    v = __fadd_rn(v, __fmul_rn(__fsub_rn(__fmul_rn(g, g), v), omb2));
    p = __fsub_rn(p, __fdiv_rn(__fmul_rn(m, alpha), __fadd_rn(__fsqrt_rn(v), eps)));
}

// Column sums over the 32 lanes of a warp (lane = tile row) of C per-lane values, in a fixed butterfly order:
// afterwards lane l holds the total of column (C == 32 ? l : l >> 1).  31 (C = 32) / 16 (C = 16) shuffles.
template <int C>
__device__ __forceinline__ float warp_colsum(float (&v)[C], int lane) {
    static_assert(C == 16 || C == 32, "16 or 32 columns per lane");
#pragma unroll
    for (int step = 0; step < (C == 32 ? 5 : 4); ++step) {
        const int off = 16 >> step, half = (C / 2) >> step;
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float send = up ? v[i] : v[i + half];
            const float recv = __shfl_xor_sync(0xffffffffu, send, off);
            v[i] = (up ? v[i + half] : v[i]) + recv;
        }
    }
    float r = v[0];
    if (C == 16) r += __shfl_xor_sync(0xffffffffu, r, 1);
    return r;
}

// gather the rows of mini-batch `step` into the pair buffer X[buf] / y[buf]; one warp per row, 16-byte lanes
__device__ __forceinline__ void gather_rows(const PassArgs& a, const PassDyn& pd, int step, int buf, int warp_rank, int n_warps, int lane, bool rnd, bool x3) {
    const long long off = (long long)step * a.bs;
    const long long left = pd.n_data - off;
    const int rows = left < a.bs ? (int)left : a.bs;
    const int K0 = a.du + a.di;
    const long long xz = (long long)a.max_rows * K0;
    float* X = a.X[buf];
    float* y = a.y[buf];
    for (int r = warp_rank; r < rows; r += n_warps) {
        const long long o = pd.order ? (long long)__ldg(pd.order + off + r) : off + r;
        const long long u = __ldg(pd.uid + o), p = __ldg(pd.pid + o);
        const float* su = a.Eu + u * a.du;
        const float* si = a.Ei + p * a.di;
        float* xr = X + (long long)r * K0;
        for (int c = lane * 4; c < a.du; c += 128) store_pair4(xr + c, xz, ldg_f4(su + c), rnd, x3);
        for (int c = lane * 4; c < a.di; c += 128) store_pair4(xr + a.du + c, xz, ldg_f4(si + c), rnd, x3);
        if (lane == 0) y[r] = __ldg(pd.label + o);
    }
}

// ---- the kernel ---------------------------------------------------------------------------------------------------
template <int NL>
__global__ void __launch_bounds__(kThreads, 1)
pass_kernel(const __grid_constant__ MapTable maps, const __grid_constant__ PassArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar_full[kStages], bar_empty[kStages], bar_done, bar_tfree;
    __shared__ uint32_t tmem_base_s;
    __shared__ float s_thr[kMaxThr];
    __shared__ float s_beff[64];
    __shared__ float s_wd[64];
    __shared__ float s_z[2][128];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = gridDim.x, cta = blockIdx.x;
    const int passes = a.passes;
    const bool x3 = passes == 3;    // 3xTF32: pair operands, N-concatenated MMAs
    const bool rnd = passes == 1;   // 1-pass TF32: GEMM operands are stored pre-rounded (RN) to tf32
    unsigned char* scratch = smem + (size_t)kStages * STAGE_BYTES;

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            tc::mbar_init(&bar_full[s], 1);
            tc::mbar_init(&bar_empty[s], 1);
        }
        tc::mbar_init(&bar_done, 1);
        tc::mbar_init(&bar_tfree, kWorkers);
        tc::fence_barrier_init();
    }
    if (warp == kMmaWarp) tc::tmem_alloc(&tmem_base_s, kTmemCols);
    if (warp == kProdWarp && lane == 0) {
        for (int b = 0; b < 2; ++b) { tc::tma_prefetch_desc(&maps.xk[b]); tc::tma_prefetch_desc(&maps.xmn[b]); }
        for (int l = 0; l < a.L; ++l) {
            tc::tma_prefetch_desc(&maps.wf[l]);
            tc::tma_prefetch_desc(&maps.dzmn[l]);
            if (l >= 1) { tc::tma_prefetch_desc(&maps.hk[l]); tc::tma_prefetch_desc(&maps.hmn[l]); tc::tma_prefetch_desc(&maps.dzk[l]); tc::tma_prefetch_desc(&maps.wb[l]); }
        }
    }
    for (int i = tid; i < a.T && i < kMaxThr; i += kThreads) s_thr[i] = a.thr ? a.thr[i] : 0.f;
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const uint32_t smem_base = tc::smem_u32(smem);

    // replicated optimizer scalars: every thread advances its own copy with the same fp32 operations
    long long step_ctr = a.state->step;
    float b1pow = a.state->b1pow, b2pow = a.state->b2pow;
    unsigned int bar_target = 0;
    int ring_s = 0;            // pipeline stage cursor of this warp's role (producer / MMA issuer)
    uint32_t ring_ph = 0;      // phase bit of the current lap
    uint32_t ring_n = 0;       // chunks handled so far (the first kStages need no empty-wait)
    uint32_t njob = 0;         // GEMM jobs run so far by this CTA (parity of bar_done / bar_tfree)
    const int K0 = a.n[0];
    const int L = a.L;
    const float inv_keep = a.dropout_enabled && a.train ? a.dropout_scale : 1.0f;
    const int nz = x3 ? 2 : 1;

    // refresh the pair shadow of the kernels from the parameter arena (float4 items of this thread)
    auto refresh_wpair = [&]() {
        for (int q = 0; q < a.nseg; ++q) {
            if (a.seg[q].kind != SEG_KERNEL) continue;
            const long long o0 = a.seg[q].off;
            for (int i = (cta * kThreads + tid) * 4; i < a.seg[q].numel; i += G * kThreads * 4)
                store_pair4(a.wpair + o0 + i, a.wz, ldcg_f4(a.params + o0 + i), rnd, x3);
        }
    };

    const ProgOp* ops = a.ops ? a.ops : &a.inline_op;
    int pass_idx = 0;
    for (int oi = 0; oi < a.n_ops; ++oi) {
    const ProgOp& op = ops[oi];
    if (op.kind == PROG_META) {
        // ---------- element-wise DN / DR meta sweep over arenas (K9 / K10), all CTAs
        MetaArgs ma;
        ma.w0 = op.w0; ma.w1 = op.w1; ma.r0 = op.r0; ma.r1 = op.r1; ma.r2 = op.r2;
        ma.f0 = op.f0; ma.f1 = op.f1; ma.method = op.method >> 8; ma.n = op.n;
        const int mop = op.method & 0xff;
        const int64_t nv4 = op.n >> 2;
        for (int64_t i = (int64_t)cta * kThreads + tid; i < nv4; i += (int64_t)G * kThreads) meta_float4(mop, ma, i);
        grid_barrier(a.bar, bar_target);
        continue;
    }
    PassDyn pd;
    pd.dom = op.domain; pd.steps = op.steps; pd.n_data = op.n_data;
    pd.uid = op.uid; pd.pid = op.pid; pd.order = op.order; pd.label = op.label;
    pd.losses = op.losses; pd.probs = op.probs;
    int* const hist_cur = a.hist + (pass_idx & 1) * (2 * (kMaxThr + 1));

    // ---- prologue of a pass: gather mini-batch 0; pair shadow of the kernels (the arena may have been rewritten by a
    // meta sweep or by the host since the last pass); |E_d|^2 for the inference loss
    refresh_wpair();
    if (warp < kWorkerWarps) {
        gather_rows(a, pd, 0, 0, cta * kWorkerWarps + warp, G * kWorkerWarps, lane, rnd, x3);
        if (!a.train && cta == G - 1) {
            double sq = 0.0;
            for (int i = tid; i < a.n_domain * a.dd; i += kWorkers) { const double e = ldcg_f(a.params + a.off_Ed + i); sq += e * e; }
            double* red = reinterpret_cast<double*>(scratch);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
            if (lane == 0) red[warp] = sq;
            worker_sync();
            if (tid == 0) {
                double tot = 0.0;
                for (int w = 0; w < kWorkerWarps; ++w) tot += red[w];
                *a.ed_sq = tot;
            }
        }
    }
    grid_barrier(a.bar, bar_target);

    const int n_phases = a.train ? 2 * L + 1 : L;
    for (int step = 0; step < pd.steps; ++step) {
        const long long left = pd.n_data - (long long)step * a.bs;
        const int rows = left < a.bs ? (int)left : a.bs;
        const int mt = cdiv(rows, 128);
        const int buf = step & 1;
        DropoutParams dp;
        dp.enabled = a.dropout_enabled && a.train;
        dp.threshold = a.dropout_threshold;
        dp.scale = a.dropout_scale;
        dp.step = (uint32_t)(step_ctr & 0xffffffffll);
        dp.seed = a.dropout_seed;

        const int jobs_last_bwd = a.train ? phase_jobs(a, 2 * L - 1, rows) : 0;
        const bool early_done = a.train && G - jobs_last_bwd >= G / 4;   // enough idle CTAs in the last backward phase
        // fixed-order reduction of the split-K / per-tile partials fused with the optimizer apply, for the float4 items
        // first, first + stride, ...;  which: 0 = all, 1 = early items only, 2 = late items only
        auto update_items = [&](long long first, long long stride, int which) {
            int chunks, S, cps;
            split_plan(rows, chunks, S, cps);
            const float alpha = __fdiv_rn(__fmul_rn(a.lr, __fsqrt_rn(__fsub_rn(1.0f, b2pow))), __fsub_rn(1.0f, b1pow));
            const float omb1 = __fsub_rn(1.0f, a.beta1), omb2 = __fsub_rn(1.0f, a.beta2);
            const float two_l2 = 2.0f * a.l2_emb;
            const long long nv4 = a.arena >> 2;
            for (long long i4 = first; i4 < nv4; i4 += stride) {
                const long long o = i4 << 2;
                int si = -1;
                for (int q = 0; q < a.nseg; ++q)
                    if (o >= a.seg[q].off && o < a.seg[q].off + a.seg[q].numel) si = q;
                if (si < 0) continue;   // alignment padding stays zero
                const Seg sg = a.seg[si];
                // early items: everything whose gradient is final before the last backward phase (layers >= 1, dense, global bias)
                const bool late = sg.kind == SEG_ED || ((sg.kind == SEG_KERNEL || sg.kind == SEG_BIAS) && sg.layer == 0);
                if ((which == 1 && late) || (which == 2 && !late)) continue;
                const int e = (int)(o - sg.off);
                const float4 P = ldcg_f4(a.params + o);
                float4 M = make_float4(0.f, 0.f, 0.f, 0.f), V = M;
                if (a.opt_kind == 0) { M = ldcg_f4(a.m + o); V = ldcg_f4(a.v + o); }
                float g[4] = {0.f, 0.f, 0.f, 0.f};
                if (sg.kind == SEG_KERNEL) {
                    const int l = sg.layer, N = a.n[l + 1];
                    const int k = e / N, c = e - k * N;
                    if (l == 0 && k >= K0) {   // domain block: rank-1  E_d[dom]^T (x) db_0
                        const float ev = ldcg_f(a.ed_row + (k - K0));
                        const float4 d4 = ldcg_f4(a.db0_red + c);
                        g[0] = __fmul_rn(ev, d4.x); g[1] = __fmul_rn(ev, d4.y); g[2] = __fmul_rn(ev, d4.z); g[3] = __fmul_rn(ev, d4.w);
                    } else {
                        const int bn = dw_bn(a, l), NT = N / bn, tiles = cdiv(a.n[l], 128) * NT;
                        const int tile = (k >> 7) * NT + c / bn;
                        const float* src = a.partials[l] + ((long long)tile * 128 + (k & 127)) * bn + (c % bn);
                        const long long zstride = (long long)tiles * 128 * bn;
                        float4 q4[kMaxSplit];
#pragma unroll
                        for (int z = 0; z < kMaxSplit; ++z) q4[z] = z < S ? ldcg_f4(src + z * zstride) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int z = 0; z < kMaxSplit; ++z)
                            if (z < S) { g[0] += q4[z].x; g[1] += q4[z].y; g[2] += q4[z].z; g[3] += q4[z].w; }
                    }
                } else if (sg.kind == SEG_BIAS) {
                    const int N = a.n[sg.layer + 1];
                    const float* src = sg.layer == 0 ? nullptr : a.db_part[sg.layer] + e;
                    if (sg.layer == 0) {
                        const float4 q4 = ldcg_f4(a.db0_red + e);
                        g[0] = q4.x; g[1] = q4.y; g[2] = q4.z; g[3] = q4.w;
                    } else {
                        for (int m0 = 0; m0 < mt; m0 += 8) {
                            float4 q4[8];
#pragma unroll
                            for (int u = 0; u < 8; ++u) q4[u] = m0 + u < mt ? ldcg_f4(src + (long long)(m0 + u) * N) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                            for (int u = 0; u < 8; ++u)
                                if (m0 + u < mt) { g[0] += q4[u].x; g[1] += q4[u].y; g[2] += q4[u].z; g[3] += q4[u].w; }
                        }
                    }
                } else if (sg.kind == SEG_DENSE) {
                    for (int m0 = 0; m0 < mt; m0 += 8) {
                        float4 q4[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) q4[u] = m0 + u < mt ? ldcg_f4(a.dw_part + (m0 + u) * NL + e) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            if (m0 + u < mt) { g[0] += q4[u].x; g[1] += q4[u].y; g[2] += q4[u].z; g[3] += q4[u].w; }
                    }
                } else if (sg.kind == SEG_GBIAS) {
                    for (int m = 0; m < mt; ++m) g[0] += ldcg_f(a.dg_part + m);
                } else {   // SEG_ED: L2 term on every row (+ the batch row's data gradient)
                    const float4 p4 = P;
                    g[0] = __fmul_rn(two_l2, p4.x); g[1] = __fmul_rn(two_l2, p4.y); g[2] = __fmul_rn(two_l2, p4.z); g[3] = __fmul_rn(two_l2, p4.w);
                    if (e / a.dd == pd.dom) {
                        const float4 d4 = ldcg_f4(a.gEd_row + (e - pd.dom * a.dd));
                        g[0] = __fadd_rn(g[0], d4.x); g[1] = __fadd_rn(g[1], d4.y); g[2] = __fadd_rn(g[2], d4.z); g[3] = __fadd_rn(g[3], d4.w);
                    }
                }
                float pp[4] = {P.x, P.y, P.z, P.w};
                if (a.opt_kind == 0) {
                    float mm[4] = {M.x, M.y, M.z, M.w}, vv[4] = {V.x, V.y, V.z, V.w};
#pragma unroll
                    for (int t = 0; t < 4; ++t) adam1(pp[t], mm[t], vv[t], g[t], alpha, omb1, omb2, a.eps);
                    *reinterpret_cast<float4*>(a.m + o) = make_float4(mm[0], mm[1], mm[2], mm[3]);
                    *reinterpret_cast<float4*>(a.v + o) = make_float4(vv[0], vv[1], vv[2], vv[3]);
                } else {
#pragma unroll
                    for (int t = 0; t < 4; ++t) pp[t] = __fsub_rn(pp[t], __fmul_rn(g[t], a.lr));
                }
                const float4 pnew = make_float4(pp[0], pp[1], pp[2], pp[3]);
                *reinterpret_cast<float4*>(a.params + o) = pnew;
                if (sg.kind == SEG_KERNEL) store_pair4(a.wpair + o, a.wz, pnew, rnd, x3);
                if (a.grads) *reinterpret_cast<float4*>(a.grads + o) = make_float4(g[0], g[1], g[2], g[3]);
            }
        };
        for (int phase = 0; phase < n_phases; ++phase) {
            const long long tslot = (((long long)step * n_phases + phase) * G + cta) * 16;
            const bool tim = a.timing && tslot + 15 < a.timing_cap;
            if (tim && tid == 0) { a.timing[tslot] = gtime(); a.timing[tslot + 7] = (unsigned long long)clock64(); }
            if (phase < 2 * L) {
                const int njobs = phase_jobs(a, phase, rows);
                for (int j = cta; j < njobs; j += G) {
                    const Job J = decode_job(a, phase, rows, j);
                    const int l = J.layer;
                    if (J.type == J_DOM) {
                        // ---------- domain-embedding job q of kDomJobs (workers): db_0 (every job, into smem), rows
                        // [q*per, (q+1)*per) of dE_d[dom] = W_0dom . db_0; job 0 also publishes db_0, E_d[dom], |E_d|^2
                        if (warp < kWorkerWarps) {
                            const int n1 = a.n[1];
                            float* s_db0 = reinterpret_cast<float*>(scratch);           // [n1] (n1 <= 4096: 16 KB)
                            double* red = reinterpret_cast<double*>(scratch + 16384);
                            for (int c = tid * 4; c < n1; c += kWorkers * 4) {
                                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                                float4 part[8];
                                for (int m0 = 0; m0 < mt; m0 += 8) {
#pragma unroll
                                    for (int u = 0; u < 8; ++u)
                                        part[u] = m0 + u < mt ? ldcg_f4(a.db_part[0] + (long long)(m0 + u) * n1 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                                    for (int u = 0; u < 8; ++u)
                                        if (m0 + u < mt) { acc.x += part[u].x; acc.y += part[u].y; acc.z += part[u].z; acc.w += part[u].w; }
                                }
                                *reinterpret_cast<float4*>(s_db0 + c) = acc;
                                if (J.m_tile == 0) *reinterpret_cast<float4*>(a.db0_red + c) = acc;
                            }
                            if (J.m_tile == 0) {
                                double sq = 0.0;
                                for (int i = tid; i < a.n_domain * a.dd; i += kWorkers) { const double e = ldcg_f(a.params + a.off_Ed + i); sq += e * e; }
#pragma unroll
                                for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
                                if (lane == 0) red[warp] = sq;
                                for (int k = tid; k < a.dd; k += kWorkers) a.ed_row[k] = ldcg_f(a.params + a.off_Ed + (long long)pd.dom * a.dd + k);
                            }
                            worker_sync();
                            if (J.m_tile == 0 && tid == 0) {
                                double tot = 0.0;
                                for (int w = 0; w < kWorkerWarps; ++w) tot += red[w];
                                *a.ed_sq = tot;
                            }
                            const float* W0dom = a.params + a.off_W[0] + (long long)K0 * n1;
                            const int per = cdiv(a.dd, kDomJobs);
                            const int k_end = (J.m_tile + 1) * per < a.dd ? (J.m_tile + 1) * per : a.dd;
                            for (int k = J.m_tile * per + warp; k < k_end; k += kWorkerWarps) {
                                const float* wr = W0dom + (long long)k * n1;
                                float s = 0.f;
                                for (int c0 = 0; c0 < n1; c0 += 512) {   // 4 x 128 floats per trip, loads first
                                    float4 wv[4];
#pragma unroll
                                    for (int u = 0; u < 4; ++u) {
                                        const int c = c0 + u * 128 + lane * 4;
                                        wv[u] = c < n1 ? ldcg_f4(wr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                                    }
#pragma unroll
                                    for (int u = 0; u < 4; ++u) {
                                        const int c = c0 + u * 128 + lane * 4;
                                        if (c < n1) {
                                            const float4 d4 = *reinterpret_cast<const float4*>(s_db0 + c);
                                            s = fmaf(d4.x, wv[u].x, s); s = fmaf(d4.y, wv[u].y, s); s = fmaf(d4.z, wv[u].z, s); s = fmaf(d4.w, wv[u].w, s);
                                        }
                                    }
                                }
#pragma unroll
                                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                                if (lane == 0) a.gEd_row[k] = s;
                            }
                            worker_sync();
                        }
                        continue;
                    }
                    const bool a_mn = J.type == J_DW, b_mn = J.type != J_DH;
                    if (warp == kProdWarp) {
                        // ---------- TMA producer (whole warp converged, one elected lane issues)
                        const CUtensorMap* ma;
                        const CUtensorMap* mb;
                        if (J.type == J_FWD || J.type == J_HEAD) { ma = l == 0 ? &maps.xk[buf] : &maps.hk[l]; mb = &maps.wf[l]; }
                        else if (J.type == J_DH) { ma = &maps.dzk[l]; mb = &maps.wb[l]; }
                        else { ma = l == 0 ? &maps.xmn[buf] : &maps.hmn[l]; mb = &maps.dzmn[l]; }
                        const uint32_t b_tile = (uint32_t)J.bn * KCH * 4;
                        const uint32_t tx = (uint32_t)(A_BYTES + b_tile) * (uint32_t)nz;
                        const int a_row = J.m_tile * 128, b_col = J.n_tile * J.bn, ngb = J.bn >> 5;
                        for (int i = 0; i < J.nch; ++i) {
                            const int kc = (J.c_beg + i) * KCH;
                            if (ring_n >= (uint32_t)kStages) tc::mbar_wait(&bar_empty[ring_s], ring_ph ^ 1);
                            unsigned char* sA = smem + (size_t)ring_s * STAGE_BYTES;
                            unsigned char* sB = sA + 2 * A_BYTES;
                            uint64_t* fb = &bar_full[ring_s];
                            if (tc::elect_one()) {
                                tc::mbar_arrive_expect_tx(fb, tx);
                                for (int z = 0; z < nz; ++z) {
                                    if (a_mn) {
#pragma unroll
                                        for (int g = 0; g < 4; ++g) tc::tma_load_3d(sA + z * A_BYTES + g * 4096, ma, fb, a_row + g * 32, kc, z);
                                    } else {
                                        tc::tma_load_3d(sA + z * A_BYTES, ma, fb, kc, a_row, z);
                                    }
                                    if (b_mn) {
                                        for (int g = 0; g < ngb; ++g) tc::tma_load_3d(sB + z * b_tile + g * 4096, mb, fb, b_col + g * 32, kc, z);
                                    } else {
                                        tc::tma_load_3d(sB + z * b_tile, mb, fb, kc, b_col, z);
                                    }
                                }
                                if (tim) {
                                    if (i == 0) a.timing[tslot + 2] = (unsigned long long)clock64();
                                    if (i == J.nch - 1) a.timing[tslot + 3] = (unsigned long long)clock64();
                                }
                            }
                            __syncwarp();
                            ++ring_n;
                            if (++ring_s == kStages) { ring_s = 0; ring_ph ^= 1; }
                        }
                    } else if (warp == kMmaWarp) {
                        // ---------- MMA issuer (whole warp converged, one elected lane issues)
                        if (njob > 0) tc::mbar_wait(&bar_tfree, (njob - 1) & 1);
                        tc::tc_fence_after();
                        const uint32_t idesc1 = tc::make_idesc_tf32(128, J.bn, a_mn ? 1 : 0, b_mn ? 1 : 0);
                        const uint32_t idesc2 = tc::make_idesc_tf32(128, 2 * J.bn, a_mn ? 1 : 0, b_mn ? 1 : 0);
                        const uint32_t a_hiw = a_mn ? tc::kDescHiMN : tc::kDescHiK, a_low = a_mn ? tc::kDescLoMN : tc::kDescLoK;
                        const uint32_t b_hiw = b_mn ? tc::kDescHiMN : tc::kDescHiK, b_low = b_mn ? tc::kDescLoMN : tc::kDescLoK;
                        const uint32_t a_k = a_mn ? (1024u >> 4) : (32u >> 4), b_k = b_mn ? (1024u >> 4) : (32u >> 4);   // per k-step of 8
                        uint32_t acc = 0;
                        for (int i = 0; i < J.nch; ++i) {
                            tc::mbar_wait(&bar_full[ring_s], ring_ph);
                            tc::tc_fence_after();
                            const uint32_t st = (smem_base + (uint32_t)ring_s * STAGE_BYTES) >> 4;
                            if (tc::elect_one()) {
                                if (tim && i == 0) a.timing[tslot + 4] = (unsigned long long)clock64();
                                const uint32_t aw = st | a_low, alw = aw + (A_BYTES >> 4);
                                const uint32_t bw = (st + ((2 * A_BYTES) >> 4)) | b_low;
                                if (x3) {
#pragma unroll
                                    for (int k = 0; k < KCH / 8; ++k) {
                                        // A.[B | B_lo] -> columns [0, bn) and [bn, 2 bn);  A_lo.B -> columns [0, bn)
                                        tc::mma_tf32(tmem, tc::desc_words(aw + k * a_k, a_hiw), tc::desc_words(bw + k * b_k, b_hiw), idesc2, acc);
                                        tc::mma_tf32(tmem, tc::desc_words(alw + k * a_k, a_hiw), tc::desc_words(bw + k * b_k, b_hiw), idesc1, 1u);
                                        acc = 1;
                                    }
                                } else {
#pragma unroll
                                    for (int k = 0; k < KCH / 8; ++k) {
                                        tc::mma_tf32(tmem, tc::desc_words(aw + k * a_k, a_hiw), tc::desc_words(bw + k * b_k, b_hiw), idesc1, acc);
                                        acc = 1;
                                    }
                                }
                                tc::mma_commit(&bar_empty[ring_s]);
                            }
                            __syncwarp();
                            if (++ring_s == kStages) { ring_s = 0; ring_ph ^= 1; }
                        }
                        if (tc::elect_one()) {
                            tc::mma_commit(&bar_done);
                            if (tim) a.timing[tslot + 5] = (unsigned long long)clock64();
                        }
                        __syncwarp();
                    } else {
                        // ---------- workers: prelude in the shadow of the mainloop, then the epilogue
                        const int q = warp & 3, hf = warp >> 2;           // TMEM lane quarter, column half of the tile
                        const int rloc = q * 32 + lane;                   // row inside the tile = TMEM lane
                        const int row = J.m_tile * 128 + rloc;
                        const bool valid = row < rows;
                        const int N = (J.type == J_DH) ? a.n[l] : a.n[l + 1];   // output row pitch
                        const int col0 = J.n_tile * J.bn;
                        if (J.type == J_FWD || J.type == J_HEAD) {
                            // effective bias of this tile's columns; layer 0 adds E_d[dom] . W_0[K0:, cols] in fp32
                            const int bn = J.bn;
                            worker_sync();   // the previous job's epilogue may still be reading s_beff / scratch
                            if (l == 0) {
                                float* part = reinterpret_cast<float*>(scratch);   // [groups][bn]
                                const int groups = kWorkers / bn, per = cdiv(a.dd, groups);
                                const int c = tid % bn, gq = tid / bn;
                                const float* W0dom = a.params + a.off_W[0] + (long long)K0 * N + col0 + c;
                                const float* ed = a.params + a.off_Ed + (long long)pd.dom * a.dd;
                                float s = 0.f;
                                const int k_end = (gq + 1) * per < a.dd ? (gq + 1) * per : a.dd;
#pragma unroll 16
                                for (int k = gq * per; k < k_end; ++k) s = fmaf(ldcg_f(ed + k), ldcg_f(W0dom + (long long)k * N), s);
                                part[gq * bn + c] = s;
                                worker_sync();
                                if (tid < bn) {
                                    float dsum = 0.f;
                                    for (int g2 = 0; g2 < groups; ++g2) dsum += part[g2 * bn + tid];
                                    s_beff[tid] = ldcg_f(a.params + a.off_b[0] + col0 + tid) + dsum;
                                }
                            } else if (tid < bn) {
                                s_beff[tid] = ldcg_f(a.params + a.off_b[l] + col0 + tid);
                            }
                            if (J.type == J_HEAD && tid >= 64 && tid < 64 + NL) s_wd[tid - 64] = ldcg_f(a.params + a.off_w + tid - 64);
                            worker_sync();
                        }
                        // dropout keep bits of this thread's columns (bit c: column cbase + c of this row survives)
                        constexpr int NLH = NL / 2;
                        const int cbase = J.type == J_HEAD ? hf * NLH : hf * 16;   // first column (inside the tile) of this thread
                        uint32_t keepmask = 0xffffffffu;
                        if (dp.enabled && (J.type == J_FWD || J.type == J_HEAD)) {
                            DropoutParams dq = dp;
                            dq.seed = a.dropout_seed + (uint32_t)l;
                            keepmask = 0u;
                            const uint32_t e0 = (uint32_t)row * (uint32_t)N + (uint32_t)(col0 + cbase);
                            const int ncol = J.type == J_HEAD ? NLH : 16;
#pragma unroll
                            for (int c = 0; c < 32; c += 4) {
                                if (c < ncol) {
                                    const uint4 w = dropout_words4(dq, e0 + c);
                                    const uint32_t nib = (w.x < dq.threshold ? 1u : 0u) | (w.y < dq.threshold ? 2u : 0u) |
                                                         (w.z < dq.threshold ? 4u : 0u) | (w.w < dq.threshold ? 8u : 0u);
                                    keepmask |= nib << c;
                                }
                            }
                        }
                        float4 hmask[4];   // dH: the forward activations of this thread's 16 columns (all loads before the wait)
                        if (J.type == J_DH) {
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                hmask[u] = valid ? ldcg_f4(a.H[l] + (long long)row * N + col0 + cbase + u * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                        WSTAMP(12);
                        tc::mbar_wait(&bar_done, njob & 1);
                        tc::tc_fence_after();
                        if (tim && tid == 0) a.timing[tslot + 6] = (unsigned long long)clock64();
                        const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);

                        if (J.type == J_FWD) {
                            WSTAMP(8);
                            float* out = a.H[l + 1] + (long long)row * N + col0 + cbase;
                            const long long oz = (long long)a.max_rows * N;
                            const float dscale = dp.enabled ? dp.scale : 1.0f;
                            float vv[16], v2[16];
                            tc::tmem_ld16(tlane + cbase, vv);
                            if (x3) tc::tmem_ld16(tlane + 32 + cbase, v2);
                            tc::tc_fence_before();
                            tc::mbar_arrive(&bar_tfree);
                            if (valid) {
#pragma unroll
                                for (int jx = 0; jx < 16; jx += 4) {
                                    float h[4];
#pragma unroll
                                    for (int t = 0; t < 4; ++t) {
                                        const float accv = x3 ? vv[jx + t] + v2[jx + t] : vv[jx + t];
                                        h[t] = fmaxf(accv + s_beff[cbase + jx + t], 0.f);
                                        h[t] = (keepmask >> (jx + t)) & 1u ? h[t] * dscale : 0.f;
                                    }
                                    store_pair4(out + jx, oz, make_float4(h[0], h[1], h[2], h[3]), rnd, x3);
                                }
                            }
                            WSTAMP(9);
                        } else if (J.type == J_HEAD) {
                            // last hidden layer + Dense(1) + sigmoid + BCE + ds + dZ_{L-1} + per-tile partials + AUC bins
                            float h[NLH];
                            float zp = 0.f;
                            const float dscale = dp.enabled ? dp.scale : 1.0f;
#pragma unroll
                            for (int n0 = 0; n0 < NLH; n0 += 16) {
                                float vv[16], v2[16];
                                tc::tmem_ld16(tlane + cbase + n0, vv);
                                if (x3) tc::tmem_ld16(tlane + NL + cbase + n0, v2);
#pragma unroll
                                for (int jx = 0; jx < 16; ++jx) {
                                    const float accv = x3 ? vv[jx] + v2[jx] : vv[jx];
                                    float hh = fmaxf(accv + s_beff[cbase + n0 + jx], 0.f);
                                    hh = (keepmask >> (n0 + jx)) & 1u ? hh * dscale : 0.f;
                                    h[n0 + jx] = valid ? hh : 0.f;
                                    zp = fmaf(hh, s_wd[cbase + n0 + jx], zp);
                                }
                            }
                            tc::tc_fence_before();
                            tc::mbar_arrive(&bar_tfree);
                            WSTAMP(8);
                            s_z[hf][rloc] = zp;
                            worker_sync();
                            const float z = s_z[0][rloc] + s_z[1][rloc];
                            const float lo_c = 1e-7f, hi_c = 1.0f - 1e-7f;
                            float dsv = 0.f;
                            double bce = 0.0;
                            if (valid) {
                                const float sgm = z + ldcg_f(a.params + a.off_g);
                                const float pv = 1.0f / (1.0f + expf(-sgm));
                                const float yv = a.y[buf][row];
                                if (a.train) dsv = (fabsf(sgm) <= MAMDR_LOGIT_CLIP) ? __fdiv_rn(__fsub_rn(pv, yv), (float)rows) : 0.f;
                                if (hf == 0) {
                                    const float ph = fminf(fmaxf(pv, lo_c), hi_c);
                                    const float lg = logf(ph / (1.0f - ph));
                                    bce = (double)(fmaxf(lg, 0.f) - lg * yv + log1pf(expf(-fabsf(lg))));
                                    if (pd.probs) pd.probs[(long long)step * a.bs + row] = pv;
                                    if (a.auc_acc) {
                                        int lo_i = 0, hi_i = a.T;
                                        while (lo_i < hi_i) {
                                            const int mid = (lo_i + hi_i) >> 1;
                                            if (s_thr[mid] < pv) lo_i = mid + 1; else hi_i = mid;
                                        }
                                        atomicAdd(&hist_cur[(yv != 0.f ? (a.T + 1) : 0) + lo_i], 1);
                                    }
                                }
                            }
                            WSTAMP(9);
                            double* red_d = reinterpret_cast<double*>(scratch);               // [4]
                            float* red_f = reinterpret_cast<float*>(scratch + 64);            // [4]
                            float* csum = reinterpret_cast<float*>(scratch + 128);            // [2 kinds][8 warps][NLH]
                            if (hf == 0) {
                                double bs = bce;
                                float dgs = dsv;
#pragma unroll
                                for (int o = 16; o > 0; o >>= 1) {
                                    bs += __shfl_xor_sync(0xffffffffu, bs, o);
                                    dgs += __shfl_xor_sync(0xffffffffu, dgs, o);
                                }
                                if (lane == 0) { red_d[q] = bs; red_f[q] = dgs; }
                            }
                            if (a.train) {
                                float* dZ = a.dZ[L - 1] + (long long)row * NL + cbase;
                                const long long oz = (long long)a.max_rows * NL;
                                const bool store = row < a.max_rows;
                                float hd[NLH];   // h * ds  (column sums -> gradient of the Dense(1) kernel)
#pragma unroll
                                for (int c = 0; c < NLH; c += 4) {
                                    float dz[4];
#pragma unroll
                                    for (int t = 0; t < 4; ++t) {
                                        const float dh = __fmul_rn(dsv, s_wd[cbase + c + t]);
                                        dz[t] = h[c + t] > 0.f ? __fmul_rn(dh, inv_keep) : 0.f;
                                        hd[c + t] = h[c + t] * dsv;
                                        h[c + t] = dz[t];
                                    }
                                    // rows past the batch get zeros: dW's K loop runs over whole 32-row chunks
                                    if (store) store_pair4(dZ + c, oz, make_float4(dz[0], dz[1], dz[2], dz[3]), rnd, x3);
                                }
                                WSTAMP(10);
                                const float s_hd = warp_colsum<NLH>(hd, lane);
                                const float s_dz = warp_colsum<NLH>(h, lane);
                                const int cl = NLH == 32 ? lane : lane >> 1;
                                if (NLH == 32 || (lane & 1) == 0) {
                                    csum[(0 * kWorkerWarps + warp) * NLH + cl] = s_hd;
                                    csum[(1 * kWorkerWarps + warp) * NLH + cl] = s_dz;
                                }
                            }
                            worker_sync();
                            if (tid == 0) {
                                a.loss_part[buf * kMaxMT + J.m_tile] = red_d[0] + red_d[1] + red_d[2] + red_d[3];
                                a.dg_part[J.m_tile] = red_f[0] + red_f[1] + red_f[2] + red_f[3];
                            }
                            if (a.train && tid < 2 * NL) {
                                // column c of the tile (half c / NLH): the four row-quarter warps in order
                                const int kind = tid / NL, c = tid - kind * NL, h2 = c / NLH, cc = c - h2 * NLH;
                                float s = 0.f;
#pragma unroll
                                for (int qq = 0; qq < 4; ++qq) s += csum[(kind * kWorkerWarps + h2 * 4 + qq) * NLH + cc];
                                (kind == 0 ? a.dw_part : a.db_part[L - 1])[J.m_tile * NL + c] = s;
                            }
                            worker_sync();
                            WSTAMP(11);
                        } else if (J.type == J_DH) {
                            // dZ_{l-1}[row, col0 + cbase .. +16) = acc * inv_keep * 1[H_l > 0]; per-tile column sums -> db_{l-1}
                            float* out = a.dZ[l - 1] + (long long)row * N + col0 + cbase;
                            const long long oz = (long long)a.max_rows * N;
                            const bool store = row < a.max_rows;
                            float vv[16], v2[16], dzv[16];
                            tc::tmem_ld16(tlane + cbase, vv);
                            if (x3) tc::tmem_ld16(tlane + 32 + cbase, v2);
                            tc::tc_fence_before();
                            tc::mbar_arrive(&bar_tfree);
#pragma unroll
                            for (int jx = 0; jx < 16; jx += 4) {
                                const float4 hq = hmask[jx >> 2];
                                const float hm[4] = {hq.x, hq.y, hq.z, hq.w};
#pragma unroll
                                for (int t = 0; t < 4; ++t) {
                                    const float accv = x3 ? vv[jx + t] + v2[jx + t] : vv[jx + t];
                                    dzv[jx + t] = (valid && hm[t] > 0.f) ? accv * inv_keep : 0.f;
                                }
                                if (store) store_pair4(out + jx, oz, make_float4(dzv[jx], dzv[jx + 1], dzv[jx + 2], dzv[jx + 3]), rnd, x3);
                            }
                            float* csum = reinterpret_cast<float*>(scratch);   // [8 warps][16]
                            const float s_dz = warp_colsum<16>(dzv, lane);
                            worker_sync();   // the previous job's readers of csum are done
                            if ((lane & 1) == 0) csum[warp * 16 + (lane >> 1)] = s_dz;
                            worker_sync();
                            if (tid < 32) {
                                const int h2 = tid >> 4, cc = tid & 15;
                                float s = 0.f;
#pragma unroll
                                for (int qq = 0; qq < 4; ++qq) s += csum[(h2 * 4 + qq) * 16 + cc];
                                a.db_part[l - 1][(long long)J.m_tile * N + col0 + tid] = s;
                            }
                        } else {   // J_DW: split-K partial of dW_l -> partials[l][z][tile][128][bn]
                            const int tile = J.m_tile * J.NT + J.n_tile;
                            const int half = J.bn >> 1;   // columns per thread: 32 (bn = 64) or 16 (bn = 32)
                            float* mine = a.partials[l] + (((long long)J.z * J.tiles + tile) * 128 + rloc) * J.bn + hf * half;
                            for (int n0 = 0; n0 < half; n0 += 16) {
                                float vv[16], v2[16];
                                tc::tmem_ld16(tlane + hf * half + n0, vv);
                                if (x3) tc::tmem_ld16(tlane + J.bn + hf * half + n0, v2);
#pragma unroll
                                for (int jx = 0; jx < 16; jx += 4) {
                                    float4 o4;
                                    if (x3) o4 = make_float4(vv[jx] + v2[jx], vv[jx + 1] + v2[jx + 1], vv[jx + 2] + v2[jx + 2], vv[jx + 3] + v2[jx + 3]);
                                    else o4 = make_float4(vv[jx], vv[jx + 1], vv[jx + 2], vv[jx + 3]);
                                    __stcg(reinterpret_cast<float4*>(mine + n0 + jx), o4);
                                }
                            }
                            tc::tc_fence_before();
                            tc::mbar_arrive(&bar_tfree);
                        }
                    }
                    ++njob;
                }
                if (phase == 2 * L - 1 && early_done && cta >= jobs_last_bwd)
                    update_items((long long)(cta - jobs_last_bwd) * kThreads + tid, (long long)(G - jobs_last_bwd) * kThreads, 1);
                // while the few head tiles run, everyone else stages the next mini-batch
                if (phase == L - 1 && step + 1 < pd.steps && warp < kWorkerWarps) {
                    const int first = G > 2 * mt ? mt : 0;
                    if (cta >= first) gather_rows(a, pd, step + 1, buf ^ 1, (cta - first) * kWorkerWarps + warp, (G - first) * kWorkerWarps, lane, rnd, x3);
                }
            } else {
                // ---------- update phase: Adam / SGD on the parameters whose gradients became final in the last backward
                // phase (E_d, W_0, b_0); the rest was already applied by the idle CTAs of that phase (early_done)
                update_items((long long)cta * kThreads + tid, (long long)G * kThreads, early_done ? 2 : 0);
                if (a.opt_kind == 0) {
                    b1pow = __fmul_rn(b1pow, a.beta1);
                    b2pow = __fmul_rn(b2pow, a.beta2);
                }
                step_ctr += 1;
            }
            if (a.timing) {
                __syncthreads();
                if (tid == 0 && tim) a.timing[tslot + 1] = gtime();
            }
            grid_barrier(a.bar, bar_target);
        }
        // the Keras loss of this mini-batch (value only).  loss_part is double-buffered by step parity: the next
        // write to this buffer is two head phases (>= one grid barrier that this thread also passes) away.
        if (cta == G - 1 && tid == 0) {
            double bs = 0.0;
            for (int m = 0; m < mt; ++m) bs += __ldcg(a.loss_part + buf * kMaxMT + m);
            pd.losses[step] = (float)(bs / (double)rows + (double)a.frozen_reg + (double)a.l2_emb * __ldcg(a.ed_sq));
        }
    }

    // ---- epilogue of the pass: AUC accumulators (suffix sums of the pass histogram; the histogram is double-buffered
    // by pass parity, so the next pass of a program may already be filling the other one)
    if (cta == 0) {
        if (a.auc_acc && warp < 2) {
            // warp 0: negatives, warp 1: positives.  lane owns a contiguous run of bins; suffix scan across lanes.
            const int T1 = a.T + 1;
            int* hsrc = hist_cur + warp * T1;
            int* sfx = reinterpret_cast<int*>(scratch) + warp * (kMaxThr + 32);
            const int per = cdiv(T1, 32);
            const int b0 = lane * per, b1 = (b0 + per < T1) ? b0 + per : T1;
            int run = 0;
            for (int b = b1 - 1; b >= b0; --b) { run += __ldcg(hsrc + b); sfx[b] = run; }
            int tot = run, incl = run;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int up = __shfl_down_sync(0xffffffffu, incl, o);
                if (lane + o < 32) incl += up;
            }
            const int above = incl - tot;   // sum over the lanes after this one
            for (int b = b0; b < b1; ++b) { sfx[b] += above; hsrc[b] = 0; }
        }
        __syncthreads();
        if (a.auc_acc) {
            const int T1 = a.T + 1;
            const int* neg = reinterpret_cast<const int*>(scratch);
            const int* pos = neg + (kMaxThr + 32);
            for (int j = tid; j < a.T; j += kThreads) {
                const int npos = pos[0], nneg = neg[0];
                const int tp = j + 1 < T1 ? pos[j + 1] : 0, fp = j + 1 < T1 ? neg[j + 1] : 0;
                a.auc_acc[0 * a.T + j] += (float)tp;
                a.auc_acc[1 * a.T + j] += (float)fp;
                a.auc_acc[2 * a.T + j] += (float)(npos - tp);
                a.auc_acc[3 * a.T + j] += (float)(nneg - fp);
            }
        }
        __syncthreads();
    }
    // the next op of a program may reset or read the accumulators: order it after the fold
    if (a.auc_acc && oi + 1 < a.n_ops) grid_barrier(a.bar, bar_target);
    ++pass_idx;
    }   // program ops

    if (cta == 0 && tid == 0 && a.train) {
        a.state->step = step_ctr;
        a.state->b1pow = b1pow;
        a.state->b2pow = b2pow;
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) tc::tmem_dealloc(tmem, kTmemCols);
}

// ---- workspace layout ---------------------------------------------------------------------------------------------
struct PassWs {
    size_t bar, hist, X[2], y[2], H[MAMDR_MAX_LAYERS], dZ[MAMDR_MAX_LAYERS], partials[MAMDR_MAX_LAYERS], db_part[MAMDR_MAX_LAYERS];
    size_t dw_part, dg_part, loss_part, db0_red, gEd_row, ed_row, ed_sq, wpair, wz, total;
};

inline PassWs pass_ws(const mamdr_mlp_desc& d, int B) {
    PassWs w;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off = (off + bytes + 1023) / 1024 * 1024;
        return o;
    };
    const int L = d.n_layers;
    const int K0 = d.emb_dim[0] + d.emb_dim[1];
    const int Bp = (B + 127) / 128 * 128;   // whole 128-row tiles: the epilogues write zero rows up to the tile edge
    const int mt = Bp / 128;
    w.bar = take(64);
    w.hist = take((size_t)2 * 2 * (kMaxThr + 1) * 4);   // double-buffered by pass parity
    // GEMM operands are pair arrays [2][Bp][width]: plane 0 = value, plane 1 = its 3xTF32 "lo" part
    for (int b = 0; b < 2; ++b) { w.X[b] = take((size_t)2 * Bp * K0 * 4); w.y[b] = take((size_t)Bp * 4); }
    for (int l = 0; l < L; ++l) {
        w.H[l] = l >= 1 ? take((size_t)2 * Bp * d.hidden[l - 1] * 4) : 0;
        w.dZ[l] = take((size_t)2 * Bp * d.hidden[l] * 4);
        const int in = l == 0 ? K0 : d.hidden[l - 1];
        const int bn = d.hidden[l] < 64 ? d.hidden[l] : 64;
        const size_t tiles = (size_t)((in + 127) / 128) * (d.hidden[l] / bn);
        w.partials[l] = take(tiles * kMaxSplit * 128 * bn * 4);
        w.db_part[l] = take((size_t)mt * d.hidden[l] * 4);
    }
    w.dw_part = take((size_t)mt * d.hidden[L - 1] * 4);
    w.dg_part = take((size_t)mt * 4);
    w.loss_part = take((size_t)2 * kMaxMT * 8);
    w.db0_red = take((size_t)d.hidden[0] * 4);
    w.gEd_row = take((size_t)d.emb_dim[2] * 4);
    w.ed_row = take((size_t)d.emb_dim[2] * 4);
    w.ed_sq = take(8);
    w.wz = ((size_t)(d.arena_floats - d.off_domain_emb) + 31) / 32 * 32;   // floats between the two planes of the kernel shadow
    w.wpair = take(2 * w.wz * 4);                                          // dense span of the arena only
    w.total = off;
    return w;
}

inline size_t smem_bytes() { return (size_t)kStages * STAGE_BYTES + kScratchBytes + 1024; }

}  // namespace passk

using namespace passk;

static int pass_supported(mamdr_ctx* ctx, const mamdr_mlp_desc* d, int max_batch) {
    MAMDR_REQUIRE(ctx, d->n_layers >= 1 && d->n_layers <= MAMDR_MAX_LAYERS, MAMDR_E_INVALID, "n_layers out of range");
    MAMDR_REQUIRE(ctx, !d->emb_trainable, MAMDR_E_UNSUPPORTED, "the pass kernel covers frozen user/item tables (trainable tables use the per-step path)");
    const int K0 = d->emb_dim[0] + d->emb_dim[1];
    MAMDR_REQUIRE(ctx, d->emb_dim[0] % 4 == 0 && d->emb_dim[1] % 4 == 0 && d->emb_dim[2] % 4 == 0 && K0 % 32 == 0, MAMDR_E_UNSUPPORTED,
                  "pass kernel needs emb dims that are multiples of 4 and user+item width a multiple of 32");
    for (int l = 0; l < d->n_layers; ++l)
        MAMDR_REQUIRE(ctx, d->hidden[l] % 32 == 0 && (d->hidden[l] == 32 || d->hidden[l] % 64 == 0) && d->hidden[l] <= 4096, MAMDR_E_UNSUPPORTED,
                      "pass kernel needs hidden widths of 32 or multiples of 64");
    const int nl = d->hidden[d->n_layers - 1];
    MAMDR_REQUIRE(ctx, nl == 32 || nl == 64, MAMDR_E_UNSUPPORTED, "pass kernel needs a last hidden width of 32 or 64");
    MAMDR_REQUIRE(ctx, max_batch >= 1 && max_batch <= 128 * kMaxMT, MAMDR_E_UNSUPPORTED, "batch too large for the pass kernel");
    MAMDR_REQUIRE(ctx, d->dropout_rate >= 0.f && d->dropout_rate < 1.f, MAMDR_E_INVALID, "dropout_rate must be in [0,1)");
    return MAMDR_OK;
}

int mamdr_pass_init_kernels(mamdr_ctx* ctx) {
    MAMDR_CUDA_OK(ctx, cudaFuncSetAttribute(pass_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes()));
    MAMDR_CUDA_OK(ctx, cudaFuncSetAttribute(pass_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes()));
    return MAMDR_OK;
}

// debug hook (not part of the product contract): per-phase globaltimer stamps of the next pass launches
extern "C" void mamdr_program_abort(mamdr_ctx* ctx);

extern "C" int mamdr_debug_pass_timing(mamdr_ctx* ctx, void* buf_dev, int64_t capacity_u64) {
    if (!ctx) return MAMDR_E_INVALID;
    ctx->dbg_timing = buf_dev;
    ctx->dbg_timing_cap = buf_dev ? capacity_u64 : 0;
    return MAMDR_OK;
}

void mamdr_pass_free_ctx(mamdr_ctx* ctx) {
    mlptc::free_tmap_cache(ctx);
    mamdr_program_abort(ctx);
}

extern "C" size_t mamdr_mlp_pass_workspace_bytes(const mamdr_mlp_desc* desc, int32_t max_batch) {
    if (!desc || max_batch < 1 || desc->n_layers < 1 || desc->n_layers > MAMDR_MAX_LAYERS) return 0;
    return pass_ws(*desc, max_batch).total;
}

extern "C" int mamdr_mlp_pass_supported(mamdr_ctx* ctx, const mamdr_mlp_desc* desc, int32_t max_batch) {
    if (!ctx || !desc) return MAMDR_E_INVALID;
    return pass_supported(ctx, desc, max_batch);
}

struct Program {
    std::vector<ProgOp> ops;
    bool has_pass = false;
    PassArgs a;
    MapTable mp;
};

int mamdr_meta_launch(mamdr_ctx* ctx, int meta_op, const MetaArgs& a, mamdr_stream stream);   // optim.cu

static int launch_program(mamdr_ctx* ctx, const MapTable& mp, const PassArgs& a, cudaStream_t st) {
    MAMDR_CUDA_OK(ctx, cudaMemsetAsync(a.bar, 0, 64, st));
    const size_t smem = smem_bytes();
    void* kargs[] = {(void*)&mp, (void*)&a};
    const void* fn = a.n[a.L] == 64 ? (const void*)pass_kernel<64> : (const void*)pass_kernel<32>;
    MAMDR_CUDA_OK(ctx, cudaLaunchCooperativeKernel(fn, dim3(ctx->sm_count), dim3(kThreads), kargs, smem, st));
    return MAMDR_OK;
}

static int run_pass(mamdr_ctx* ctx, const mamdr_mlp_desc* d, const mamdr_pass* ps, const float* ut, const float* it, float* params,
                    float* m, float* v, float* grads, void* ws_, size_t ws_bytes, void* opt_state, float* losses, float* probs,
                    float* auc_acc, const float* thr, int T, int train, int optimizer, float lr, float beta1, float beta2,
                    float eps, int precision_mode, cudaStream_t st) {
    MAMDR_REQUIRE(ctx, ctx && d && ps, MAMDR_E_INVALID, "NULL ctx/desc/pass");
    int rc = pass_supported(ctx, d, ps->batch_size);
    if (rc) return rc;
    MAMDR_REQUIRE(ctx, precision_mode == MAMDR_PREC_TF32 || precision_mode == MAMDR_PREC_TF32X3, MAMDR_E_UNSUPPORTED,
                  "the pass kernel runs the tcgen05 modes (tf32, tf32x3); fp32 uses the per-step SIMT path");
    MAMDR_REQUIRE(ctx, ps->steps >= 1 && ps->n_data >= 1 && ps->batch_size >= 1, MAMDR_E_INVALID, "empty pass");
    MAMDR_REQUIRE(ctx, (int64_t)(ps->steps - 1) * ps->batch_size < ps->n_data, MAMDR_E_INVALID, "steps * batch_size runs past n_data");
    MAMDR_REQUIRE(ctx, ps->domain >= 0 && ps->domain < d->n_domain, MAMDR_E_INVALID, "domain id out of range");
    MAMDR_REQUIRE(ctx, ps->uid_dev && ps->pid_dev && ps->label_dev && ut && it, MAMDR_E_INVALID, "NULL data column / table");
    MAMDR_REQUIRE(ctx, params && aligned16(params) && ws_ && aligned16(ws_) && losses && opt_state, MAMDR_E_INVALID, "NULL or misaligned buffer");
    if (train) MAMDR_REQUIRE(ctx, optimizer == 1 || (m && v && aligned16(m) && aligned16(v)), MAMDR_E_INVALID, "Adam slots NULL or misaligned");
    if (auc_acc) MAMDR_REQUIRE(ctx, thr && T >= 2 && T <= kMaxThr - 1, MAMDR_E_INVALID, "bad AUC thresholds (2 <= T <= 1023)");
    const PassWs w = pass_ws(*d, ps->batch_size);
    MAMDR_REQUIRE(ctx, ws_bytes >= w.total, MAMDR_E_WORKSPACE, "workspace too small: %zu < %zu", ws_bytes, w.total);
    unsigned char* ws = (unsigned char*)ws_;
    const int L = d->n_layers;
    const int Bp = (ps->batch_size + 127) / 128 * 128;

    PassArgs a;
    memset(&a, 0, sizeof(a));
    a.L = L;
    a.n[0] = d->emb_dim[0] + d->emb_dim[1];
    for (int l = 0; l < L; ++l) a.n[l + 1] = d->hidden[l];
    a.du = d->emb_dim[0]; a.di = d->emb_dim[1]; a.dd = d->emb_dim[2];
    a.n_domain = d->n_domain;
    a.off_Ed = d->off_domain_emb; a.off_w = d->off_dense_kernel; a.off_g = d->off_global_bias; a.arena = d->arena_floats;
    int ns = 0;
    a.seg[ns++] = Seg{d->off_domain_emb, d->n_domain * d->emb_dim[2], SEG_ED, 0};
    for (int l = 0; l < L; ++l) {
        a.off_W[l] = d->off_kernel[l]; a.off_b[l] = d->off_bias[l];
        const int in = l == 0 ? a.n[0] + a.dd : a.n[l];
        a.seg[ns++] = Seg{d->off_kernel[l], in * a.n[l + 1], SEG_KERNEL, l};
        a.seg[ns++] = Seg{d->off_bias[l], a.n[l + 1], SEG_BIAS, l};
    }
    a.seg[ns++] = Seg{d->off_dense_kernel, a.n[L], SEG_DENSE, 0};
    a.seg[ns++] = Seg{d->off_global_bias, 1, SEG_GBIAS, 0};
    a.nseg = ns;
    a.params = params; a.m = m; a.v = v; a.grads = grads; a.Eu = ut; a.Ei = it;
    a.bs = ps->batch_size; a.max_rows = Bp;
    ProgOp op;
    memset(&op, 0, sizeof(op));
    op.kind = PROG_PASS; op.domain = ps->domain; op.steps = ps->steps; op.n_data = ps->n_data;
    op.uid = ps->uid_dev; op.pid = ps->pid_dev; op.order = ps->order_dev; op.label = ps->label_dev;
    op.losses = losses; op.probs = probs;
    for (int b = 0; b < 2; ++b) { a.X[b] = (float*)(ws + w.X[b]); a.y[b] = (float*)(ws + w.y[b]); }
    for (int l = 0; l < L; ++l) {
        a.H[l] = l >= 1 ? (float*)(ws + w.H[l]) : nullptr;
        a.dZ[l] = (float*)(ws + w.dZ[l]);
        a.partials[l] = (float*)(ws + w.partials[l]);
        a.db_part[l] = (float*)(ws + w.db_part[l]);
    }
    a.dw_part = (float*)(ws + w.dw_part); a.dg_part = (float*)(ws + w.dg_part); a.loss_part = (double*)(ws + w.loss_part);
    a.db0_red = (float*)(ws + w.db0_red); a.gEd_row = (float*)(ws + w.gEd_row); a.ed_row = (float*)(ws + w.ed_row);
    a.ed_sq = (double*)(ws + w.ed_sq);
    a.wpair = (float*)(ws + w.wpair) - d->off_domain_emb;   // indexed with arena offsets
    a.wz = (long long)w.wz;
    a.hist = (int*)(ws + w.hist);
    a.bar = (unsigned int*)(ws + w.bar);
    a.state = (OptState*)opt_state;
    a.opt_kind = optimizer; a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps;
    const float keep = 1.0f - d->dropout_rate;
    a.dropout_enabled = d->dropout_rate > 0.f ? 1 : 0;
    a.dropout_seed = d->dropout_seed;
    const double thrd = floor((double)keep * 4294967296.0);
    a.dropout_threshold = thrd >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)thrd;
    a.dropout_scale = 1.0f / keep;
    a.l2_emb = d->l2_emb; a.frozen_reg = d->frozen_reg;
    a.auc_acc = auc_acc; a.thr = thr; a.T = auc_acc ? T : 0;
    a.train = train ? 1 : 0;
    a.passes = precision_mode == MAMDR_PREC_TF32X3 ? 3 : 1;
    a.timing = (unsigned long long*)ctx->dbg_timing;
    a.timing_cap = ctx->dbg_timing_cap;

    MapTable mp;
    memset(&mp, 0, sizeof(mp));
    bool ok = true;
    // every GEMM operand is a pair array; the lo plane lies Bp * width floats (kernels: wz floats) behind the hi plane
    const float* wsrc = a.wpair;   // B operands of fwd / dH: the pair shadow of the kernels
    for (int b = 0; b < 2; ++b) {
        const uint64_t z = (uint64_t)Bp * a.n[0];
        ok = ok && mlptc::pair_kmajor_map(ctx, &mp.xk[b], a.X[b], Bp, a.n[0], 128, z) && mlptc::pair_mnmajor_map(ctx, &mp.xmn[b], a.X[b], Bp, a.n[0], z);
    }
    for (int l = 0; l < L; ++l) {
        const uint64_t zh = (uint64_t)Bp * a.n[l], zd = (uint64_t)Bp * a.n[l + 1];
        if (l >= 1) {
            ok = ok && mlptc::pair_kmajor_map(ctx, &mp.hk[l], a.H[l], Bp, a.n[l], 128, zh) && mlptc::pair_mnmajor_map(ctx, &mp.hmn[l], a.H[l], Bp, a.n[l], zh);
            ok = ok && mlptc::pair_kmajor_map(ctx, &mp.dzk[l], a.dZ[l], Bp, a.n[l + 1], 128, zd);
            ok = ok && mlptc::pair_kmajor_map(ctx, &mp.wb[l], wsrc + d->off_kernel[l], a.n[l], a.n[l + 1], 32, (uint64_t)a.wz);
        }
        ok = ok && mlptc::pair_mnmajor_map(ctx, &mp.dzmn[l], a.dZ[l], Bp, a.n[l + 1], zd);
        ok = ok && mlptc::pair_mnmajor_map(ctx, &mp.wf[l], wsrc + d->off_kernel[l], a.n[l], a.n[l + 1], (uint64_t)a.wz);
    }
    MAMDR_REQUIRE(ctx, ok, MAMDR_E_CUDA, "cuTensorMapEncodeTiled failed (pass kernel)");

    if (ctx->prog) {
        // recording: the first pass fixes the launch-wide arguments, later passes must agree with them
        Program* pr = static_cast<Program*>(ctx->prog);
        MAMDR_REQUIRE(ctx, train, MAMDR_E_INVALID, "only training passes and meta sweeps can be recorded into a program");
        if (!pr->has_pass) {
            pr->a = a; pr->mp = mp; pr->has_pass = true;
        } else {
            const PassArgs& c = pr->a;
            MAMDR_REQUIRE(ctx, c.params == a.params && c.m == a.m && c.v == a.v && c.X[0] == a.X[0] && c.bs == a.bs && c.passes == a.passes &&
                              c.opt_kind == a.opt_kind && c.lr == a.lr && c.Eu == a.Eu && c.Ei == a.Ei && c.state == a.state &&
                              c.auc_acc == a.auc_acc && c.grads == a.grads && c.arena == a.arena,
                          MAMDR_E_INVALID, "passes of one program must share model, workspace, batch size, optimizer and precision");
        }
        pr->ops.push_back(op);
        return MAMDR_OK;
    }
    a.ops = nullptr; a.n_ops = 1; a.inline_op = op;
    return launch_program(ctx, mp, a, st);
}

bool mamdr_prog_recording(const mamdr_ctx* ctx) { return ctx && ctx->prog; }

int mamdr_prog_push_meta(mamdr_ctx* ctx, int meta_op, const MetaArgs& m) {
    Program* pr = static_cast<Program*>(ctx->prog);
    ProgOp op;
    memset(&op, 0, sizeof(op));
    op.kind = PROG_META; op.method = (meta_op & 0xff) | (m.method << 8); op.n = m.n;
    op.w0 = m.w0; op.w1 = m.w1; op.r0 = m.r0; op.r1 = m.r1; op.r2 = m.r2; op.f0 = m.f0; op.f1 = m.f1;
    pr->ops.push_back(op);
    return MAMDR_OK;
}

extern "C" int mamdr_program_begin(mamdr_ctx* ctx) {
    MAMDR_REQUIRE(ctx, ctx != nullptr, MAMDR_E_INVALID, "ctx is NULL");
    MAMDR_REQUIRE(ctx, ctx->prog == nullptr, MAMDR_E_INVALID, "a program is already being recorded");
    ctx->prog = new Program();
    return MAMDR_OK;
}

extern "C" void mamdr_program_abort(mamdr_ctx* ctx) {
    if (!ctx || !ctx->prog) return;
    delete static_cast<Program*>(ctx->prog);
    ctx->prog = nullptr;
}

extern "C" int64_t mamdr_program_op_bytes(void) { return (int64_t)sizeof(ProgOp); }

extern "C" int mamdr_program_end(mamdr_ctx* ctx, void* ops_dev, size_t ops_dev_bytes, int32_t* n_ops_out, mamdr_stream stream) {
    MAMDR_REQUIRE(ctx, ctx && ctx->prog, MAMDR_E_INVALID, "no program is being recorded");
    Program* pr = static_cast<Program*>(ctx->prog);
    ctx->prog = nullptr;   // from here on every call executes immediately again
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = pr->ops.size();
    if (n_ops_out) *n_ops_out = (int32_t)n;
    int rc = MAMDR_OK;
    if (n == 0) {
        // nothing recorded
    } else if (!pr->has_pass) {
        // meta sweeps only: replay them as ordinary launches
        for (size_t i = 0; i < n && rc == MAMDR_OK; ++i) {
            const ProgOp& o = pr->ops[i];
            MetaArgs m{o.w0, o.w1, o.r0, o.r1, o.r2, o.f0, o.f1, o.method >> 8, (int64_t)o.n};
            rc = mamdr_meta_launch(ctx, o.method & 0xff, m, stream);
        }
    } else if (!ops_dev || ops_dev_bytes < n * sizeof(ProgOp) || !aligned16(ops_dev)) {
        MAMDR_SET_ERR(ctx, "program buffer too small or misaligned: %zu ops need %zu bytes", n, n * sizeof(ProgOp));
        rc = MAMDR_E_WORKSPACE;
    } else {
        cudaError_t e = cudaMemcpyAsync(ops_dev, pr->ops.data(), n * sizeof(ProgOp), cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) {
            MAMDR_SET_ERR(ctx, "cudaMemcpyAsync(program): %s", cudaGetErrorString(e));
            rc = MAMDR_E_CUDA;
        } else {
            pr->a.ops = (const ProgOp*)ops_dev;
            pr->a.n_ops = (int)n;
            rc = launch_program(ctx, pr->mp, pr->a, st);
        }
    }
    delete pr;
    return rc;
}

extern "C" int mamdr_mlp_train_pass(mamdr_ctx* ctx, const mamdr_mlp_desc* desc, const mamdr_pass* pass,
                                    const float* user_table_dev, const float* item_table_dev, float* params_dev,
                                    float* m_dev, float* v_dev, float* grads_dev, void* ws_dev, size_t ws_bytes,
                                    void* opt_state_dev, float* losses_dev, float* auc_acc_dev, const float* thresholds_dev,
                                    int32_t num_thresholds, int32_t optimizer, float lr, float beta1, float beta2, float eps,
                                    int32_t precision_mode, mamdr_stream stream) {
    MAMDR_REQUIRE(ctx, optimizer == 0 || optimizer == 1, MAMDR_E_INVALID, "optimizer must be 0 (Adam) or 1 (SGD)");
    return run_pass(ctx, desc, pass, user_table_dev, item_table_dev, params_dev, m_dev, v_dev, grads_dev, ws_dev, ws_bytes,
                    opt_state_dev, losses_dev, nullptr, auc_acc_dev, thresholds_dev, num_thresholds, 1, optimizer, lr, beta1,
                    beta2, eps, precision_mode, (cudaStream_t)stream);
}

extern "C" int mamdr_mlp_eval_pass(mamdr_ctx* ctx, const mamdr_mlp_desc* desc, const mamdr_pass* pass,
                                   const float* user_table_dev, const float* item_table_dev, const float* params_dev,
                                   void* ws_dev, size_t ws_bytes, void* opt_state_dev, float* losses_dev, float* probs_dev,
                                   float* auc_acc_dev, const float* thresholds_dev, int32_t num_thresholds,
                                   int32_t precision_mode, mamdr_stream stream) {
    return run_pass(ctx, desc, pass, user_table_dev, item_table_dev, const_cast<float*>(params_dev), nullptr, nullptr, nullptr,
                    ws_dev, ws_bytes, opt_state_dev, losses_dev, probs_dev, auc_acc_dev, thresholds_dev, num_thresholds, 0, 0,
                    0.f, 0.f, 0.f, 0.f, precision_mode, (cudaStream_t)stream);
}
