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
// the B200 rooflines; everything is bound by the latency of DEPENDENT steps.  v1 / v2 of this kernel walked 7 phases
// of 128-row tile jobs per mini-batch, separated by grid barriers (75 / 64 us per mini-batch; each phase re-loaded
// 128 x K activation tiles into 64 SMs at the 76 B/clk/SM TMA rate).  v3 (this file) removes the dependence between
// CTAs from the forward / backward-activation chain altogether:
//
//   phase 0  "chain":  the mini-batch is cut into row groups of CR = 16 rows; ONE CTA runs the whole chain
//            X -> H_1 -> .. -> H_L -> Dense(1) -> sigmoid-BCE -> dZ_{L-1} -> .. -> dZ_0 for its rows.  Every GEMM is
//            computed TRANSPOSED ("swap A/B"): out^T[feature, row] = W^T . act^T, i.e. the WEIGHTS are the M = 128
//            operand of tcgen05.mma (streamed once through a 4-stage TMA ring, 32 KB per k-chunk) and the activations
//            are the N = 16 (+16 "lo") operand, which never leaves shared memory: an epilogue thread owns ONE output
//            feature (= TMEM lane = k index of the next GEMM) for 8 rows, applies bias / ReLU / dropout (or the ReLU mask
//            of the backward pass) and writes the next B operand straight into the 128B-swizzled K-major layout.
//            No grid barrier, no global round trip and no pipeline refill between layers; the TMA ring keeps streaming
//            weights across layer boundaries.  H_l / dZ_l also go to global memory (pair arrays) for phase 1.
//            CTAs without a row group gather the NEXT mini-batch's embedding rows meanwhile (frozen tables).
//   phase 1  "dW + update":  every weight gradient dW_l = H_l^T . dZ_l as split-K tile jobs (128 x 64 tiles, K = 128
//            rows per job) over all CTAs; when the S split-K partials of a tile are in memory (a per-tile counter,
//            release / acquire) each of the S CTAs reduces ITS 128 / S rows of the tile in fixed order and applies
//            Adam / SGD to them -- no separate update phase.  The domain jobs (db_0, dE_d[dom], the domain block of W_0
//            and the next fold) and the column-sum job (biases, Dense(1), the other rows of E_d) apply theirs directly.
// = 2 grid barriers per mini-batch instead of 7.
//
// CTA = 10 warps: warps 0-7 = epilogue / element-wise workers (warp w <-> TMEM lanes 32 (w % 4) .. +31, row half
// w / 4), warp 8 = TMA producer, warp 9 = tcgen05.mma issuer (owns the TMEM allocation).  Both issue warps run their
// loops CONVERGED, one elected lane issuing: TMA and MMA descriptors then live in uniform registers (under
// `if (lane == 0)` every instruction pays an R2UR waterfall: 130-200 cycles per MMA instead of 50,
// profiles/r2_probe_mainloop.txt).
//
// 3xTF32 (MAMDR_PREC_TF32X3): every GEMM operand exists as a PAIR: plane 0 = hi = rn_tf32(x), plane 1 = lo =
// rn_tf32(x - hi), written by whoever produces the operand (hi rounded to NEAREST: |lo| <= 2^-12 |x|).  The
// N operand is [hi rows | lo rows], so a k-step is two MMAs: W_hi.[act_hi | act_lo] (two accumulator halves) and
// W_lo.act_hi (first half); the epilogue adds the halves.  Dropped: lo.lo (2^-24 relative: fp32-level products).
//
// The domain embedding row is the same for every sample of a batch (utils/dataset.py:73-99: per-domain
// datasets), so X is only [E_u | E_i] and the domain block of layer 0 is folded into its bias in fp32
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
constexpr int A_BYTES = 128 * KCH * 4;        // 16 KB: one M = 128 operand tile
constexpr int B_BYTES = 64 * KCH * 4;         // 8 KB: one dW B tile at BN = 64
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;   // dW jobs: [A | A_lo | B | B_lo] = 48 KB
constexpr int kStages = 4;
constexpr int CR = 16;                        // rows of a chain job (row group)
constexpr int kChainStage = 2 * A_BYTES;      // chain: [W | W_lo] = 32 KB per stage
constexpr int kRingBytes = kStages * kChainStage;        // 128 KB; the two activation buffers follow
constexpr int kMaxWidth = 256;                // widest layer the chain keeps in shared memory
constexpr int kBChunk = 2 * CR * KCH * 4;     // 4 KB: one K chunk of the N operand, [hi rows | lo rows] x 32 k
constexpr int kActBytes = (kMaxWidth / KCH) * kBChunk;   // 32 KB
static_assert(kRingBytes + 2 * kActBytes == kStages * STAGE_BYTES, "the chain and dW layouts overlay the same 192 KB");
constexpr int kScratchBytes = 20 * 1024;
constexpr int kMaxThr = 1024;
constexpr int kMaxSplit = 8;
constexpr int kDomJobs = 16;                  // the domain block (gradient GEMV, optimizer apply, fold) is split over this many CTAs
constexpr int kMaxGroups = 8192 / CR;         // row groups per mini-batch (batch <= 8192)
constexpr int kMaxSegs = 4 * MAMDR_MAX_LAYERS;
constexpr int kBarBytes = 64 + 4 * 8 * MAMDR_MAX_LAYERS;   // grid-barrier counter + one counter per dW tile (<= 8 per layer)
constexpr int kAccGroups = 4;                 // chain: K chunks of a GEMM are spread over this many TMEM accumulators, summed with
                                              // round-to-nearest adds in the epilogue (the tensor core's own fp32 accumulation
                                              // truncates: fewer dependent adds on smaller partial sums per accumulator)
constexpr uint32_t kSlotCols = kAccGroups * 2 * CR;   // 128 columns per accumulator slot
constexpr uint32_t kTmemCols = 2 * kSlotCols; // chain: two slots; dW: one accumulator of up to 2 x 64 columns

enum { J_NONE = 0, J_DW, J_DOM, J_RED };
enum { S_FWD = 0, S_HEAD, S_DH, S_DX };   // S_DX: dX = dZ_0 . W_0[0:K0]^T (trainable user / item tables only)
enum { SEG_ED = 0, SEG_KERNEL, SEG_BIAS, SEG_DENSE, SEG_GBIAS };

// chain-phase layout of the scratch area
constexpr int kScrMask = 0;                   // [kMaxSegs / 2][256] bytes: ReLU / dropout survivor bits of the 8 rows of a thread
constexpr int kScrZp = 6144;                  // [2][4][8] floats: logit partials per (row half, lane quarter)
constexpr int kScrDs = 6400;                  // [16] floats: ds of the rows
constexpr int kScrComb = 6656;                // [2][2][128] floats: row-half combine of column sums

struct MapTable {   // kernel parameter (param space is a legal tensor-map address space); every map is a pair map
    CUtensorMap xk[2], xmn[2];                  // X double buffer: K-major [B, K0] (box CR rows) / MN-major view
    CUtensorMap hmn[MAMDR_MAX_LAYERS];          // H_l  MN-major (A of dW_l), l = 1..L-1
    CUtensorMap dzmn[MAMDR_MAX_LAYERS];         // dZ_l MN-major (B of dW_l), l = 0..L-1
    CUtensorMap wf[MAMDR_MAX_LAYERS];           // W_l^T K-major [M = out, K = in] (A of the transposed forward GEMM)
    CUtensorMap wb[MAMDR_MAX_LAYERS];           // W_l  K-major  [M = in, K = out] (A of the transposed dH GEMM), l = 1..L-1
};

struct Seg { long long off; int numel; int kind; int layer; };

struct SegD {   // one GEMM of the chain: an M tile of a forward layer or of a dH layer
    int kind, layer, mtile, nch, ngrp;   // nch = K chunks; ngrp = valid 32-feature groups of the tile (1..4)
    int dep;                             // the last segment whose epilogue wrote this segment's N operand (-1: X)
    int bsrc, bdst;                      // activation buffer read by the MMAs / written by the epilogue
    int mseg;                            // S_DH: the forward segment that holds the ReLU mask bits of these features
};

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
    float* wpair;                     // arena-indexed pair shadow of the kernels W_l [in, out] (M operand of the dH GEMMs):
    long long wz;                     // plane 0 at wpair, plane 1 at wpair + wz
    float* wT[MAMDR_MAX_LAYERS];      // pair shadow of W_l^T [out, in] (M operand of the forward GEMMs; layer 0: the K0 user / item
                                      // rows only); plane 1 lies n[l + 1] * n[l] floats behind plane 0
    const float *Eu, *Ei;
    // ---- data
    int bs, max_rows;
    // ---- workspace (pair arrays: the lo plane lies max_rows * width floats behind the hi plane)
    float *X[2], *y[2], *H[MAMDR_MAX_LAYERS], *dZ[MAMDR_MAX_LAYERS], *partials[MAMDR_MAX_LAYERS];
    float *db_part[MAMDR_MAX_LAYERS];   // per-row-group column sums of dZ_l
    float *dw_part, *dg_part, *fold_part;   // fold_part: [kDomJobs][n1] k-slice partials of E_d[dom] . W_0dom
    unsigned int* tile_ctr;             // per dW tile: split-K partials written so far (monotonic over the launch)
    // ---- trainable user / item tables (config #2): the tables live in the arena, one mini-batch per launch; the kernel
    // leaves dX [rows, K0] and the batch's ids for the sparse-gradient de-duplication + table sweeps that follow it
    int emb_trainable;
    float* dX;
    int32_t* bid[2];
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
    int type, layer, m_tile, n_tile, z, bn, nch, c_beg, tiles, NT, gtile, S;
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

// jobs of the dW phase: [domain jobs | column-sum job | split-K tile jobs of dW_0, dW_1, ...]
__host__ __device__ __noinline__ int dw_phase_jobs(const PassArgs& a, int rows) {
    int chunks, S, cps;
    split_plan(rows, chunks, S, cps);
    int n = kDomJobs + 1;
    for (int l = 0; l < a.L; ++l) n += dw_tiles(a, l) * S;
    return n;
}

__device__ __noinline__ Job decode_dw_job(const PassArgs& a, int rows, int j) {
    Job J;
    J.z = 0; J.c_beg = 0; J.tiles = 0; J.NT = 1; J.n_tile = 0; J.layer = 0; J.bn = 0; J.nch = 0; J.m_tile = 0; J.gtile = 0; J.S = 1;
    if (j < kDomJobs) { J.type = J_DOM; J.m_tile = j; return J; }
    if (j == kDomJobs) { J.type = J_RED; return J; }
    int jj = j - kDomJobs - 1;
    int chunks, S, cps;
    split_plan(rows, chunks, S, cps);
    int l = 0;
    for (; l < a.L - 1; ++l) {
        const int cnt = dw_tiles(a, l) * S;
        if (jj < cnt) break;
        jj -= cnt;
    }
    J.type = J_DW;
    J.layer = l;
    J.S = S;
    J.gtile = 0;
    for (int l2 = 0; l2 < l; ++l2) J.gtile += dw_tiles(a, l2);
    J.bn = dw_bn(a, l);
    J.NT = a.n[l + 1] / J.bn;
    J.tiles = cdiv(a.n[l], 128) * J.NT;
    J.z = jj / J.tiles;
    const int t = jj - J.z * J.tiles;
    J.m_tile = t / J.NT;
    J.n_tile = t - J.m_tile * J.NT;
    J.gtile += t;
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
// 3xTF32 split of x: hi = rn_tf32(x) (the tensor core's own truncation of a raw fp32 operand would leave |lo| < 2^-11 |x| and
// a dropped lo.lo term of 2^-22; rounding hi to NEAREST halves |lo|, so the dropped term is 2^-24: fp32-level products),
// lo = rn_tf32(x - hi) (x - hi is exact in fp32)
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = rn_tf32(x);
    lo = rn_tf32(x - hi);
}

// store 4 consecutive elements of a pair array: plane 0 = hi (the pre-rounded value in both tensor-core modes), plane 1 = lo
// (3-pass mode only)
__device__ __forceinline__ void store_pair4(float* hi, long long z, float4 v, bool rnd, bool x3) {
    const float4 h = (rnd || x3) ? rn_tf32_4(v) : v;
    *reinterpret_cast<float4*>(hi) = h;
    if (x3) *reinterpret_cast<float4*>(hi + z) = rn_tf32_4(make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w));
}
__device__ __forceinline__ void store_pair1(float* hi, long long z, float v, bool rnd, bool x3) {
    const float h = (rnd || x3) ? rn_tf32(v) : v;
    *hi = h;
    if (x3) hi[z] = rn_tf32(v - h);
}

__device__ __forceinline__ void adam1(float& p, float& m, float& v, float g, float alpha, float omb1, float omb2, float eps) {
    m = __fadd_rn(m, __fmul_rn(__fsub_rn(g, m), omb1));
    v = __fadd_rn(v, __fmul_rn(__fsub_rn(__fmul_rn(g, g), v), omb2));
    p = __fsub_rn(p, __fdiv_rn(__fmul_rn(m, alpha), __fadd_rn(__fsqrt_rn(v), eps)));
}

// Sums over the 32 lanes of a warp of 8 per-lane values, in a fixed butterfly order: afterwards every lane holds the
// total of value index ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1).  9 shuffles.
__device__ __forceinline__ float warp_sum8(float (&v)[8], int lane) {
#pragma unroll
    for (int step = 0; step < 3; ++step) {
        const int off = 16 >> step, half = 4 >> step;
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float send = up ? v[i] : v[i + half];
            const float recv = __shfl_xor_sync(0xffffffffu, send, off);
            v[i] = (up ? v[i + half] : v[i]) + recv;
        }
    }
    float r = v[0];
    r += __shfl_xor_sync(0xffffffffu, r, 2);
    r += __shfl_xor_sync(0xffffffffu, r, 1);
    return r;
}
__device__ __forceinline__ int warp_sum8_index(int lane) { return ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1); }

// fixed-order sum over the row groups g = 0 .. ng-1 of p[g * stride] by one warp (lane-strided, then a butterfly)
__device__ __forceinline__ float warp_group_sum_f(const float* p, int ng, long long stride, int lane) {
    float s = 0.f;
    for (int g0 = 0; g0 < ng; g0 += 128) {
        float q[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int g = g0 + u * 32 + lane; q[u] = g < ng ? ldcg_f(p + g * stride) : 0.f; }
#pragma unroll
        for (int u = 0; u < 4; ++u) s += q[u];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
}
__device__ __forceinline__ double warp_group_sum_d(const double* p, int ng, int lane) {
    double s = 0.0;
    for (int g = lane; g < ng; g += 32) s += __ldcg(p + g);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
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
        if (lane == 0) {
            y[r] = __ldg(pd.label + o);
            if (a.bid[0]) { a.bid[0][r] = (int32_t)u; a.bid[1][r] = (int32_t)p; }
        }
    }
}

// k rows [kb, ke) of the domain block handled by domain job q
__device__ __forceinline__ void dom_rows(const PassArgs& a, int q, int& kb, int& ke) {
    const int per = cdiv(a.dd, kDomJobs);
    kb = q * per;
    ke = kb + per < a.dd ? kb + per : a.dd;
    if (kb > ke) kb = ke;
}

// fold partial of domain job q from the CURRENT parameters: fold_part[q][c] = sum_{k in rows(q)} E_d[dom][k] W_0[K0 + k, c]
// (pass prologue; inside a training pass the domain jobs publish it from the values they have just updated -- same
// operands, same order, same bits)
__device__ __forceinline__ void fold_from_params(const PassArgs& a, int dom, int q, int tid) {
    int kb, ke;
    dom_rows(a, q, kb, ke);
    const int n1 = a.n[1], K0 = a.n[0];
    const float* ed = a.params + a.off_Ed + (long long)dom * a.dd;
    for (int c = tid; c < n1; c += kWorkers) {
        const float* w = a.params + a.off_W[0] + (long long)(K0 + kb) * n1 + c;
        float fp = 0.f;
        for (int k0 = 0; k0 < ke - kb; k0 += 16) {
            float wv[16], ev[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const bool in = k0 + u < ke - kb;
                wv[u] = in ? ldcg_f(w + (long long)(k0 + u) * n1) : 0.f;
                ev[u] = in ? ldcg_f(ed + kb + k0 + u) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 16; ++u)
                if (k0 + u < ke - kb) fp = fmaf(ev[u], wv[u], fp);
        }
        a.fold_part[q * n1 + c] = fp;
    }
}

// ---- the kernel ---------------------------------------------------------------------------------------------------
template <int PASSES>   // 3 = 3xTF32 (pair operands), 1 = 1-pass TF32 (operands pre-rounded)
__global__ void __launch_bounds__(kThreads, 1)
pass_kernel(const __grid_constant__ MapTable maps, const __grid_constant__ PassArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar_full[kStages], bar_empty[kStages], bar_done, bar_tfree, bar_x, bar_acc[2], bar_epi[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ float s_thr[kMaxThr];
    __shared__ SegD s_seg[kMaxSegs];
    __shared__ int s_nfwd, s_nseg;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = gridDim.x, cta = blockIdx.x;
    constexpr bool x3 = PASSES == 3;    // 3xTF32: pair operands, N-concatenated MMAs
    constexpr bool rnd = PASSES == 1;   // 1-pass TF32: GEMM operands are stored pre-rounded (RN) to tf32
    unsigned char* scratch = smem + (size_t)kStages * STAGE_BYTES;
    const int L = a.L;

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            tc::mbar_init(&bar_full[s], 1);
            tc::mbar_init(&bar_empty[s], 1);
        }
        tc::mbar_init(&bar_done, 1);
        tc::mbar_init(&bar_tfree, kWorkers);
        tc::mbar_init(&bar_x, 1);
        for (int s = 0; s < 2; ++s) {
            tc::mbar_init(&bar_acc[s], 1);
            tc::mbar_init(&bar_epi[s], kWorkers);
        }
        tc::fence_barrier_init();
        // the GEMMs of the chain, in issue order
        int ns = 0, last = -1;
        for (int l = 0; l < L; ++l) {
            const int tiles = cdiv(a.n[l + 1], 128);
            for (int t = 0; t < tiles; ++t) {
                SegD sd;
                sd.kind = l == L - 1 ? S_HEAD : S_FWD; sd.layer = l; sd.mtile = t; sd.nch = a.n[l] / KCH;
                sd.ngrp = (a.n[l + 1] - t * 128) / 32 < 4 ? (a.n[l + 1] - t * 128) / 32 : 4;
                sd.dep = last; sd.bsrc = l & 1; sd.bdst = (l + 1) & 1; sd.mseg = ns;
                s_seg[ns++] = sd;
            }
            last = ns - 1;
        }
        s_nfwd = ns;
        int bsel = L & 1;   // the head writes dZ_{L-1} here
        for (int l = L - 1; l >= 1; --l) {
            const int tiles = cdiv(a.n[l], 128);
            // forward segments of layer l-1 (their features are this layer's inputs), in tile order
            int fseg = 0;
            for (int q = 0; q < s_nfwd; ++q)
                if (s_seg[q].layer == l - 1) { fseg = q; break; }
            const int prev_last = last;
            for (int t = 0; t < tiles; ++t) {
                SegD sd;
                sd.kind = S_DH; sd.layer = l; sd.mtile = t; sd.nch = a.n[l + 1] / KCH;
                sd.ngrp = (a.n[l] - t * 128) / 32 < 4 ? (a.n[l] - t * 128) / 32 : 4;
                sd.dep = prev_last; sd.bsrc = bsel; sd.bdst = bsel ^ 1; sd.mseg = fseg + t;
                s_seg[ns++] = sd;
            }
            last = ns - 1;
            bsel ^= 1;
        }
        if (a.emb_trainable) {   // dX^T[k, row] = W_0[k, :] . dZ_0[row, :]  (k < K0)
            const int tiles = cdiv(a.n[0], 128);
            const int prev_last = last;
            for (int t = 0; t < tiles; ++t) {
                SegD sd;
                sd.kind = S_DX; sd.layer = 0; sd.mtile = t; sd.nch = a.n[1] / KCH;
                sd.ngrp = (a.n[0] - t * 128) / 32 < 4 ? (a.n[0] - t * 128) / 32 : 4;
                sd.dep = prev_last; sd.bsrc = bsel; sd.bdst = bsel ^ 1; sd.mseg = 0;
                s_seg[ns++] = sd;
            }
        }
        s_nseg = ns;
    }
    if (warp == kMmaWarp) tc::tmem_alloc(&tmem_base_s, kTmemCols);
    if (warp == kProdWarp && lane == 0) {
        for (int b = 0; b < 2; ++b) { tc::tma_prefetch_desc(&maps.xk[b]); tc::tma_prefetch_desc(&maps.xmn[b]); }
        for (int l = 0; l < L; ++l) {
            tc::tma_prefetch_desc(&maps.wf[l]);
            tc::tma_prefetch_desc(&maps.dzmn[l]);
            if (l >= 1) { tc::tma_prefetch_desc(&maps.hmn[l]); tc::tma_prefetch_desc(&maps.wb[l]); }
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
    unsigned int tile_target = 0;   // split-K partials every dW tile has received so far in this launch
    int ring_s = 0;            // pipeline stage cursor of this warp's role (producer / MMA issuer)
    uint32_t ring_ph = 0;      // phase bit of the current lap
    uint32_t ring_n = 0;       // chunks handled so far (the first kStages need no empty-wait)
    uint32_t njob = 0;         // dW jobs run so far by this CTA (parity of bar_done / bar_tfree)
    uint32_t cseg = 0;         // chain segments run so far by this CTA (slot / parity of bar_acc / bar_epi)
    uint32_t waited = 0;       // MMA warp: segment epilogues consumed so far
    uint32_t xjobs = 0;        // chain jobs run so far by this CTA (parity of bar_x)
    const int K0 = a.n[0];
    const int NL = a.n[L];
    const float inv_keep = a.dropout_enabled && a.train ? a.dropout_scale : 1.0f;
    constexpr int nz = x3 ? 2 : 1;

    // the GEMM shadows of 4 consecutive elements W_l[k, c .. c+3] (arena offset o): the pair copy of W_l itself (M operand of
    // dH_l, l >= 1) and the pair copy of W_l^T (M operand of the forward GEMM; layer 0: rows k < K0 only)
    auto store_shadows = [&](int l, int k, int c, long long o, float4 w4) {
        if (l >= 1 || (a.emb_trainable && k < a.n[0])) store_pair4(a.wpair + o, a.wz, w4, rnd, x3);
        const int K = a.n[l];
        if (k < K) {
            float* t = a.wT[l] + (long long)c * K + k;
            const long long tz = (long long)a.n[l + 1] * K;
            store_pair1(t, tz, w4.x, rnd, x3);
            store_pair1(t + K, tz, w4.y, rnd, x3);
            store_pair1(t + 2 * K, tz, w4.z, rnd, x3);
            store_pair1(t + 3 * K, tz, w4.w, rnd, x3);
        }
    };
    // refresh the shadows from the parameter arena (float4 items of this thread)
    auto refresh_wpair = [&]() {
        for (int q = 0; q < a.nseg; ++q) {
            if (a.seg[q].kind != SEG_KERNEL) continue;
            const long long o0 = a.seg[q].off;
            const int l = a.seg[q].layer, N = a.n[l + 1];
            for (int i = (cta * kThreads + tid) * 4; i < a.seg[q].numel; i += G * kThreads * 4) {
                const int k = i / N;
                store_shadows(l, k, i - k * N, o0 + i, ldcg_f4(a.params + o0 + i));
            }
        }
    };

    const ProgOp* ops = a.ops ? a.ops : &a.inline_op;
    int pass_idx = 0;
    bool shadows_valid = false;   // the GEMM shadows match the arena: true after a training pass, until a meta sweep runs
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
        shadows_valid = false;   // (conservative: the sweep may have rewritten the live arena)
        continue;
    }
    PassDyn pd;
    pd.dom = op.domain; pd.steps = op.steps; pd.n_data = op.n_data;
    pd.uid = op.uid; pd.pid = op.pid; pd.order = op.order; pd.label = op.label;
    pd.losses = op.losses; pd.probs = op.probs;
    int* const hist_cur = a.hist + (pass_idx & 1) * (2 * (kMaxThr + 1));

    // ---- prologue of a pass: gather mini-batch 0; pair shadow of the kernels (the arena may have been rewritten by a
    // meta sweep or by the host since the last pass); |E_d|^2 for the inference loss
    if (!shadows_valid) refresh_wpair();
    if (warp < kWorkerWarps) {
        if (cta < kDomJobs) fold_from_params(a, pd.dom, cta, tid);
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

    const int n_phases = a.train ? 2 : 1;
    // the dW job of this CTA for a full mini-batch, decoded once (integer divisions off the per-step path)
    const int njobs_full = dw_phase_jobs(a, a.bs);
    const Job job_full = decode_dw_job(a, a.bs, cta);
    for (int step = 0; step < pd.steps; ++step) {
        const long long left = pd.n_data - (long long)step * a.bs;
        const int rows = left < a.bs ? (int)left : a.bs;
        // row groups of the chain: whole 32-row K chunks of the dW GEMMs when training (rows past the batch give zero dZ)
        const int ngroups = a.train ? cdiv(rows, KCH) * (KCH / CR) : cdiv(rows, CR);
        const int buf = step & 1;
        DropoutParams dp;
        dp.enabled = a.dropout_enabled && a.train;
        dp.threshold = a.dropout_threshold;
        dp.scale = a.dropout_scale;
        dp.step = (uint32_t)(step_ctr & 0xffffffffll);
        dp.seed = a.dropout_seed;
        dp.row0 = 0;

        const float alpha = __fdiv_rn(__fmul_rn(a.lr, __fsqrt_rn(__fsub_rn(1.0f, b2pow))), __fsub_rn(1.0f, b1pow));
        const float omb1 = __fsub_rn(1.0f, a.beta1), omb2 = __fsub_rn(1.0f, a.beta2);
        // optimizer apply on 4 consecutive parameters at arena offset o with gradient g (TF ApplyAdam order / plain SGD)
        auto apply4 = [&](long long o, const float (&g)[4], int kl, int kk, int kc) {   // kl >= 0: element [kk, kc..] of kernel kl
            const float4 P = ldcg_f4(a.params + o);
            float pp[4] = {P.x, P.y, P.z, P.w};
            if (a.opt_kind == 0) {
                const float4 M = ldcg_f4(a.m + o), V = ldcg_f4(a.v + o);
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
            if (kl >= 0) store_shadows(kl, kk, kc, o, pnew);
            if (a.grads) *reinterpret_cast<float4*>(a.grads + o) = make_float4(g[0], g[1], g[2], g[3]);
        };
        auto apply1 = [&](long long o, float g) {
            float pe = ldcg_f(a.params + o);
            if (a.opt_kind == 0) {
                float me = ldcg_f(a.m + o), ve = ldcg_f(a.v + o);
                adam1(pe, me, ve, g, alpha, omb1, omb2, a.eps);
                a.m[o] = me; a.v[o] = ve;
            } else {
                pe = __fsub_rn(pe, __fmul_rn(g, a.lr));
            }
            a.params[o] = pe;
            if (a.grads) a.grads[o] = g;
            return pe;
        };

        for (int phase = 0; phase < n_phases; ++phase) {
            const long long tslot = (((long long)step * n_phases + phase) * G + cta) * 16;
            const bool tim = a.timing && tslot + 15 < a.timing_cap;
            if (tim && tid == 0) { a.timing[tslot] = gtime(); a.timing[tslot + 7] = (unsigned long long)clock64(); }
            if (phase == 0) {
                // =========================================================================================== chain
                const int nsegs = a.train ? s_nseg : s_nfwd;
                for (int j = cta; j < ngroups; j += G) {
                    const int row0 = j * CR;
                    if (warp == kProdWarp) {
                        // ---------- TMA producer: the group's X rows, then every weight tile of the chain in issue order
                        if (tim && lane == 0) a.timing[tslot + 6] = (unsigned long long)clock64();
                        if (tc::elect_one()) {
                            const int nch0 = K0 / KCH;
                            tc::mbar_arrive_expect_tx(&bar_x, (uint32_t)(nch0 * CR * KCH * 4 * nz));
                            for (int c = 0; c < nch0; ++c)   // box = [plane][CR rows][32 k]: hi rows, then lo rows
                                tc::tma_load_3d(smem + kRingBytes + c * kBChunk, &maps.xk[buf], &bar_x, c * KCH, row0, 0);
                        }
                        __syncwarp();
                        for (int s = 0; s < nsegs; ++s) {
                            const SegD sd = s_seg[s];
                            const int l = sd.layer;
                            const bool fwd = sd.kind == S_FWD || sd.kind == S_HEAD;
                            const int mw = fwd ? a.n[l + 1] : a.n[l];       // M extent of the weight operand
                            const int brow = mw < 128 ? mw : 128;           // box rows of its map
                            const uint32_t tx = (uint32_t)(brow * KCH * 4) * (uint32_t)nz;
                            const CUtensorMap* mp = fwd ? &maps.wf[l] : &maps.wb[l];
                            for (int i0 = 0; i0 < sd.nch; ++i0) {
                                const int i = (i0 + j) % sd.nch;   // chunk order rotated per row group (see the MMA warp)
                                if (ring_n >= (uint32_t)kStages) tc::mbar_wait(&bar_empty[ring_s], ring_ph ^ 1);
                                unsigned char* sA = smem + (size_t)ring_s * kChainStage;
                                uint64_t* fb = &bar_full[ring_s];
                                if (tc::elect_one()) {
                                    tc::mbar_arrive_expect_tx(fb, tx);
                                    tc::tma_load_3d(sA, mp, fb, i * KCH, sd.mtile * 128, 0);   // one box = [plane][brow rows][32 k]
#ifdef PASS_DBG_SEG
                                    if (tim && s == PASS_DBG_SEG) {
                                        if (i0 == 0) a.timing[tslot + 2] = (unsigned long long)clock64();
                                        if (i0 == 2) a.timing[tslot + 3] = (unsigned long long)clock64();
                                        if (i0 == 4) a.timing[tslot + 4] = (unsigned long long)clock64();
                                        if (i0 == 6) a.timing[tslot + 5] = (unsigned long long)clock64();
                                        if (i0 == 7) a.timing[tslot + 6] = (unsigned long long)clock64();
                                    }
#else
                                    if (tim) {
                                        if (s == 0 && i0 == 0) a.timing[tslot + 2] = (unsigned long long)clock64();
                                        if (s == nsegs - 1 && i0 == sd.nch - 1) a.timing[tslot + 3] = (unsigned long long)clock64();
                                    }
#endif
                                }
                                __syncwarp();
                                ++ring_n;
                                if (++ring_s == kStages) { ring_s = 0; ring_ph ^= 1; }
                            }
                        }
                    } else if (warp == kMmaWarp) {
                        // ---------- MMA issuer: out^T[feature, row] (+)= W^T . act^T, accumulators alternate between two TMEM slots
                        for (int s = 0; s < nsegs; ++s) {
                            const SegD sd = s_seg[s];
                            const uint32_t cs = cseg + (uint32_t)s;
                            // the slot must be drained (epilogue of segment cs - 2) and the N operand written (epilogue of `dep`)
                            uint32_t need = cs >= 1 ? cs - 1 : 0;
                            if (sd.dep >= 0 && cseg + (uint32_t)sd.dep + 1 > need) need = cseg + (uint32_t)sd.dep + 1;
                            while (waited < need) { tc::mbar_wait(&bar_epi[waited & 1], (waited >> 1) & 1); ++waited; }
                            if (s == 0) tc::mbar_wait(&bar_x, xjobs & 1);
                            tc::tc_fence_after();
                            const bool fwd = sd.kind == S_FWD || sd.kind == S_HEAD;
                            constexpr uint32_t idN2 = tc::make_idesc_tf32(128, 2 * CR, 0, 0);
                            constexpr uint32_t idN1 = tc::make_idesc_tf32(128, CR, 0, 0);
                            constexpr uint32_t a_hiw = tc::kDescHiK, a_low = tc::kDescLoK;
                            constexpr uint32_t a_k = 32u >> 4;   // per k-step of 8
                            const int mw = fwd ? a.n[sd.layer + 1] : a.n[sd.layer];
                            const uint32_t lo_off = (uint32_t)((mw < 128 ? mw : 128) * KCH * 4) >> 4;   // the lo plane follows the hi box
                            const uint32_t bbase = ((smem_base + (uint32_t)kRingBytes + (uint32_t)sd.bsrc * kActBytes) >> 4) | tc::kDescLoK;
                            const uint32_t dslot = tmem + (cs & 1u) * kSlotCols;
                            for (int i0 = 0; i0 < sd.nch; ++i0) {
                                // every CTA streams the same weights: starting the K loop at chunk j mod nch keeps the CTAs from
                                // hammering the same L2 lines in lock-step (the K order of a row group is fixed: deterministic)
                                const int i = (i0 + j) % sd.nch;
                                tc::mbar_wait(&bar_full[ring_s], ring_ph);
                                tc::tc_fence_after();
                                const uint32_t st = (smem_base + (uint32_t)ring_s * kChainStage) >> 4;
                                if (tc::elect_one()) {
#ifdef PASS_DBG_SEG
                                    if (tim && s == PASS_DBG_SEG && i0 < 8) a.timing[tslot + 8 + i0] = (unsigned long long)clock64();
#else
                                    if (tim && s == 0 && i0 == 0) a.timing[tslot + 4] = (unsigned long long)clock64();
#endif
                                    const uint32_t aw = st | a_low, alw = aw + lo_off;
                                    const uint32_t bw = bbase + (uint32_t)i * (kBChunk >> 4);
                                    // chunk i0 accumulates into group i0 mod kAccGroups of the slot
                                    const uint32_t dcol = dslot + (uint32_t)(i0 & (kAccGroups - 1)) * (2 * CR);
                                    uint32_t acc = i0 < kAccGroups ? 0u : 1u;
                                    if (x3) {
#pragma unroll
                                        for (int k = 0; k < KCH / 8; ++k) {
                                            // W.[act | act_lo] -> columns [0, CR) and [CR, 2 CR);  W_lo.act -> columns [CR, 2 CR) as well:
                                            // the first half only ever sees the large hi.hi products
                                            tc::mma_tf32(dcol, tc::desc_words(aw + k * a_k, a_hiw), tc::desc_words(bw + k * 2u, tc::kDescHiK), idN2, acc);
                                            tc::mma_tf32(dcol + CR, tc::desc_words(alw + k * a_k, a_hiw), tc::desc_words(bw + k * 2u, tc::kDescHiK), idN1, 1u);
                                            acc = 1;
                                        }
                                    } else {
#pragma unroll
                                        for (int k = 0; k < KCH / 8; ++k) {
                                            tc::mma_tf32(dcol, tc::desc_words(aw + k * a_k, a_hiw), tc::desc_words(bw + k * 2u, tc::kDescHiK), idN1, acc);
                                            acc = 1;
                                        }
                                    }
                                    tc::mma_commit(&bar_empty[ring_s]);
                                }
                                __syncwarp();
                                if (++ring_s == kStages) { ring_s = 0; ring_ph ^= 1; }
                            }
                            if (tc::elect_one()) {
                                tc::mma_commit(&bar_acc[cs & 1]);
#ifndef PASS_DBG_SEG
                                if (tim && s == nsegs - 1) a.timing[tslot + 5] = (unsigned long long)clock64();
#endif
                            }
                            __syncwarp();
                        }
                        // every epilogue of the job has run: the accumulators and the activation buffers are free again
                        while (waited < cseg + (uint32_t)nsegs) { tc::mbar_wait(&bar_epi[waited & 1], (waited >> 1) & 1); ++waited; }
                    } else {
                        // ---------- workers: thread <-> one output feature (TMEM lane) x 8 rows
                        const int q = warp & 3, hf = warp >> 2;
                        const int wt = q * 32 + lane;                    // feature inside the M tile
                        const int r_first = row0 + hf * 8;               // batch row of this thread's value 0
                        unsigned char* s_mask = scratch + kScrMask;
                        float* s_zp = reinterpret_cast<float*>(scratch + kScrZp);
                        float* s_ds = reinterpret_cast<float*>(scratch + kScrDs);
                        float* s_comb = reinterpret_cast<float*>(scratch + kScrComb);
                        const float dscale = dp.enabled ? dp.scale : 1.0f;
                        for (int s = 0; s < nsegs; ++s) {
                            const SegD sd = s_seg[s];
                            const uint32_t cs = cseg + (uint32_t)s;
                            const int l = sd.layer;
                            const int f = sd.mtile * 128 + wt;            // output feature of this thread
                            const bool on = q < sd.ngrp;                  // warp-uniform: the tile has this lane quarter
                            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (cs & 1u) * kSlotCols + (uint32_t)(hf * 8);
                            // this thread's 8 values of the N operand chunk f / 32: row line hf * 8 + i, swizzled 16-byte chunk
                            unsigned char* bdst = smem + kRingBytes + sd.bdst * kActBytes + (f >> 5) * kBChunk + (hf * 8) * 128 + (lane & 3) * 4;
                            float bias = 0.f, wd = 0.f, gbias = 0.f, ylab = 0.f;
                            uint32_t keep = 0xffu;
                            if (sd.kind == S_FWD || sd.kind == S_HEAD) {
                                // ---- prelude in the shadow of the MMAs: effective bias, dropout keep bits
                                if (l == 0) {
                                    // layer 0 adds E_d[dom] . W_0[K0:, f] in fp32: the k-slice partials published by the domain jobs
                                    // of the previous update (or by the pass prologue)
                                    if (on) {
                                        float q8[kDomJobs];
#pragma unroll
                                        for (int u = 0; u < kDomJobs; ++u) q8[u] = ldcg_f(a.fold_part + u * a.n[1] + f);
                                        float fs = 0.f;
#pragma unroll
                                        for (int u = 0; u < kDomJobs; ++u) fs += q8[u];
                                        bias = ldcg_f(a.params + a.off_b[0] + f) + fs;
                                    }
                                } else if (on) {
                                    bias = ldcg_f(a.params + a.off_b[l] + f);
                                }
                                if (sd.kind == S_HEAD) {
                                    if (on) wd = ldcg_f(a.params + a.off_w + f);
                                    if (warp == 0) {   // the row lanes of the sigmoid-BCE head: label and global bias up front
                                        gbias = ldcg_f(a.params + a.off_g);
                                        if (lane < CR && row0 + lane < rows) ylab = a.y[buf][row0 + lane];
                                    }
                                }
                                if (dp.enabled && on) {
                                    // the 4 lanes of a feature quad share one Philox counter per row: each lane draws two rows' words
                                    DropoutParams dq = dp;
                                    dq.seed = a.dropout_seed + (uint32_t)l;
                                    const uint32_t N = (uint32_t)a.n[l + 1];
                                    uint32_t nib[2];
#pragma unroll
                                    for (int u = 0; u < 2; ++u) {
                                        const uint32_t row = (uint32_t)(r_first + (lane & 3) + 4 * u);
                                        const uint4 w = dropout_words4(dq, row * N + (uint32_t)(f & ~3));
                                        nib[u] = (w.x < dq.threshold ? 1u : 0u) | (w.y < dq.threshold ? 2u : 0u) |
                                                 (w.z < dq.threshold ? 4u : 0u) | (w.w < dq.threshold ? 8u : 0u);
                                    }
                                    keep = 0u;
#pragma unroll
                                    for (int i = 0; i < 8; ++i) {
                                        const uint32_t nb = __shfl_sync(0xffffffffu, nib[i >> 2], (lane & ~3) | (i & 3));
                                        keep |= ((nb >> (lane & 3)) & 1u) << i;
                                    }
                                }
                            }
                            tc::mbar_wait_warp(&bar_acc[cs & 1], (cs >> 1) & 1);
                            tc::tc_fence_after();
#ifdef PASS_DBG_HEAD
                            if (sd.kind == S_HEAD) WSTAMP(8);
#endif
                            float vv[8], v2[8];
                            if (on) {
                                // sum of the accumulator groups, (g0 + g1) + (g2 + g3), with round-to-nearest adds
                                const int ng = sd.nch < kAccGroups ? sd.nch : kAccGroups;
                                float t0[8], t1[8];
#pragma unroll
                                for (int part = 0; part < (x3 ? 2 : 1); ++part) {
                                    float* dst = part == 0 ? vv : v2;
                                    const uint32_t ta = taddr + (part == 0 ? 0u : (uint32_t)CR);
                                    tc::tmem_ld8(ta, dst);
                                    if (ng > 1) {
                                        tc::tmem_ld8(ta + 2 * CR, t0);
#pragma unroll
                                        for (int i = 0; i < 8; ++i) dst[i] += t0[i];
                                    }
                                    if (ng > 2) {
                                        tc::tmem_ld8(ta + 4 * CR, t0);
                                        if (ng > 3) {
                                            tc::tmem_ld8(ta + 6 * CR, t1);
#pragma unroll
                                            for (int i = 0; i < 8; ++i) t0[i] += t1[i];
                                        }
#pragma unroll
                                        for (int i = 0; i < 8; ++i) dst[i] += t0[i];
                                    }
                                }
                            }
                            tc::tc_fence_before();
#ifdef PASS_DBG_HEAD
                            if (sd.kind == S_HEAD) WSTAMP(9);
#endif

                            if (sd.kind == S_FWD) {
                                // H_{l+1}[row, f] = dropout(relu(acc + b)) -> N operand of the next layer (+ global copy for dW)
                                float h[8];
                                if (on) {
                                    uint32_t mb = 0u;
#pragma unroll
                                    for (int i = 0; i < 8; ++i) {
                                        const float accv = x3 ? vv[i] + v2[i] : vv[i];
                                        float hh = fmaxf(accv + bias, 0.f);
                                        hh = (keep >> i) & 1u ? hh * dscale : 0.f;
                                        h[i] = hh;
                                        mb |= (hh > 0.f ? 1u : 0u) << i;
                                        unsigned char* d = bdst + i * 128 + ((((lane >> 2) ^ i) & 7) << 4);
                                        store_pair1(reinterpret_cast<float*>(d), CR * 32, hh, rnd, x3);
                                    }
                                    s_mask[sd.mseg * 256 + tid] = (unsigned char)mb;
                                }
                                tc::fence_proxy_async();
                                tc::mbar_arrive(&bar_epi[cs & 1]);
                                if (on && a.train) {
                                    const int N = a.n[l + 1];
                                    float* out = a.H[l + 1] + (long long)r_first * N + f;
                                    const long long oz = (long long)a.max_rows * N;
#pragma unroll
                                    for (int i = 0; i < 8; ++i)
                                        if (r_first + i < rows) store_pair1(out + (long long)i * N, oz, h[i], rnd, x3);
                                }
                            } else if (sd.kind == S_HEAD) {
                                // last hidden layer + Dense(1) + sigmoid + BCE + ds + dZ_{L-1} + column-sum partials + AUC bins
                                float h[8], zp[8];
#pragma unroll
                                for (int i = 0; i < 8; ++i) { h[i] = 0.f; zp[i] = 0.f; }
                                if (on) {
#pragma unroll
                                    for (int i = 0; i < 8; ++i) {
                                        const float accv = x3 ? vv[i] + v2[i] : vv[i];
                                        float hh = fmaxf(accv + bias, 0.f);
                                        hh = (keep >> i) & 1u ? hh * dscale : 0.f;
                                        h[i] = hh;
                                        zp[i] = hh * wd;
                                    }
                                    const float zs = warp_sum8(zp, lane);
                                    if ((lane & 3) == 0) s_zp[(hf * 4 + q) * 8 + warp_sum8_index(lane)] = zs;
                                }
#ifdef PASS_DBG_HEAD
                                WSTAMP(10);
#endif
                                worker_sync();
#ifdef PASS_DBG_HEAD
                                WSTAMP(11);
#endif
                                float pv = 0.f, yv = 0.f, dsv0 = 0.f;   // warp 0: row lane of the group
                                const int hrow = row0 + lane;
                                const bool hvalid = warp == 0 && lane < CR && hrow < rows;
                                if (warp == 0) {
                                    if (hvalid) {
                                        float z = 0.f;
                                        for (int qq = 0; qq < sd.ngrp; ++qq) z += s_zp[((lane >> 3) * 4 + qq) * 8 + (lane & 7)];
                                        const float sgm = z + gbias;
                                        pv = 1.0f / (1.0f + expf(-sgm));
                                        yv = ylab;
                                        if (a.train) dsv0 = (fabsf(sgm) <= MAMDR_LOGIT_CLIP) ? __fdiv_rn(__fsub_rn(pv, yv), (float)rows) : 0.f;
                                    }
                                    if (lane < CR) s_ds[lane] = dsv0;
                                }
#ifdef PASS_DBG_HEAD
                                WSTAMP(12);
#endif
                                worker_sync();
#ifdef PASS_DBG_HEAD
                                WSTAMP(13);
#endif
                                float hd = 0.f, dbs = 0.f;
                                float dz[8];
                                if (on && a.train) {
#pragma unroll
                                    for (int i = 0; i < 8; ++i) {
                                        const float dsv = s_ds[hf * 8 + i];
                                        const float dh = __fmul_rn(dsv, wd);
                                        dz[i] = h[i] > 0.f ? __fmul_rn(dh, inv_keep) : 0.f;   // rows past the batch: ds = 0
                                        hd += h[i] * dsv;
                                        dbs += dz[i];
                                        unsigned char* d = bdst + i * 128 + ((((lane >> 2) ^ i) & 7) << 4);
                                        store_pair1(reinterpret_cast<float*>(d), CR * 32, dz[i], rnd, x3);
                                    }
                                }
                                tc::fence_proxy_async();
                                tc::mbar_arrive(&bar_epi[cs & 1]);
#ifdef PASS_DBG_HEAD
                                WSTAMP(14);
#endif
                                if (warp == 0) {
                                    // off the critical path: Keras BCE of the clipped probability, probabilities, AUC bins
                                    const float lo_c = 1e-7f, hi_c = 1.0f - 1e-7f;
                                    double bce = 0.0;
                                    if (hvalid) {
                                        const float ph = fminf(fmaxf(pv, lo_c), hi_c);
                                        const float lg = logf(ph / (1.0f - ph));
                                        bce = (double)(fmaxf(lg, 0.f) - lg * yv + log1pf(expf(-fabsf(lg))));
                                        if (pd.probs) pd.probs[(long long)step * a.bs + hrow] = pv;
                                        if (a.auc_acc) {
                                            int lo_i = 0, hi_i = a.T;
                                            while (lo_i < hi_i) {
                                                const int mid = (lo_i + hi_i) >> 1;
                                                if (s_thr[mid] < pv) lo_i = mid + 1; else hi_i = mid;
                                            }
                                            atomicAdd(&hist_cur[(yv != 0.f ? (a.T + 1) : 0) + lo_i], 1);
                                        }
                                    }
                                    double bs = bce;
                                    float dgs = dsv0;
#pragma unroll
                                    for (int o = 16; o > 0; o >>= 1) {
                                        bs += __shfl_xor_sync(0xffffffffu, bs, o);
                                        dgs += __shfl_xor_sync(0xffffffffu, dgs, o);
                                    }
                                    if (lane == 0) { a.loss_part[buf * kMaxGroups + j] = bs; a.dg_part[j] = dgs; }
                                }
                                if (a.train) {
                                    if (on) {
                                        float* out = a.dZ[L - 1] + (long long)r_first * NL + f;
                                        const long long oz = (long long)a.max_rows * NL;
#pragma unroll
                                        for (int i = 0; i < 8; ++i) store_pair1(out + (long long)i * NL, oz, dz[i], rnd, x3);
                                    }
                                    // the two row halves in order -> per-group column sums (Dense(1) kernel gradient, db_{L-1})
                                    float* comb = s_comb + (cs & 1u) * 256;
                                    if (on && hf == 1) { comb[wt] = hd; comb[128 + wt] = dbs; }
                                    worker_sync();
                                    if (on && hf == 0) {
                                        a.dw_part[(long long)j * NL + f] = hd + comb[wt];
                                        a.db_part[L - 1][(long long)j * NL + f] = dbs + comb[128 + wt];
                                    }
                                }
                            } else if (sd.kind == S_DX) {
                                // dX[row, k] (fp32, plain): gradient rows of the gathered user / item embeddings
                                tc::mbar_arrive(&bar_epi[cs & 1]);
                                if (on) {
                                    float* out = a.dX + (long long)r_first * K0 + f;
#pragma unroll
                                    for (int i = 0; i < 8; ++i)
                                        if (r_first + i < rows) out[(long long)i * K0] = x3 ? vv[i] + v2[i] : vv[i];
                                }
                            } else {
                                // dZ_{l-1}[row, f] = acc * inv_keep * 1[H_l > 0]; per-group column sums -> db_{l-1}
                                float dz[8];
                                float dbs = 0.f;
                                const bool feed = l >= 2 || a.emb_trainable;   // dZ_{l-1} is the N operand of dH_{l-1} (of the dX GEMM for l = 1)
                                if (on) {
                                    const uint32_t mb = s_mask[sd.mseg * 256 + tid];
#pragma unroll
                                    for (int i = 0; i < 8; ++i) {
                                        const float accv = x3 ? vv[i] + v2[i] : vv[i];
                                        dz[i] = (r_first + i < rows && ((mb >> i) & 1u)) ? accv * inv_keep : 0.f;
                                        dbs += dz[i];
                                        if (feed) {
                                            unsigned char* d = bdst + i * 128 + ((((lane >> 2) ^ i) & 7) << 4);
                                            store_pair1(reinterpret_cast<float*>(d), CR * 32, dz[i], rnd, x3);
                                        }
                                    }
                                }
                                tc::fence_proxy_async();
                                tc::mbar_arrive(&bar_epi[cs & 1]);
                                const int N = a.n[l];
                                if (on) {
                                    float* out = a.dZ[l - 1] + (long long)r_first * N + f;
                                    const long long oz = (long long)a.max_rows * N;
#pragma unroll
                                    for (int i = 0; i < 8; ++i) store_pair1(out + (long long)i * N, oz, dz[i], rnd, x3);
                                }
                                float* comb = s_comb + (cs & 1u) * 256;
                                if (on && hf == 1) comb[wt] = dbs;
                                worker_sync();
                                if (on && hf == 0) a.db_part[l - 1][(long long)j * N + f] = dbs + comb[wt];
                            }
#if !defined(PASS_DBG_SEG) && !defined(PASS_DBG_HEAD)
                            if (tim && tid == 0 && s < 8) a.timing[tslot + 8 + s] = (unsigned long long)clock64();
#endif
#ifdef PASS_DBG_HEAD
                            if (sd.kind == S_HEAD) WSTAMP(15);
#endif
                        }
                    }
                    cseg += (uint32_t)nsegs;
                    ++xjobs;
                    __syncthreads();   // the next job's X load overwrites activation buffer 0
                }
                // CTAs without a row group stage the next mini-batch meanwhile
                if (step + 1 < pd.steps && warp < kWorkerWarps) {
                    const int first = G - ngroups >= G / 4 ? ngroups : 0;
                    if (cta >= first) gather_rows(a, pd, step + 1, buf ^ 1, (cta - first) * kWorkerWarps + warp, (G - first) * kWorkerWarps, lane, rnd, x3);
                }
            } else if (phase == 1) {
                // =========================================================================================== dW
                const int njobs = rows == a.bs ? njobs_full : dw_phase_jobs(a, rows);
                for (int j = cta; j < njobs; j += G) {
                    const Job J = (rows == a.bs && j == cta) ? job_full : decode_dw_job(a, rows, j);
                    const int l = J.layer;
                    if (J.type == J_DOM) {
                        // ---------- domain job q of kDomJobs (workers): db_0 (every job, into smem); for the k rows of this job:
                        // dE_d[dom][k] = W_0dom[k, :] . db_0, the optimizer apply on E_d[dom][k] and on the rank-1 gradient
                        // E_d[dom][k] (x) db_0 of W_0dom[k, :], and the fold partial of the UPDATED values for the next chain.
                        // Job 0 also publishes db_0 and |E_d|^2 (pre-update, for the loss).
                        if (warp < kWorkerWarps) {
                            const int n1 = a.n[1];
                            float* s_db0 = reinterpret_cast<float*>(scratch);           // [n1 <= 256]
                            float4* s_sl = reinterpret_cast<float4*>(scratch + 1024);   // [slices][n1 / 4]
                            float* s_g = reinterpret_cast<float*>(scratch + 5120);      // [32] dE_d[dom][k]
                            float* s_eo = s_g + 32;                                     // [32] E_d[dom][k] before the apply
                            float* s_en = s_g + 64;                                     // [32] ... after
                            const int ncol4 = n1 >> 2, nsl = kWorkers / ncol4;
                            int kb, ke;
                            dom_rows(a, J.m_tile, kb, ke);
                            {
                                // column sums of the per-group partials: slice sl sums its groups in order, then the slices in order
                                const int c4 = tid % ncol4, sl = tid / ncol4;
                                const int per = cdiv(ngroups, nsl);
                                const int g_end = (sl + 1) * per < ngroups ? (sl + 1) * per : ngroups;
                                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                                if (sl < nsl) {
                                    for (int g0 = sl * per; g0 < g_end; g0 += 16) {
                                        float4 part[16];
#pragma unroll
                                        for (int u = 0; u < 16; ++u)
                                            part[u] = g0 + u < g_end ? ldcg_f4(a.db_part[0] + (long long)(g0 + u) * n1 + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                                        for (int u = 0; u < 16; ++u) { acc.x += part[u].x; acc.y += part[u].y; acc.z += part[u].z; acc.w += part[u].w; }
                                    }
                                    s_sl[sl * ncol4 + c4] = acc;
                                }
                                worker_sync();
                                if (tid < ncol4) {
                                    float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
                                    for (int s2 = 0; s2 < nsl; ++s2) { const float4 p4 = s_sl[s2 * ncol4 + tid]; t4.x += p4.x; t4.y += p4.y; t4.z += p4.z; t4.w += p4.w; }
                                    *reinterpret_cast<float4*>(s_db0 + tid * 4) = t4;
                                    if (J.m_tile == 0) {   // b_0
                                        const float g4[4] = {t4.x, t4.y, t4.z, t4.w};
                                        apply4(a.off_b[0] + tid * 4, g4, -1, 0, 0);
                                    }
                                }
                            }
                            worker_sync();
                            WSTAMP(11);
                            float* W0dom = a.params + a.off_W[0] + (long long)K0 * n1;
                            for (int k = kb + warp; k < ke; k += kWorkerWarps) {
                                const float* wr = W0dom + (long long)k * n1;
                                float sv = 0.f;
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
                                            sv = fmaf(d4.x, wv[u].x, sv); sv = fmaf(d4.y, wv[u].y, sv); sv = fmaf(d4.z, wv[u].z, sv); sv = fmaf(d4.w, wv[u].w, sv);
                                        }
                                    }
                                }
#pragma unroll
                                for (int o = 16; o > 0; o >>= 1) sv += __shfl_xor_sync(0xffffffffu, sv, o);
                                if (lane == 0) s_g[k - kb] = sv;
                            }
                            worker_sync();
                            if (tid < ke - kb) {
                                // E_d[dom][k]: g = 2 l2 E + dE   (the other rows of E_d take their l2-only gradient in the column-sum job)
                                const long long o = a.off_Ed + (long long)pd.dom * a.dd + kb + tid;
                                const float pe = ldcg_f(a.params + o);
                                s_eo[tid] = pe;
                                s_en[tid] = apply1(o, __fadd_rn(__fmul_rn(2.0f * a.l2_emb, pe), s_g[tid]));
                            }
                            WSTAMP(12);
                            worker_sync();
                            if (tid == 0) {   // this job's share of |E_d|^2 (pre-update values, for the loss)
                                double sq = 0.0;
                                for (int k = 0; k < ke - kb; ++k) sq += (double)s_eo[k] * (double)s_eo[k];
                                a.ed_sq[1 + J.m_tile] = sq;
                            }
                            for (int c = tid; c < n1; c += kWorkers) {
                                const long long o0 = a.off_W[0] + (long long)(K0 + kb) * n1 + c;
                                const float dbc = s_db0[c];
                                float fp = 0.f;
                                for (int k0 = 0; k0 < ke - kb; k0 += 16) {
                                    float pw[16], mw[16], vw[16];
#pragma unroll
                                    for (int u = 0; u < 16; ++u) {
                                        const bool in = k0 + u < ke - kb;
                                        const long long o = o0 + (long long)(k0 + u) * n1;
                                        pw[u] = in ? ldcg_f(a.params + o) : 0.f;
                                        mw[u] = in && a.opt_kind == 0 ? ldcg_f(a.m + o) : 0.f;
                                        vw[u] = in && a.opt_kind == 0 ? ldcg_f(a.v + o) : 0.f;
                                    }
#pragma unroll
                                    for (int u = 0; u < 16; ++u) {
                                        if (k0 + u < ke - kb) {
                                            const long long o = o0 + (long long)(k0 + u) * n1;
                                            const float g = __fmul_rn(s_eo[k0 + u], dbc);   // rank-1: E_d[dom]^T (x) db_0
                                            if (a.opt_kind == 0) {
                                                adam1(pw[u], mw[u], vw[u], g, alpha, omb1, omb2, a.eps);
                                                a.m[o] = mw[u]; a.v[o] = vw[u];
                                            } else {
                                                pw[u] = __fsub_rn(pw[u], __fmul_rn(g, a.lr));
                                            }
                                            a.params[o] = pw[u];
                                            if (a.grads) a.grads[o] = g;
                                            fp = fmaf(s_en[k0 + u], pw[u], fp);
                                        }
                                    }
                                }
                                a.fold_part[J.m_tile * n1 + c] = fp;
                            }
                            worker_sync();
                            WSTAMP(13);
                        }
                        continue;
                    }
                    if (J.type == J_RED) {
                        // ---------- column-sum job (workers): totals of the per-group partials of db_1.., the Dense(1) kernel
                        // gradient, the global-bias gradient (each applied right away) and the mini-batch loss; thread per column
                        if (warp < kWorkerWarps) {
                            int total = NL;
                            for (int l2 = 1; l2 < L; ++l2) total += a.n[l2 + 1];
                            for (int c = tid; c < total; c += kWorkers) {
                                // column c of [db_1 | db_2 | .. | d(Dense kernel)]
                                int cc = c, l2 = 1;
                                for (; l2 < L && cc >= a.n[l2 + 1]; ++l2) cc -= a.n[l2 + 1];
                                const int N = l2 < L ? a.n[l2 + 1] : NL;
                                const float* src = (l2 < L ? a.db_part[l2] : a.dw_part) + cc;
                                float sum = 0.f;
                                for (int g0 = 0; g0 < ngroups; g0 += 16) {
                                    float pv[16];
#pragma unroll
                                    for (int u = 0; u < 16; ++u) pv[u] = g0 + u < ngroups ? ldcg_f(src + (long long)(g0 + u) * N) : 0.f;
#pragma unroll
                                    for (int u = 0; u < 16; ++u) sum += pv[u];
                                }
                                apply1((l2 < L ? a.off_b[l2] : a.off_w) + cc, sum);
                            }
                            if (warp == 0) {
                                const float s = warp_group_sum_f(a.dg_part, ngroups, 1, lane);
                                if (lane == 0) apply1(a.off_g, s);
                            }
                            // the rows of E_d other than the batch's domain: l2 term only
                            double sq = 0.0;
                            for (int i = tid; i < a.n_domain * a.dd; i += kWorkers) {
                                if (i / a.dd == pd.dom) continue;
                                const float pe = ldcg_f(a.params + a.off_Ed + i);
                                sq += (double)pe * (double)pe;
                                apply1(a.off_Ed + i, __fmul_rn(2.0f * a.l2_emb, pe));
                            }
                            double* red = reinterpret_cast<double*>(scratch + 16384);
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
                            if (lane == 0) red[warp] = sq;
                            worker_sync();
                            if (tid == 0) {
                                double tot = 0.0;
                                for (int w = 0; w < kWorkerWarps; ++w) tot += red[w];
                                a.ed_sq[0] = tot;   // + the domain jobs' shares of the batch's own row
                            }
                            if (warp == 1) {
                                const double bs = warp_group_sum_d(a.loss_part + buf * kMaxGroups, ngroups, lane);
                                // |E_d|^2 of this mini-batch is published by domain job 0 in this same phase: the loss is
                                // completed after the barrier (below)
                                if (lane == 0) a.loss_part[2 * kMaxGroups + buf] = bs;
                            }
                            worker_sync();
                            WSTAMP(14);
                        }
                        continue;
                    }
                    // ---------- split-K tile job of dW_l: partial[z][tile] = H_l[rows of z]^T . dZ_l[rows of z]
                    if (warp == kProdWarp) {
                        const CUtensorMap* ma = l == 0 ? &maps.xmn[buf] : &maps.hmn[l];
                        const CUtensorMap* mb = &maps.dzmn[l];
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
#pragma unroll
                                    for (int g = 0; g < 4; ++g) tc::tma_load_3d(sA + z * A_BYTES + g * 4096, ma, fb, a_row + g * 32, kc, z);
                                    for (int g = 0; g < ngb; ++g) tc::tma_load_3d(sB + z * b_tile + g * 4096, mb, fb, b_col + g * 32, kc, z);
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
                        if (njob > 0) tc::mbar_wait(&bar_tfree, (njob - 1) & 1);
                        tc::tc_fence_after();
                        const uint32_t idesc1 = tc::make_idesc_tf32(128, J.bn, 1, 1);
                        const uint32_t idesc2 = tc::make_idesc_tf32(128, 2 * J.bn, 1, 1);
                        const uint32_t a_k = 1024u >> 4, b_k = 1024u >> 4;   // per k-step of 8
                        uint32_t acc = 0;
                        for (int i = 0; i < J.nch; ++i) {
                            tc::mbar_wait(&bar_full[ring_s], ring_ph);
                            tc::tc_fence_after();
                            const uint32_t st = (smem_base + (uint32_t)ring_s * STAGE_BYTES) >> 4;
                            if (tc::elect_one()) {
                                if (tim && i == 0) a.timing[tslot + 4] = (unsigned long long)clock64();
                                const uint32_t aw = st | tc::kDescLoMN, alw = aw + (A_BYTES >> 4);
                                const uint32_t bw = (st + ((2 * A_BYTES) >> 4)) | tc::kDescLoMN;
                                if (x3) {
#pragma unroll
                                    for (int k = 0; k < KCH / 8; ++k) {
                                        // A.[B | B_lo] -> columns [0, bn) and [bn, 2 bn);  A_lo.B -> columns [bn, 2 bn) too (small terms apart)
                                        tc::mma_tf32(tmem, tc::desc_words(aw + k * a_k, tc::kDescHiMN), tc::desc_words(bw + k * b_k, tc::kDescHiMN), idesc2, acc);
                                        tc::mma_tf32(tmem + (uint32_t)J.bn, tc::desc_words(alw + k * a_k, tc::kDescHiMN), tc::desc_words(bw + k * b_k, tc::kDescHiMN), idesc1, 1u);
                                        acc = 1;
                                    }
                                } else {
#pragma unroll
                                    for (int k = 0; k < KCH / 8; ++k) {
                                        tc::mma_tf32(tmem, tc::desc_words(aw + k * a_k, tc::kDescHiMN), tc::desc_words(bw + k * b_k, tc::kDescHiMN), idesc1, acc);
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
                        // ---------- workers: the split-K partial tile -> partials[l][z][tile][128][bn]
                        const int q = warp & 3, hf = warp >> 2;
                        const int rloc = q * 32 + lane;
                        tc::mbar_wait_warp(&bar_done, njob & 1);
                        tc::tc_fence_after();
                        if (tim && tid == 0) a.timing[tslot + 6] = (unsigned long long)clock64();
                        const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
                        const int tile = J.m_tile * J.NT + J.n_tile;
                        const int half = J.bn >> 1;   // columns per thread: 32 (bn = 64) or 16 (bn = 32)
                        float* ptile = a.partials[l] + ((long long)J.z * J.tiles + tile) * 128 * J.bn;
                        // last job of this CTA in the phase: the ring is idle -> transpose through shared memory for coalesced stores
                        const bool staged = j + G >= njobs;
                        float* s_p = reinterpret_cast<float*>(smem);   // [128][bn + 4]
                        for (int n0 = 0; n0 < half; n0 += 16) {
                            float vv[16], v2[16];
                            tc::tmem_ld16(tlane + hf * half + n0, vv);
                            if (x3) tc::tmem_ld16(tlane + J.bn + hf * half + n0, v2);
#pragma unroll
                            for (int jx = 0; jx < 16; jx += 4) {
                                float4 o4;
                                if (x3) o4 = make_float4(vv[jx] + v2[jx], vv[jx + 1] + v2[jx + 1], vv[jx + 2] + v2[jx + 2], vv[jx + 3] + v2[jx + 3]);
                                else o4 = make_float4(vv[jx], vv[jx + 1], vv[jx + 2], vv[jx + 3]);
                                if (staged) *reinterpret_cast<float4*>(s_p + rloc * (J.bn + 4) + hf * half + n0 + jx) = o4;
                                else __stcg(reinterpret_cast<float4*>(ptile + (long long)rloc * J.bn + hf * half + n0 + jx), o4);
                            }
                        }
                        tc::tc_fence_before();
                        tc::mbar_arrive(&bar_tfree);
                        if (staged) {
                            worker_sync();
                            const int bn4 = J.bn >> 2;
                            for (int idx = tid; idx < 128 * bn4; idx += kWorkers) {
                                const int r = idx / bn4, c = (idx - r * bn4) * 4;
                                __stcg(reinterpret_cast<float4*>(ptile + (long long)r * J.bn + c), *reinterpret_cast<const float4*>(s_p + r * (J.bn + 4) + c));
                            }
                        }
                        // publish: the release covers the partial-tile stores of every worker (ordered by the barrier)
                        worker_sync();
                        if (tid == 0) red_release_add_u32(a.tile_ctr + J.gtile, 1u);
                        WSTAMP(8);
                    }
                    ++njob;
                }
                // ---------- second sweep over this CTA's tile jobs (after ALL its partials are published: no wait cycle between
                // CTAs): when the S partials of the tile are in memory, reduce rows [z, z+1) * 128 / S of it in split order and
                // apply the optimizer to them
                {
                    int chunks, S, cps;
                    split_plan(rows, chunks, S, cps);
                    tile_target += (unsigned int)S;
                }
                if (warp < kWorkerWarps) {
                    for (int j = cta; j < njobs; j += G) {
                        const Job J = (rows == a.bs && j == cta) ? job_full : decode_dw_job(a, rows, j);
                        if (J.type != J_DW) continue;
                        if (tid == 0) {
                            while (ld_acquire_u32(a.tile_ctr + J.gtile) < tile_target) {}
                        }
                        worker_sync();
                        WSTAMP(9);
                        const int l = J.layer, N = a.n[l + 1], bn = J.bn, bn4 = bn >> 2;
                        const int tile = J.m_tile * J.NT + J.n_tile;
                        const int vrows = a.n[l] - J.m_tile * 128 < 128 ? a.n[l] - J.m_tile * 128 : 128;
                        const int rper = cdiv(128, J.S);
                        const int r_beg = J.z * rper, r_end = r_beg + rper < vrows ? r_beg + rper : vrows;
                        const long long zstride = (long long)J.tiles * 128 * bn;
                        const int nr = r_end - r_beg;
                        float* s_t = reinterpret_cast<float*>(smem);   // [nr][bn + 1] updated values (the ring is idle: every job of this CTA is done)
                        for (int idx = tid; idx < nr * bn4; idx += kWorkers) {
                            const int rr = idx / bn4, r = r_beg + rr, c = (idx - rr * bn4) * 4;
                            const float* src = a.partials[l] + ((long long)tile * 128 + r) * bn + c;
                            const long long o = a.off_W[l] + (long long)(J.m_tile * 128 + r) * N + J.n_tile * bn + c;
                            // every load of the item first: one round trip to L2
                            float4 q4[kMaxSplit];
#pragma unroll
                            for (int z = 0; z < kMaxSplit; ++z) q4[z] = z < J.S ? ldcg_f4(src + z * zstride) : make_float4(0.f, 0.f, 0.f, 0.f);
                            const float4 P = ldcg_f4(a.params + o);
                            float4 M = make_float4(0.f, 0.f, 0.f, 0.f), V = M;
                            if (a.opt_kind == 0) { M = ldcg_f4(a.m + o); V = ldcg_f4(a.v + o); }
                            float g[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                            for (int z = 0; z < kMaxSplit; ++z)
                                if (z < J.S) { g[0] += q4[z].x; g[1] += q4[z].y; g[2] += q4[z].z; g[3] += q4[z].w; }
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
                            if (l >= 1 || a.emb_trainable) store_pair4(a.wpair + o, a.wz, pnew, rnd, x3);
                            if (a.grads) *reinterpret_cast<float4*>(a.grads + o) = make_float4(g[0], g[1], g[2], g[3]);
#pragma unroll
                            for (int t = 0; t < 4; ++t) s_t[rr * (bn + 1) + c + t] = pp[t];
                        }
                        worker_sync();
                        {
                            // W_l^T[col, k]: consecutive lanes walk the k rows of the slice -> contiguous runs per column
                            const int K = a.n[l];
                            float* wt = a.wT[l] + (long long)(J.n_tile * bn) * K + J.m_tile * 128 + r_beg;
                            const long long tz = (long long)N * K;
                            for (int idx = tid; idx < nr * bn; idx += kWorkers) {
                                const int cc = idx / nr, rr = idx - cc * nr;
                                store_pair1(wt + (long long)cc * K + rr, tz, s_t[rr * (bn + 1) + cc], rnd, x3);
                            }
                        }
                        worker_sync();   // the next job of the sweep reuses the transpose buffer
                        WSTAMP(10);
                    }
                }
                if (a.opt_kind == 0) {
                    b1pow = __fmul_rn(b1pow, a.beta1);
                    b2pow = __fmul_rn(b2pow, a.beta2);
                }
                step_ctr += 1;
            }
            if (warp == kProdWarp && lane == 0) {
                // the next phase's first tensor maps: fetched into the descriptor cache while this CTA waits at the barrier
                if (phase == 0 && a.train) { tc::tma_prefetch_desc(&maps.xmn[buf]); tc::tma_prefetch_desc(&maps.dzmn[0]); }
                else { tc::tma_prefetch_desc(&maps.xk[buf ^ 1]); tc::tma_prefetch_desc(&maps.wf[0]); }
            }
            if (a.timing) {
                __syncthreads();
                if (tid == 0 && tim) a.timing[tslot + 1] = gtime();
            }
            grid_barrier(a.bar, bar_target);
        }
        // the Keras loss of this mini-batch (value only).  loss_part is double-buffered by step parity: the next
        // write to this buffer is two chain phases (>= one grid barrier that this thread also passes) away.
        if (cta == G - 1 && warp == 0) {
            double bs;
            if (a.train) bs = __ldcg(a.loss_part + 2 * kMaxGroups + buf);
            else bs = warp_group_sum_d(a.loss_part + buf * kMaxGroups, ngroups, lane);
            if (lane == 0) {
                double esq = __ldcg(a.ed_sq);
                if (a.train)
                    for (int q = 0; q < kDomJobs; ++q) esq += __ldcg(a.ed_sq + 1 + q);
                pd.losses[step] = (float)(bs / (double)rows + (double)a.frozen_reg + (double)a.l2_emb * esq);
            }
        }
    }

    // ---- epilogue of the pass: AUC accumulators (suffix sums of the pass histogram; the histogram is double-buffered
    // by pass parity, so the next pass of a program may already be filling the other one).  Done by the LAST CTA: it has no
    // row group (batch <= 16 x (G - 1) rows) and no dW job at the usual sizes, so the fold overlaps the next pass's first chain
    // phase instead of delaying CTA 0's chain
    if (cta == G - 1) {
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
    // a following sweep over the accumulators (AUC.reset_states recorded into the program) must come after the fold; any
    // other op is ordered by its own barriers (the histogram is double-buffered by pass parity)
    if (a.auc_acc && oi + 1 < a.n_ops && ops[oi + 1].kind == PROG_META && ops[oi + 1].w0 == a.auc_acc) grid_barrier(a.bar, bar_target);
    shadows_valid = a.train != 0;
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
    size_t dw_part, dg_part, loss_part, fold_part, ed_sq, wpair, wz, wT[MAMDR_MAX_LAYERS], dX, bid[2], uids[2], urows[2], ucnt, sws[2], total;
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
    const int ng = Bp / CR;                 // row groups of the chain
    w.bar = take(kBarBytes);                 // grid-barrier counter + the dW tile counters
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
        w.db_part[l] = take((size_t)ng * d.hidden[l] * 4);
    }
    w.dw_part = take((size_t)ng * d.hidden[L - 1] * 4);
    w.dg_part = take((size_t)ng * 4);
    w.loss_part = take((size_t)(2 * kMaxGroups + 2) * 8);
    w.fold_part = take((size_t)kDomJobs * d.hidden[0] * 4);
    w.ed_sq = take((size_t)(1 + kDomJobs) * 8);   // eval: [0] = |E_d|^2; training: [0] = rows other than the batch's, [1 + q] = domain job q's share
    w.wz = ((size_t)(d.arena_floats - d.off_domain_emb) + 31) / 32 * 32;   // floats between the two planes of the kernel shadow
    w.wpair = take(2 * w.wz * 4);                                          // dense span of the arena only
    for (int l = 0; l < L; ++l) w.wT[l] = take((size_t)2 * d.hidden[l] * (l == 0 ? K0 : d.hidden[l - 1]) * 4);
    w.dX = 0; w.ucnt = 0;
    for (int t = 0; t < 2; ++t) w.bid[t] = w.uids[t] = w.urows[t] = w.sws[t] = 0;
    if (d.emb_trainable) {   // what the sparse-gradient de-duplication behind a launch consumes / produces
        w.dX = take((size_t)Bp * K0 * 4);
        w.ucnt = take(64);
        for (int t = 0; t < 2; ++t) {
            w.bid[t] = take((size_t)Bp * 4);
            w.uids[t] = take((size_t)Bp * 4);
            w.urows[t] = take((size_t)Bp * d.emb_dim[t] * 4);
            w.sws[t] = take(mamdr_scatter_workspace_bytes(Bp));
        }
    }
    w.total = off;
    return w;
}

inline size_t smem_bytes() { return (size_t)kStages * STAGE_BYTES + kScratchBytes + 1024; }

}  // namespace passk

using namespace passk;

static int pass_supported(mamdr_ctx* ctx, const mamdr_mlp_desc* d, int max_batch) {
    MAMDR_REQUIRE(ctx, d->n_layers >= 1 && d->n_layers <= MAMDR_MAX_LAYERS, MAMDR_E_INVALID, "n_layers out of range");
    if (d->emb_trainable)
        MAMDR_REQUIRE(ctx, max_batch <= mamdr_scatter_max_n() && d->off_user_emb >= 0 && d->off_item_emb >= 0, MAMDR_E_UNSUPPORTED,
                      "trainable tables: the batch must fit the sparse-gradient de-duplication and the tables must live in the arena");
    const int K0 = d->emb_dim[0] + d->emb_dim[1];
    MAMDR_REQUIRE(ctx, d->emb_dim[0] % 4 == 0 && d->emb_dim[1] % 4 == 0 && d->emb_dim[2] % 4 == 0 && K0 % 32 == 0, MAMDR_E_UNSUPPORTED,
                  "pass kernel needs emb dims that are multiples of 4 and user+item width a multiple of 32");
    for (int l = 0; l < d->n_layers; ++l)
        MAMDR_REQUIRE(ctx, d->hidden[l] % 32 == 0 && (d->hidden[l] == 32 || d->hidden[l] % 64 == 0) && d->hidden[l] <= kMaxWidth, MAMDR_E_UNSUPPORTED,
                      "pass kernel needs hidden widths of 32 or multiples of 64, at most 256 (the row-local chain keeps a layer in shared memory)");
    MAMDR_REQUIRE(ctx, K0 <= kMaxWidth && d->emb_dim[2] <= 256, MAMDR_E_UNSUPPORTED, "pass kernel needs user+item width <= 256 and domain width <= 256");
    const int nl = d->hidden[d->n_layers - 1];
    MAMDR_REQUIRE(ctx, nl == 32 || nl == 64, MAMDR_E_UNSUPPORTED, "pass kernel needs a last hidden width of 32 or 64");
    MAMDR_REQUIRE(ctx, max_batch >= 1 && max_batch <= CR * kMaxGroups, MAMDR_E_UNSUPPORTED, "batch too large for the pass kernel");
    MAMDR_REQUIRE(ctx, d->dropout_rate >= 0.f && d->dropout_rate < 1.f, MAMDR_E_INVALID, "dropout_rate must be in [0,1)");
    return MAMDR_OK;
}

int mamdr_pass_init_kernels(mamdr_ctx* ctx) {
    MAMDR_CUDA_OK(ctx, cudaFuncSetAttribute(pass_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes()));
    MAMDR_CUDA_OK(ctx, cudaFuncSetAttribute(pass_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes()));
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
    MAMDR_CUDA_OK(ctx, cudaMemsetAsync(a.bar, 0, kBarBytes, st));
    const size_t smem = smem_bytes();
    void* kargs[] = {(void*)&mp, (void*)&a};
    const void* fn = a.passes == 3 ? (const void*)pass_kernel<3> : (const void*)pass_kernel<1>;
    MAMDR_CUDA_OK(ctx, cudaLaunchCooperativeKernel(fn, dim3(ctx->pass_ctas > 0 ? ctx->pass_ctas : ctx->sm_count), dim3(kThreads), kargs, smem, st));
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
    MAMDR_REQUIRE(ctx, ps->uid_dev && ps->pid_dev && ps->label_dev && (d->emb_trainable || (ut && it)), MAMDR_E_INVALID, "NULL data column / table");
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
    a.emb_trainable = d->emb_trainable ? 1 : 0;
    if (d->emb_trainable) {
        MAMDR_REQUIRE(ctx, !train || (ps->steps == 1 && !ctx->prog), MAMDR_E_UNSUPPORTED,
                      "trainable tables: one mini-batch per launch (the table sweeps run between mini-batches), not recordable");
        a.Eu = params + d->off_user_emb;
        a.Ei = params + d->off_item_emb;
    }
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
    a.fold_part = (float*)(ws + w.fold_part);
    if (d->emb_trainable) {
        a.dX = (float*)(ws + w.dX);
        a.bid[0] = (int32_t*)(ws + w.bid[0]); a.bid[1] = (int32_t*)(ws + w.bid[1]);
    }
    a.ed_sq = (double*)(ws + w.ed_sq);
    a.wpair = (float*)(ws + w.wpair) - d->off_domain_emb;   // indexed with arena offsets
    a.wz = (long long)w.wz;
    for (int l = 0; l < L; ++l) a.wT[l] = (float*)(ws + w.wT[l]);
    a.hist = (int*)(ws + w.hist);
    a.bar = (unsigned int*)(ws + w.bar);
    a.tile_ctr = a.bar + 16;
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
    const float* wsrc = a.wpair;   // M operands of the dH GEMMs: the pair shadow of the kernels
    const uint32_t nzp = a.passes == 3 ? 2 : 1;   // planes per TMA box
    for (int b = 0; b < 2; ++b) {
        const uint64_t z = (uint64_t)Bp * a.n[0];
        ok = ok && mlptc::pair_kmajor_map(ctx, &mp.xk[b], a.X[b], Bp, a.n[0], CR, z, nzp) && mlptc::pair_mnmajor_map(ctx, &mp.xmn[b], a.X[b], Bp, a.n[0], z);
    }
    for (int l = 0; l < L; ++l) {
        const uint64_t zh = (uint64_t)Bp * a.n[l], zd = (uint64_t)Bp * a.n[l + 1];
        if (l >= 1) {
            ok = ok && mlptc::pair_mnmajor_map(ctx, &mp.hmn[l], a.H[l], Bp, a.n[l], zh);
            ok = ok && mlptc::pair_kmajor_map(ctx, &mp.wb[l], wsrc + d->off_kernel[l], a.n[l], a.n[l + 1], a.n[l] < 128 ? a.n[l] : 128, (uint64_t)a.wz, nzp);
        }
        if (l == 0 && d->emb_trainable)
            ok = ok && mlptc::pair_kmajor_map(ctx, &mp.wb[0], wsrc + d->off_kernel[0], a.n[0], a.n[1], a.n[0] < 128 ? a.n[0] : 128, (uint64_t)a.wz, nzp);
        ok = ok && mlptc::pair_mnmajor_map(ctx, &mp.dzmn[l], a.dZ[l], Bp, a.n[l + 1], zd);
        ok = ok && mlptc::pair_kmajor_map(ctx, &mp.wf[l], a.wT[l], a.n[l + 1], a.n[l], a.n[l + 1] < 128 ? a.n[l + 1] : 128, (uint64_t)a.n[l + 1] * a.n[l], nzp);
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
    rc = launch_program(ctx, mp, a, st);
    if (rc || !d->emb_trainable || !train) return rc;
    // de-duplicate the two sparse gradients of this mini-batch (K6): sorted unique ids + rows summed in batch order
    const int rows = (int)(ps->n_data < ps->batch_size ? ps->n_data : ps->batch_size);
    DedupJob jobs[2];
    for (int t = 0; t < 2; ++t) {
        jobs[t] = DedupJob{a.bid[t], a.dX + (t == 0 ? 0 : d->emb_dim[0]), (int64_t)a.n[0], d->emb_dim[t], (int32_t*)(ws + w.uids[t]),
                           (float*)(ws + w.urows[t]), (int32_t*)(ws + w.ucnt) + t, nullptr, nullptr};
        mamdr_scatter_job_ws(&jobs[t], ws + w.sws[t], Bp);
    }
    return mamdr_scatter_dedup_jobs(ctx, jobs, 2, rows, st);
}

/* the de-duplicated sparse gradients left by the last trainable-table launch of the pass kernel (see mamdr_mlp_sparse_grads) */
extern "C" int mamdr_mlp_pass_sparse_grads(const mamdr_mlp_desc* desc, int32_t batch_size, void* ws_dev, int32_t table,
                                           const int32_t** uniq_ids_dev, const float** uniq_rows_dev, const int32_t** n_uniq_dev) {
    if (!desc || !ws_dev || table < 0 || table > 1 || !desc->emb_trainable) return MAMDR_E_INVALID;
    const PassWs w = pass_ws(*desc, batch_size);
    unsigned char* ws = (unsigned char*)ws_dev;
    *uniq_ids_dev = (const int32_t*)(ws + w.uids[table]);
    *uniq_rows_dev = (const float*)(ws + w.urows[table]);
    *n_uniq_dev = (const int32_t*)(ws + w.ucnt) + table;
    return MAMDR_OK;
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
