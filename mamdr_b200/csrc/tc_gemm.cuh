// tcgen05 tile GEMM for the MLP tower (sm_100a): TMA-staged operands, tcgen05.mma kind::tf32 with fp32
// accumulators in TMEM, 1-pass TF32 or 3-pass error-compensated 3xTF32, fused epilogue functor.
//
//   C[m, n] = sum_k A(m, k) * B(n, k)          tile = 128 (M) x BN (N), K streamed in 32-float chunks
//   A_MN = false : A stored [M, K] (k contiguous)  -> K-major operand,  TMA SWIZZLE_128B,          UMMA layout 2
//   A_MN = true  : A stored [K, M] (m contiguous)  -> MN-major operand, TMA SWIZZLE_128B_ATOM_32B, UMMA layout 1
//   (same for B with N in place of M).  Encodings validated on a B200 by csrc/probe/tc_probe.cu.
//
// 3xTF32: the tensor core TRUNCATES fp32 inputs to tf32 (measured).  With x_lo = rn_tf32(x - trunc_tf32(x))
// precomputed by the producer of x,   A.B ~= A*B + A*B_lo + A_lo*B   (raw operands are truncated by the HW to
// their hi parts), which restores ~2^-21 relative accuracy per product with fp32 accumulation.
//
// 128 threads: warp 0 lane 0 = TMA producer, warp 1 lane 0 = MMA issuer, then all 4 warps run the epilogue
// (thread t <-> TMEM lane t <-> tile row t).  One output tile per CTA; optional split-K over gridDim.z with a
// deterministic last-CTA fix-up (partials re-read in z order; no float atomics).
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace tcg {

constexpr int KCH = 32;          // floats per K chunk (= one 128-byte swizzle row)
constexpr int A_BYTES = 128 * KCH * 4;

__device__ __forceinline__ float tf32_lo(float x) {
    const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    uint32_t u = __float_as_uint(x - hi);            // exact
    u += 0x00000FFFu + ((u >> 13) & 1u);             // round to nearest even at 13 dropped bits
    return __uint_as_float(u & 0xFFFFE000u);
}

struct Maps {
    CUtensorMap a, a_lo, b, b_lo;
};

struct SplitK {
    float*        partials;  // [z][tile][128][BN]
    unsigned int* tickets;   // [tile]
};

#ifdef TC_TIMING
__device__ unsigned long long* g_tc_timing = nullptr;   // [cta][8] globaltimer stamps (probe builds only)
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define TC_STAMP(i)                                                                                              \
    do {                                                                                                         \
        if (g_tc_timing && threadIdx.x == 0)                                                                     \
            g_tc_timing[((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8 + (i)] = gtime(); \
    } while (0)
#else
#define TC_STAMP(i)
#endif

template <int BN, bool A_MN, bool B_MN, int STAGES, class Epi>
__global__ void __launch_bounds__(128)
gemm_kernel(const __grid_constant__ Maps maps, int M_total, int chunks_total, int chunks_per_split, int passes, SplitK sk, Epi epi) {
    constexpr int B_BYTES = BN * KCH * 4;
    constexpr int STAGE_BYTES = 2 * (A_BYTES + B_BYTES);
    constexpr uint32_t TMEM_COLS = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar_full[STAGES], bar_empty[STAGES], bar_done;
    __shared__ uint32_t tmem_base_s;
    __shared__ bool is_last;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_tile = blockIdx.x, m_tile = blockIdx.y;
    const int c_beg = blockIdx.z * chunks_per_split;
    const int c_end = min(chunks_total, c_beg + chunks_per_split);
    const int nch = c_end - c_beg;
    TC_STAMP(0);

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            tc::mbar_init(&bar_full[s], 1);
            tc::mbar_init(&bar_empty[s], 1);
        }
        tc::mbar_init(&bar_done, 1);
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc(&tmem_base_s, TMEM_COLS);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    TC_STAMP(1);

    if (warp == 0 && lane == 0) {
        // ---------------- TMA producer
        const uint32_t tx = (uint32_t)(A_BYTES + B_BYTES) * (passes == 3 ? 2u : 1u);
        for (int i = 0; i < nch; ++i) {
            const int s = i % STAGES, c = c_beg + i;
            if (i >= STAGES) tc::mbar_wait(&bar_empty[s], ((i / STAGES) - 1) & 1);
            unsigned char* st = smem + (size_t)s * STAGE_BYTES;
            unsigned char *sA = st, *sAlo = st + A_BYTES, *sB = st + 2 * A_BYTES, *sBlo = st + 2 * A_BYTES + B_BYTES;
            tc::mbar_arrive_expect_tx(&bar_full[s], tx);
            if (A_MN) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    tc::tma_load_2d(sA + g * 4096, &maps.a, &bar_full[s], m_tile * 128 + g * 32, c * KCH);
                    if (passes == 3) tc::tma_load_2d(sAlo + g * 4096, &maps.a_lo, &bar_full[s], m_tile * 128 + g * 32, c * KCH);
                }
            } else {
                tc::tma_load_2d(sA, &maps.a, &bar_full[s], c * KCH, m_tile * 128);
                if (passes == 3) tc::tma_load_2d(sAlo, &maps.a_lo, &bar_full[s], c * KCH, m_tile * 128);
            }
            if (B_MN) {
#pragma unroll
                for (int g = 0; g < BN / 32; ++g) {
                    tc::tma_load_2d(sB + g * 4096, &maps.b, &bar_full[s], n_tile * BN + g * 32, c * KCH);
                    if (passes == 3) tc::tma_load_2d(sBlo + g * 4096, &maps.b_lo, &bar_full[s], n_tile * BN + g * 32, c * KCH);
                }
            } else {
                tc::tma_load_2d(sB, &maps.b, &bar_full[s], c * KCH, n_tile * BN);
                if (passes == 3) tc::tma_load_2d(sBlo, &maps.b_lo, &bar_full[s], c * KCH, n_tile * BN);
            }
        }
        TC_STAMP(2);
    } else if (warp == 1 && lane == 0) {
        // ---------------- MMA issuer
        constexpr uint32_t idesc = tc::make_idesc_tf32(128, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
        uint32_t acc = 0;
        for (int i = 0; i < nch; ++i) {
            const int s = i % STAGES;
            tc::mbar_wait(&bar_full[s], (i / STAGES) & 1);
            tc::tc_fence_after();
            const uint32_t st = tc::smem_u32(smem + (size_t)s * STAGE_BYTES);
            const uint32_t aA = st, aAlo = st + A_BYTES, aB = st + 2 * A_BYTES, aBlo = st + 2 * A_BYTES + B_BYTES;
            for (int pass = 0; pass < passes; ++pass) {
                const uint32_t pa = (pass == 2) ? aAlo : aA;   // pass 0: A*B, 1: A*B_lo, 2: A_lo*B
                const uint32_t pb = (pass == 1) ? aBlo : aB;
#pragma unroll
                for (int k = 0; k < KCH / 8; ++k) {
                    const uint64_t da = A_MN ? tc::make_smem_desc(pa + k * 1024, 4096, 512, 1)
                                             : tc::make_smem_desc(pa + k * 32, 16, 1024, tc::kSwizzle128B);
                    const uint64_t db = B_MN ? tc::make_smem_desc(pb + k * 1024, 4096, 512, 1)
                                             : tc::make_smem_desc(pb + k * 32, 16, 1024, tc::kSwizzle128B);
                    tc::mma_tf32(tmem, da, db, idesc, acc);
                    acc = 1;
                }
            }
            tc::mma_commit(&bar_empty[s]);   // frees the stage when these MMAs have read it
        }
        tc::mma_commit(&bar_done);
    }
    __syncthreads();
    TC_STAMP(3);
    tc::mbar_wait(&bar_done, 0);
    tc::tc_fence_after();
    TC_STAMP(4);

    // ---------------- epilogue: thread t owns tile row t
    // Epi protocol: begin(st,row,valid) ; cols(st,row,valid,col0,local_col0,v[16]) per 16 columns ; end(st,row,valid,m_tile,n_tile,scratch)
    // end() may use __syncthreads and the (now idle) pipeline smem as scratch, so every thread calls it.
    const int row = m_tile * 128 + tid;
    const bool valid = row < M_total;
    const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);
    typename Epi::State est;
    if (gridDim.z > 1) {
        const int tile = m_tile * gridDim.x + n_tile;
        const size_t tile_elems = (size_t)128 * BN;
        float* mine = sk.partials + ((size_t)blockIdx.z * gridDim.x * gridDim.y + tile) * tile_elems + (size_t)tid * BN;
#pragma unroll 1
        for (int n0 = 0; n0 < BN; n0 += 16) {
            float v[16];
            tc::tmem_ld16(tlane + n0, v);
#pragma unroll
            for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(mine + n0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) is_last = (atomicAdd(&sk.tickets[tile], 1u) == gridDim.z - 1);
        __syncthreads();
        if (is_last) {
            __threadfence();
            epi.begin(est, row, valid);
#pragma unroll
            for (int n0 = 0; n0 < BN; n0 += 16) {
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = 0.f;
                for (unsigned int z = 0; z < gridDim.z; ++z) {
                    const float* src = sk.partials + ((size_t)z * gridDim.x * gridDim.y + tile) * tile_elems + (size_t)tid * BN + n0;
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const float4 q = __ldcg(reinterpret_cast<const float4*>(src + j));
                        v[j] += q.x; v[j + 1] += q.y; v[j + 2] += q.z; v[j + 3] += q.w;
                    }
                }
                epi.cols(est, row, valid, n_tile * BN + n0, n0, v);
            }
            epi.end(est, row, valid, m_tile, n_tile, smem);
            if (tid == 0) sk.tickets[tile] = 0;
        }
    } else {
        epi.begin(est, row, valid);
#pragma unroll
        for (int n0 = 0; n0 < BN; n0 += 16) {
            float v[16];
            tc::tmem_ld16(tlane + n0, v);
            epi.cols(est, row, valid, n_tile * BN + n0, n0, v);
        }
        epi.end(est, row, valid, m_tile, n_tile, smem);
    }
    TC_STAMP(5);
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem, TMEM_COLS);
    TC_STAMP(6);
}

template <int BN, int STAGES>
constexpr size_t smem_bytes() { return (size_t)STAGES * 2 * (A_BYTES + BN * KCH * 4) + 1024; }

}  // namespace tcg
