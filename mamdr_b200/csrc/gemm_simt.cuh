// fp32 FFMA (SIMT) tile GEMM used by MAMDR_PREC_FP32 -- the parity mode of the tower
// (tcgen05 has no fp32-input MMA, so exact-fp32 runs use FFMA).
//
//   C[M,N] = sum_k A(m,k) * B(k,n)  followed by a fused epilogue functor.
//   A_KCONTIG: A is [M,K] row-major (k contiguous)   else A is [K,M] (m contiguous)
//   B_NCONTIG: B is [K,N] row-major (n contiguous)   else B is [N,K] (k contiguous)
//   Contiguous extents must be multiples of 4 (128-bit loads); the other extents are arbitrary.
//
// Tile 32x64x16, 128 threads, 4x4 register micro-tile.  Optional split-K across gridDim.z with a
// deterministic fix-up: every CTA parks its partial tile in the workspace; the last CTA to arrive
// (ticket) re-reads ALL partials in z order and runs the epilogue, so the result does not depend
// on CTA scheduling (no float atomics anywhere).
#pragma once
#include "common.cuh"

namespace simt {

constexpr int BM = 32, BN = 64, BK = 16, THREADS = 128, APAD = 4;

struct GemmShape {
    int M, N, K;
    int lda, ldb;  // leading dimensions (floats) of A and B in their stored orientation
};

// acc += A[:, kbeg:kend] . B[kbeg:kend, :] for this CTA's 32 x 64 tile (all threads of the CTA must call it)
template <bool A_KCONTIG, bool B_NCONTIG>
__device__ __forceinline__ void
mainloop(const float* __restrict__ A, const float* __restrict__ B, const GemmShape& s, const int kbeg, const int kend,
         float (&acc)[4][4]) {
    __shared__ __align__(16) float As[BK][BM + APAD];
    __shared__ __align__(16) float Bs[BK][BN];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    for (int k0 = kbeg; k0 < kend; k0 += BK) {
        // ---- stage A tile (BM x BK) into As[k][m]
        if (A_KCONTIG) {
            const int m = tid >> 2, k4 = (tid & 3) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            const int gm = m0 + m, gk = k0 + k4;
            if (gm < s.M && gk < kend) {
                if (gk + 3 < kend) {
                    v = ldg_f4(A + (int64_t)gm * s.lda + gk);
                } else {  // k_chunk boundaries are multiples of 4 unless K itself is ragged
                    const float* p = A + (int64_t)gm * s.lda + gk;
                    v.x = __ldg(p);
                    if (gk + 1 < kend) v.y = __ldg(p + 1);
                    if (gk + 2 < kend) v.z = __ldg(p + 2);
                }
            }
            As[k4 + 0][m] = v.x; As[k4 + 1][m] = v.y; As[k4 + 2][m] = v.z; As[k4 + 3][m] = v.w;
        } else {
            const int k = tid >> 3, m4 = (tid & 7) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            const int gk = k0 + k, gm = m0 + m4;
            if (gk < kend && gm < s.M) v = ldg_f4(A + (int64_t)gk * s.lda + gm);
            *reinterpret_cast<float4*>(&As[k][m4]) = v;
        }
        // ---- stage B tile (BK x BN) into Bs[k][n]
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int idx = tid + i * THREADS;
            if (B_NCONTIG) {
                const int k = idx >> 4, n4 = (idx & 15) * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                const int gk = k0 + k, gn = n0 + n4;
                if (gk < kend && gn < s.N) v = ldg_f4(B + (int64_t)gk * s.ldb + gn);
                *reinterpret_cast<float4*>(&Bs[k][n4]) = v;
            } else {
                const int n = idx >> 2, k4 = (idx & 3) * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                const int gn = n0 + n, gk = k0 + k4;
                if (gn < s.N && gk < kend) {
                    if (gk + 3 < kend) {
                        v = ldg_f4(B + (int64_t)gn * s.ldb + gk);
                    } else {
                        const float* p = B + (int64_t)gn * s.ldb + gk;
                        v.x = __ldg(p);
                        if (gk + 1 < kend) v.y = __ldg(p + 1);
                        if (gk + 2 < kend) v.z = __ldg(p + 2);
                    }
                }
                Bs[k4 + 0][n] = v.x; Bs[k4 + 1][n] = v.y; Bs[k4 + 2][n] = v.z; Bs[k4 + 3][n] = v.w;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
}

template <class Epilogue>
__device__ __forceinline__ void finish_tile(float (&acc)[4][4], int M, int N, float* __restrict__ partials,
                                            unsigned int* __restrict__ tickets, const Epilogue& epi, unsigned int zi, unsigned int zn);

// One CTA's tile; (zi, zn) = this CTA's split-K slice and the number of slices.
template <bool A_KCONTIG, bool B_NCONTIG, class Epilogue>
__device__ __forceinline__ void
gemm_tile(const float* __restrict__ A, const float* __restrict__ B, const GemmShape& s, int k_chunk,
          float* __restrict__ partials, unsigned int* __restrict__ tickets, const Epilogue& epi, const unsigned int zi,
          const unsigned int zn) {
    const int kbeg = (int)zi * k_chunk;
    const int kend = min(s.K, kbeg + k_chunk);

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    mainloop<A_KCONTIG, B_NCONTIG>(A, B, s, kbeg, kend, acc);
    finish_tile(acc, s.M, s.N, partials, tickets, epi, zi, zn);
}

// Split-K fix-up + epilogue of one CTA's tile: with zn > 1 every CTA parks its partial tile; the last one to arrive
// (ticket) re-reads ALL partials in z order and runs the epilogue.
template <class Epilogue>
__device__ __forceinline__ void
finish_tile(float (&acc)[4][4], const int M, const int N, float* __restrict__ partials, unsigned int* __restrict__ tickets,
            const Epilogue& epi, const unsigned int zi, const unsigned int zn) {
    __shared__ bool is_last;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int gn = n0 + tx * 4;
    if (zn > 1) {
        // park the partial tile: layout [z][tile][BM][BN]
        const int tile = blockIdx.y * gridDim.x + blockIdx.x;
        const int64_t tile_elems = (int64_t)BM * BN;
        float* mine = partials + ((int64_t)zi * gridDim.x * gridDim.y + tile) * tile_elems;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            *reinterpret_cast<float4*>(mine + (ty * 4 + i) * BN + tx * 4) =
                make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            const unsigned int t = atomicAdd(&tickets[tile], 1u);
            is_last = (t == zn - 1);
        }
        __syncthreads();
        if (!is_last) return;
        __threadfence();
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        for (unsigned int z = 0; z < zn; ++z) {
            const float* src = partials + ((int64_t)z * gridDim.x * gridDim.y + tile) * tile_elems;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 v = __ldcg(reinterpret_cast<const float4*>(src + (ty * 4 + i) * BN + tx * 4));
                acc[i][0] += v.x; acc[i][1] += v.y; acc[i][2] += v.z; acc[i][3] += v.w;
            }
        }
        if (tid == 0) tickets[tile] = 0;  // re-arm for the next launch
    }
    if (gn < N) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int gm = m0 + ty * 4 + i;
            if (gm < M) epi(gm, gn, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
        }
    }
}

template <bool A_KCONTIG, bool B_NCONTIG, class Epilogue>
__global__ void __launch_bounds__(THREADS)
gemm_kernel(const float* __restrict__ A, const float* __restrict__ B, GemmShape s, int k_chunk,
            float* __restrict__ partials, unsigned int* __restrict__ tickets, Epilogue epi) {
    gemm_tile<A_KCONTIG, B_NCONTIG, Epilogue>(A, B, s, k_chunk, partials, tickets, epi, blockIdx.z, gridDim.z);
}

// Grouped variant: up to MAX_GROUPS independent GEMMs of ONE shape in a single launch (the experts of a multi-task
// tower).  gridDim.z = n_groups * split; group g owns ticket / partial slices of its own.
constexpr int MAX_GROUPS = 8;
template <class Epilogue>
struct GroupedArgs {
    const float* A[MAX_GROUPS];
    const float* B[MAX_GROUPS];
    Epilogue     epi[MAX_GROUPS];
    int          split;           // split-K slices per group
    int64_t      partial_stride;  // floats between the partial regions of consecutive groups
    int          ticket_stride;   // tiles per group
};

template <bool A_KCONTIG, bool B_NCONTIG, class Epilogue>
__global__ void __launch_bounds__(THREADS)
gemm_grouped_kernel(const __grid_constant__ GroupedArgs<Epilogue> a, GemmShape s, int k_chunk, float* __restrict__ partials,
                    unsigned int* __restrict__ tickets) {
    const unsigned int g = blockIdx.z / a.split, z = blockIdx.z % a.split;
    gemm_tile<A_KCONTIG, B_NCONTIG, Epilogue>(a.A[g], a.B[g], s, k_chunk, partials + g * a.partial_stride,
                                              tickets + g * a.ticket_stride, a.epi[g], z, (unsigned int)a.split);
}

// K-segmented variant: C = sum_g A_g . B_g over up to MAX_GROUPS operand pairs that share M and N but not K (the input
// gradient of a layer fed to several experts: dX = sum_j dZ_j . W_j^T).  One CTA per (tile, segment); the deterministic
// split-K fix-up adds the segments in order.
struct SegmentArgs {
    const float* A[MAX_GROUPS + 1];
    const float* B[MAX_GROUPS + 1];
    int          K[MAX_GROUPS + 1];   // also lda / ldb of the segment (k-contiguous operands)
    int          n;
};

template <class Epilogue>
__global__ void __launch_bounds__(THREADS)
gemm_ksegments_kernel(const __grid_constant__ SegmentArgs a, int M, int N, float* __restrict__ partials,
                      unsigned int* __restrict__ tickets, Epilogue epi) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int g = blockIdx.z;   // gridDim.z == a.n: one segment per slice, summed in segment order by the fix-up
    const GemmShape s{M, N, a.K[g], a.K[g], a.K[g]};
    mainloop<true, false>(a.A[g], a.B[g], s, 0, a.K[g], acc);
    finish_tile(acc, M, N, partials, tickets, epi, blockIdx.z, gridDim.z);
}

struct LaunchPlan {
    dim3 grid;
    int  k_chunk;
};

// split K so that the launch has roughly `target_ctas` CTAs; k_chunk is a multiple of BK
inline LaunchPlan plan(int M, int N, int K, int target_ctas, int max_split) {
    const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    int split = 1;
    if (tiles < target_ctas) split = target_ctas / tiles;
    if (split > max_split) split = max_split;
    const int kblocks = (K + BK - 1) / BK;
    if (split > kblocks) split = kblocks;
    if (split < 1) split = 1;
    int chunk_blocks = (kblocks + split - 1) / split;
    split = (kblocks + chunk_blocks - 1) / chunk_blocks;
    LaunchPlan p;
    p.grid = dim3((N + BN - 1) / BN, (M + BM - 1) / BM, split);
    p.k_chunk = chunk_blocks * BK;
    return p;
}

}  // namespace simt
