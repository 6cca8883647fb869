// Probe (not part of the library): issue cost of tcgen05.mma kind::tf32 as a function of the instruction shape.
// One CTA, operands resident in shared memory, R back-to-back MMAs on one accumulator, commit, wait.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mma_rate mma_rate.cu
#include <cstdio>
#include "../tc_common.cuh"

__global__ void __launch_bounds__(128) probe(int M, int N, int a_mn, int b_mn, int reps, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t done;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.f;
    if (tid == 0) { tc::mbar_init(&done, 1); tc::fence_barrier_init(); }
    if (warp == 1) tc::tmem_alloc(&tmem_base, 256);
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    if (tid == 32) {
        const uint32_t idesc = tc::make_idesc_tf32(M, N, a_mn, b_mn);
        const uint32_t aA = tc::smem_u32(smem), aB = aA + 16384;
        const uint64_t dA = a_mn ? tc::make_smem_desc(aA, 4096, 512, 1) : tc::make_smem_desc(aA, 16, 1024, tc::kSwizzle128B);
        const uint64_t dB = b_mn ? tc::make_smem_desc(aB, 4096, 512, 1) : tc::make_smem_desc(aB, 16, 1024, tc::kSwizzle128B);
        const uint64_t sA = a_mn ? 64 : 2, sB = b_mn ? 64 : 2;
        for (int trial = 0; trial < 2; ++trial) {
            const long long t0 = clock64();
            for (int r = 0; r < reps; ++r) tc::mma_tf32(tmem_base, dA + (r & 3) * sA, dB + (r & 3) * sB, idesc, r > 0);
            const long long t1 = clock64();
            tc::mma_commit(&done);
            tc::mbar_wait(&done, trial & 1);
            const long long t2 = clock64();
            out[0] = t1 - t0;
            out[1] = t2 - t0;
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_base, 256);
}

int main() {
    long long* d;
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 32768 + 1024);
    const int reps = 64;
    for (int M : {128, 64})
        for (int mn = 0; mn < 2; ++mn)
            for (int N : {16, 32, 64, 128, 256}) {
                if (mn && N > 64) continue;   // MN-major B tile in this probe: 32-column boxes, keep it small
                probe<<<1, 128, 16384 + 32768 + 1024>>>(M, N, mn, mn, reps, d);
                cudaError_t e = cudaDeviceSynchronize();
                long long h[2];
                cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                printf("M=%3d N=%3d %s: issue %.0f cyc/MMA, complete %.0f cyc/MMA (%d MMAs)  %s\n", M, N, mn ? "MN-major A,B" : "K-major  A,B",
                       h[0] / (double)reps, h[1] / (double)reps, reps, e == cudaSuccess ? "" : cudaGetErrorString(e));
            }
    return 0;
}
