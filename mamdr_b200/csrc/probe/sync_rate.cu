// Probe (not part of the library): cost of the grid-wide synchronisation primitives the pass kernel can use.
//   mode 0: r1 grid barrier  -- one monotonic counter, red.release + ld.acquire spin by thread 0 of every CTA
//   mode 1: distributed      -- arrivals spread over NC counters in separate 128-byte lines, pollers sum all of them
//   mode 2: flag hop         -- a ring: CTA i waits for CTA i-1's flag, then raises its own (latency of ONE release ->
//                               acquire hand-off between two SMs, what a per-row-tile dependency counter costs)
//   mode 3: group counters   -- 8 CTAs arrive on one counter, 8 other CTAs wait for it (the fwd -> fwd hand-off)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o sync_rate sync_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned int ld_relaxed_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add_u32(unsigned int* p, unsigned int v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(320, 1) probe(unsigned int* ctr, int mode, int NC, int iters, long long* out) {
    const int G = gridDim.x, cta = blockIdx.x, tid = threadIdx.x;
    unsigned int target = 0;
    __syncthreads();
    const long long t0 = clock64();
    if (mode == 0) {
        for (int i = 0; i < iters; ++i) {
            __syncthreads();
            target += G;
            if (tid == 0) {
                red_release_add_u32(ctr, 1u);
                while (ld_acquire_u32(ctr) < target) {}
            }
            __syncthreads();
        }
    } else if (mode == 1) {
        for (int i = 0; i < iters; ++i) {
            __syncthreads();
            target += G;
            if (tid == 0) red_release_add_u32(ctr + 32 * (cta % NC), 1u);
            if (tid < 32) {
                // lane q polls counter q; the sum over lanes is the number of arrivals
                unsigned int tot;
                do {
                    unsigned int v = tid < NC ? ld_relaxed_u32(ctr + 32 * tid) : 0u;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    tot = v;
                } while (tot < target);
                __threadfence();   // acquire side
            }
            __syncthreads();
        }
    } else if (mode == 2) {
        // ring hand-off: iteration i, CTA c waits for flag[(c-1+G)%G] >= i*? -- a token travels round the ring
        if (tid == 0) {
            for (int i = 0; i < iters; ++i) {
                const unsigned int want = (unsigned int)(i * G + cta);   // token count when it is my turn
                while (ld_acquire_u32(ctr) < want) {}
                red_release_add_u32(ctr, 1u);
            }
        }
        __syncthreads();
    } else {
        // groups of 16 CTAs: the first 8 produce (arrive on the group's counter), the last 8 consume (wait), then the
        // roles swap on a second counter -- one iteration = two dependent 8 -> 8 hand-offs
        const int grp = cta / 16, r = cta % 16;
        unsigned int* c0 = ctr + 64 * grp;
        unsigned int* c1 = ctr + 64 * grp + 32;
        if (cta < (G / 16) * 16) {
            for (int i = 1; i <= iters; ++i) {
                __syncthreads();
                if (tid == 0) {
                    if (r < 8) { red_release_add_u32(c0, 1u); while (ld_acquire_u32(c1) < 8u * i) {} }
                    else { while (ld_acquire_u32(c0) < 8u * i) {} red_release_add_u32(c1, 1u); }
                }
                __syncthreads();
            }
        }
    }
    const long long t1 = clock64();
    if (tid == 0) out[cta] = t1 - t0;
}

int main() {
    unsigned int* ctr;
    cudaMalloc(&ctr, 64 * 1024);
    long long* d;
    cudaMalloc(&d, 148 * 8);
    const int iters = 2000;
    struct Cfg { int mode, NC; const char* name; };
    const Cfg cfgs[] = {{0, 1, "single counter barrier"}, {1, 4, "distributed x4"}, {1, 8, "distributed x8"}, {1, 16, "distributed x16"},
                        {1, 32, "distributed x32"}, {2, 1, "ring hop (per hop)"}, {3, 1, "8->8 group hand-off (per hand-off)"}};
    for (const Cfg& c : cfgs)
        for (int grid : {148, 64}) {
            cudaMemset(ctr, 0, 64 * 1024);
            int mode = c.mode, NC = c.NC, it = iters;
            void* args[] = {&ctr, &mode, &NC, &it, &d};
            cudaError_t e = cudaLaunchCooperativeKernel((const void*)probe, dim3(grid), dim3(320), args, 0, 0);
            if (e == cudaSuccess) e = cudaDeviceSynchronize();
            long long h[148];
            cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
            double per = mx / (double)iters;
            if (c.mode == 2) per /= grid;
            if (c.mode == 3) per /= 2;
            printf("%-36s grid %3d: %7.0f cyc = %.2f us  %s\n", c.name, grid, per, per / 1965.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    return 0;
}
