// K6: sparse embedding-gradient de-duplication: sort-by-id + warp-per-segment sum.
//
// Replaces TF's _deduplicate_indexed_slices (tf.unique + unsorted_segment_sum) that runs inside
// AdamOptimizer.apply_gradients for the IndexedSlices gradient of tf.gather
// (/root/reference/model_zoo/DeepCTR/deepctr.py:54-55,125-126).
//
//   pass 1 (one CTA): 64-bit keys (id << 32 | batch position) -> in-smem bitonic sort (unique keys
//           => the order is the stable order), head flags, exclusive scan -> sorted unique ids,
//           segment starts, permutation.
//   pass 2 (warp per unique id): rows of one id are added sequentially in batch order, each lane
//           owning 4-float column groups -> deterministic, bit-identical to numpy add.at.
// No atomics.  n <= 8192 per call (one mini-batch).
#include "common.cuh"

namespace {

constexpr int kSortThreads = 1024;
constexpr int kMaxN = 8192;

__global__ void __launch_bounds__(kSortThreads) sort_unique_kernel(const __grid_constant__ DedupArgs a) {
    extern __shared__ __align__(16) unsigned long long keys[];  // [npow2]
    __shared__ int scan_part[kSortThreads];
    const DedupJob& J = a.job[blockIdx.x];
    const int32_t* __restrict__ ids = J.ids;
    int32_t* __restrict__ uniq_ids = J.uniq_ids;
    int32_t* __restrict__ perm = J.perm;
    int32_t* __restrict__ seg_start = J.seg_start;
    const int n = a.n, npow2 = a.npow2;
    const int tid = threadIdx.x;
    __shared__ int n_valid;
    if (tid == 0) n_valid = 0;
    // negative ids are padding (fixed-capacity all-to-all buffers of the row-sharded tables): they sort behind every real id
    for (int i = tid; i < npow2; i += kSortThreads)
        keys[i] = (i < n && ids[i] >= 0) ? (((unsigned long long)(uint32_t)ids[i]) << 32) | (uint32_t)i : ~0ull;
    __syncthreads();
    for (int k = 2; k <= npow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < npow2; i += kSortThreads) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long x = keys[i], y = keys[ixj];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { keys[i] = y; keys[ixj] = x; }
                }
            }
            __syncthreads();
        }
    }
    // head flags + exclusive scan (each thread owns a contiguous chunk)
    const int per = (n + kSortThreads - 1) / kSortThreads;
    const int beg = tid * per, end = min(n, beg + per);
    int cnt = 0;
    for (int i = beg; i < end; ++i) {
        if (keys[i] == ~0ull) break;
        const uint32_t id = (uint32_t)(keys[i] >> 32);
        const bool head = (i == 0) || ((uint32_t)(keys[i - 1] >> 32) != id);
        cnt += head ? 1 : 0;
        if (i + 1 == n || keys[i + 1] == ~0ull) n_valid = i + 1;   // exactly one thread sees the last real entry
    }
    scan_part[tid] = cnt;
    __syncthreads();
    // inclusive Hillis-Steele over 1024 partials
    for (int off = 1; off < kSortThreads; off <<= 1) {
        const int add = tid >= off ? scan_part[tid - off] : 0;
        __syncthreads();
        scan_part[tid] += add;
        __syncthreads();
    }
    int seg = scan_part[tid] - cnt;  // exclusive prefix
    for (int i = beg; i < end; ++i) {
        if (keys[i] == ~0ull) break;
        const uint32_t id = (uint32_t)(keys[i] >> 32);
        const bool head = (i == 0) || ((uint32_t)(keys[i - 1] >> 32) != id);
        if (head) {
            uniq_ids[seg] = (int32_t)id;
            seg_start[seg] = i;
            ++seg;
        }
        perm[i] = (int32_t)(keys[i] & 0xffffffffull);
    }
    if (tid == kSortThreads - 1) {
        const int total = scan_part[tid];
        J.n_uniq[0] = total;
        seg_start[total] = n_valid;
    }
}

// warp per unique id; the rows of one id are added sequentially in batch order (bit-identical to numpy add.at).  The
// permutation entries of a segment are fetched 32 at a time and the row loads run 4 deep ahead of the ordered adds.
__global__ void __launch_bounds__(256) segment_sum_kernel(const __grid_constant__ DedupArgs a) {
    const DedupJob& J = a.job[blockIdx.y];
    const float* __restrict__ grad_rows = J.grad_rows;
    const int64_t grad_stride = J.grad_stride;
    const int dim = J.dim;
    const int32_t* __restrict__ perm = J.perm;
    const int32_t* __restrict__ seg_start = J.seg_start;
    float* __restrict__ uniq_rows = J.uniq_rows;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int nu = J.n_uniq[0];
    for (int k = warp; k < nu; k += nwarps) {
        const int s = seg_start[k], e = seg_start[k + 1];
        for (int base = s; base < e; base += 32) {
            const int cnt = min(32, e - base);
            const int myp = lane < cnt ? perm[base + lane] : 0;
            for (int cc = 0; cc < dim; cc += 128) {
                const int c = cc + lane * 4;
                const bool active = c < dim;
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                if (base != s && active) acc = *reinterpret_cast<const float4*>(uniq_rows + (int64_t)k * dim + c);
                for (int i = 0; i < cnt; i += 4) {
                    int p[4];
                    float4 v[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) p[q] = __shfl_sync(0xffffffffu, myp, min(i + q, cnt - 1));
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        v[q] = (active && i + q < cnt) ? *reinterpret_cast<const float4*>(grad_rows + (int64_t)p[q] * grad_stride + c)
                                                       : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (i + q >= cnt) break;
                        if (base == s && i + q == 0) {
                            acc = v[q];
                        } else {
                            acc.x = __fadd_rn(acc.x, v[q].x); acc.y = __fadd_rn(acc.y, v[q].y);
                            acc.z = __fadd_rn(acc.z, v[q].z); acc.w = __fadd_rn(acc.w, v[q].w);
                        }
                    }
                }
                if (active) *reinterpret_cast<float4*>(uniq_rows + (int64_t)k * dim + c) = acc;
            }
        }
    }
}

inline size_t al(size_t x) { return (x + 255) / 256 * 256; }

// ================================================================================================================
// Large-n path (n > kMaxN): multi-CTA.  Stable LSD radix sort of (id << 32 | position) by id (4 passes of 8 bits: per-tile
// digit histograms, one scan, a stable scatter), head flags + scan -> unique ids / segment starts / permutation, then the
// HBM-bound part: the n gradient rows are read ONCE by warps that each own a window of kWin consecutive sorted positions.
// Summation order (deterministic, restated by the numpy oracle in tests/test_gpu_kernels.py): inside a window the rows of
// an id are added sequentially in batch order; the window partials of an id that spans several windows are added in window
// order within groups of 32 windows, then the group sums in order.  An id whose rows lie inside one window: exactly the
// add.at order of the single-CTA path.
// ================================================================================================================
constexpr int kTile = 4096;      // keys per radix tile (256 threads x 16 rounds)
constexpr int kRadixThreads = 256;
constexpr int kWin = 256;        // sorted positions per window of the segment sum

__global__ void __launch_bounds__(256) lg_key_init_kernel(const int32_t* __restrict__ ids, unsigned long long* __restrict__ keys, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = ids[i] >= 0 ? (((unsigned long long)(uint32_t)ids[i]) << 32) | (unsigned long long)(uint32_t)i : ~0ull;
}

// per-tile digit histogram -> hist[digit * n_tiles + tile]
__global__ void __launch_bounds__(kRadixThreads) lg_hist_kernel(const unsigned long long* __restrict__ keys, int64_t n, int shift,
                                                                int* __restrict__ hist, int n_tiles) {
    __shared__ int cnt[256];
    const int t = threadIdx.x;
    cnt[t] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kTile;
    for (int r = 0; r < kTile / kRadixThreads; ++r) {
        const int64_t i = base + r * kRadixThreads + t;
        if (i < n) atomicAdd(&cnt[(int)((keys[i] >> shift) & 0xffull)], 1);
    }
    __syncthreads();
    hist[t * n_tiles + blockIdx.x] = cnt[t];
}

// exclusive scan of an int array by ONE block (the per-tile histograms: 256 * n_tiles entries; the per-block head counts):
// tiles of 4096 entries, 4 consecutive ints per thread (coalesced), warp shuffles + one cross-warp step per tile
__global__ void __launch_bounds__(1024) lg_scan_kernel(int* __restrict__ a, int64_t m, int* __restrict__ total_out) {
    __shared__ int wsum[32];
    __shared__ int carry_s;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (t == 0) carry_s = 0;
    __syncthreads();
    for (int64_t base = 0; base < m; base += 4096) {
        const int64_t i0 = base + (int64_t)t * 4;
        int v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = i0 + q < m ? a[i0 + q] : 0;
        const int mine = v[0] + v[1] + v[2] + v[3];
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        if (lane == 31) wsum[w] = incl;
        __syncthreads();
        const int carry = carry_s;
        if (w == 0) {
            int ws = wsum[lane];
            int wi = ws;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int up = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += up;
            }
            wsum[lane] = wi - ws;                      // exclusive prefix of the warp sums
            if (lane == 31) carry_s = carry + wi;      // running total after this tile
        }
        __syncthreads();
        int run = carry + wsum[w] + incl - mine;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (i0 + q < m) a[i0 + q] = run;
            run += v[q];
        }
        __syncthreads();
    }
    if (t == 0 && total_out) *total_out = carry_s;
}

// the histogram scan, parallel over the digits: hist is [digit][tile]; block d scans its row (exclusive) and publishes the row
// total; a second launch adds the totals of the smaller digits.  (One 1024-thread block over 256 * n_tiles entries is a
// serial tail of ~60 us per pass at 2 Mi keys.)
__global__ void __launch_bounds__(256) lg_digit_scan_kernel(int* __restrict__ hist, int n_tiles, int* __restrict__ digit_total) {
    __shared__ int wsum[8];
    __shared__ int carry_s;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    int* row = hist + (int64_t)blockIdx.x * n_tiles;
    if (t == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < n_tiles; base += 256) {
        const int i = base + t;
        const int v = i < n_tiles ? row[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        if (lane == 31) wsum[w] = incl;
        __syncthreads();
        const int carry = carry_s;
        int before = 0;
        for (int q = 0; q < w; ++q) before += wsum[q];
        if (i < n_tiles) row[i] = carry + before + incl - v;
        __syncthreads();
        if (t == 255) carry_s = carry + before + incl;
        __syncthreads();
    }
    if (t == 0) digit_total[blockIdx.x] = carry_s;
}
__global__ void __launch_bounds__(256) lg_digit_base_kernel(int* __restrict__ hist, int n_tiles, const int* __restrict__ digit_total) {
    __shared__ int base_s;
    if (threadIdx.x == 0) {
        int b = 0;
        for (int d = 0; d < (int)blockIdx.x; ++d) b += digit_total[d];
        base_s = b;
    }
    __syncthreads();
    int* row = hist + (int64_t)blockIdx.x * n_tiles;
    for (int i = threadIdx.x; i < n_tiles; i += 256) row[i] += base_s;
}

// stable scatter of one tile: 16 rounds of 256 keys in thread order; rank of a key = keys of the same digit in earlier
// tiles (scanned histogram) + earlier rounds + earlier warps of the round + earlier lanes of the warp
__global__ void __launch_bounds__(kRadixThreads) lg_scatter_kernel(const unsigned long long* __restrict__ in, unsigned long long* __restrict__ out,
                                                                   int64_t n, int shift, const int* __restrict__ hist, int n_tiles) {
    __shared__ int running[256];          // position of the next key of each digit
    __shared__ int warp_cnt[8][256];
    const int t = threadIdx.x, w = t >> 5, lane = t & 31;
    running[t] = hist[t * n_tiles + blockIdx.x];
#pragma unroll
    for (int q = 0; q < 8; ++q) warp_cnt[q][t] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kTile;
    for (int r = 0; r < kTile / kRadixThreads; ++r) {
        const int64_t i = base + r * kRadixThreads + t;
        const bool valid = i < n;
        const unsigned long long key = valid ? in[i] : 0ull;
        const int d = valid ? (int)((key >> shift) & 0xffull) : 256 + lane;   // invalid lanes never match anyone
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const int rank_in_warp = __popc(peers & ((1u << lane) - 1u));
        if (valid && rank_in_warp == 0) warp_cnt[w][d] = __popc(peers);
        __syncthreads();
        if (valid) {
            int pos = running[d] + rank_in_warp;
            for (int q = 0; q < w; ++q) pos += warp_cnt[q][d];
            out[pos] = key;
        }
        __syncthreads();
        {   // digit t: advance by this round's keys, clear the per-warp counts
            int add = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) { add += warp_cnt[q][t]; warp_cnt[q][t] = 0; }
            running[t] += add;
        }
        __syncthreads();
    }
}

// head flags: per block of 1024 sorted keys, the number of segment heads among the real (non-padding) keys
__global__ void __launch_bounds__(1024) lg_head_count_kernel(const unsigned long long* __restrict__ keys, int64_t n, int* __restrict__ blk_heads,
                                                             int* __restrict__ n_valid) {
    __shared__ int cnt;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    bool head = false;
    if (i < n && keys[i] != ~0ull) {
        head = i == 0 || (uint32_t)(keys[i - 1] >> 32) != (uint32_t)(keys[i] >> 32);
        if (i + 1 == n || keys[i + 1] == ~0ull) *n_valid = (int)(i + 1);
    }
    const unsigned b = __ballot_sync(0xffffffffu, head);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(&cnt, __popc(b));
    __syncthreads();
    if (threadIdx.x == 0) blk_heads[blockIdx.x] = cnt;
}

// unique ids, segment starts, permutation and the segment index of every sorted position
__global__ void __launch_bounds__(1024) lg_head_write_kernel(const unsigned long long* __restrict__ keys, int64_t n, const int* __restrict__ blk_base,
                                                             int32_t* __restrict__ uniq_ids, int32_t* __restrict__ seg_start, int32_t* __restrict__ perm,
                                                             int32_t* __restrict__ seg_of) {
    __shared__ int wsum[32];
    const int t = threadIdx.x, w = t >> 5, lane = t & 31;
    const int64_t i = (int64_t)blockIdx.x * 1024 + t;
    const bool real = i < n && keys[i] != ~0ull;
    const bool head = real && (i == 0 || (uint32_t)(keys[i - 1] >> 32) != (uint32_t)(keys[i] >> 32));
    const unsigned b = __ballot_sync(0xffffffffu, head);
    if (lane == 0) wsum[w] = __popc(b);
    __syncthreads();
    int before = blk_base[blockIdx.x];
    for (int q = 0; q < w; ++q) before += wsum[q];
    const int seg = before + __popc(b & ((1u << lane) - 1u)) + (head ? 1 : 0) - 1;   // index of the segment this position belongs to
    if (real) {
        perm[i] = (int32_t)(keys[i] & 0xffffffffull);
        seg_of[i] = seg;
        if (head) { uniq_ids[seg] = (int32_t)(keys[i] >> 32); seg_start[seg] = (int32_t)i; }
    }
}

__global__ void lg_finish_kernel(const int* __restrict__ total, const int* __restrict__ n_valid, int32_t* __restrict__ n_uniq, int32_t* __restrict__ seg_start) {
    n_uniq[0] = *total;
    seg_start[*total] = *n_valid;
}

// warp per window of kWin sorted positions: pieces of segments inside the window are summed sequentially in batch order.  A
// piece that is the whole segment goes to uniq_rows; a piece of a segment that started in an earlier window goes to
// part[w][0], a piece of a segment that continues into the next window to part[w][1] (flags in pflag[w]).  The window's
// permutation and segment indices are staged in shared memory first, so the row loads run 16 deep across piece boundaries
// (cold ids make one-row pieces: a per-piece dependent load chain would serialise the whole window).
__global__ void __launch_bounds__(256, 2) lg_window_sum_kernel(const float* __restrict__ grad_rows, int64_t grad_stride, int dim,
                                                               const int32_t* __restrict__ perm, const int32_t* __restrict__ seg_of,
                                                               const int32_t* __restrict__ seg_start, const int* __restrict__ n_valid_p,
                                                               float* __restrict__ uniq_rows, float* __restrict__ part, int* __restrict__ pflag) {
    __shared__ int s_perm[8][kWin], s_seg[8][kWin];
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_valid = *n_valid_p;
    const int64_t w_beg = w * kWin;
    if (w_beg >= n_valid) return;
    const int64_t w_end = w_beg + kWin < n_valid ? w_beg + kWin : n_valid;
    const int total = (int)(w_end - w_beg);
    for (int q = lane; q < total; q += 32) { s_perm[wl][q] = perm[w_beg + q]; s_seg[wl][q] = seg_of[w_beg + q]; }
    __syncwarp();
    const int first_seg = s_seg[wl][0], last_seg = s_seg[wl][total - 1];
    const bool from_before = seg_start[first_seg] < w_beg, goes_on = seg_start[last_seg + 1] > w_end;
    for (int cc = 0; cc < dim; cc += 128) {
        const int c = cc + lane * 4;
        const bool active = c < dim;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int cur = first_seg;
        bool fresh = true;
        auto flush = [&](int seg) {
            float* dst = (seg == first_seg && from_before) ? part + ((int64_t)w * 2 + 0) * dim
                       : (seg == last_seg && goes_on)      ? part + ((int64_t)w * 2 + 1) * dim
                                                           : uniq_rows + (int64_t)seg * dim;
            if (active) *reinterpret_cast<float4*>(dst + c) = acc;
        };
        auto load8 = [&](int r0, float4 (&v)[8]) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
                v[q] = (active && r0 + q < total) ? __ldcs(reinterpret_cast<const float4*>(grad_rows + (int64_t)s_perm[wl][r0 + q] * grad_stride + c))
                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        auto add8 = [&](int r0, const float4 (&v)[8]) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (r0 + q >= total) break;
                const int seg = s_seg[wl][r0 + q];
                if (seg != cur) { flush(cur); cur = seg; fresh = true; }
                if (fresh) {
                    acc = v[q];
                    fresh = false;
                } else {
                    acc.x = __fadd_rn(acc.x, v[q].x); acc.y = __fadd_rn(acc.y, v[q].y);
                    acc.z = __fadd_rn(acc.z, v[q].z); acc.w = __fadd_rn(acc.w, v[q].w);
                }
            }
        };
        float4 va[8], vb[8];
        load8(0, va);
        for (int r0 = 0; r0 < total; r0 += 16) {
            if (r0 + 8 < total) load8(r0 + 8, vb);
            add8(r0, va);
            if (r0 + 16 < total) load8(r0 + 16, va);
            if (r0 + 8 < total) add8(r0 + 8, vb);
        }
        flush(cur);
    }
    if (lane == 0) pflag[w] = (from_before ? 1 : 0) | ((goes_on && !(from_before && first_seg == last_seg)) ? 2 : 0);
}

// Segments that span several windows: the window pieces are first summed in GROUPS of kGroup consecutive windows (offset
// from the segment's first window in [kGroup g, kGroup g + kGroup), added in window order, in place into the group's first
// piece), then the group sums are added in order.  A hot id (Zipf head) spans thousands of windows: one level would be one
// long serial chain.
constexpr int kGroup = 32;

__device__ __forceinline__ void lg_sum_pieces(const float* part, int dim, int lane, int64_t x_first, int64_t x_last, int64_t step, float4* acc, int cc) {
    const int c = cc + lane * 4;
    for (int64_t x0 = x_first; x0 <= x_last; x0 += 16 * step) {   // the (independent) loads run 16 ahead of the ordered adds
        float4 v[16];
#pragma unroll
        for (int q = 0; q < 16; ++q)
            v[q] = x0 + q * step <= x_last ? __ldcs(reinterpret_cast<const float4*>(part + ((int64_t)(x0 + q * step) * 2 + 0) * dim + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            if (x0 + q * step > x_last) break;
            acc->x = __fadd_rn(acc->x, v[q].x); acc->y = __fadd_rn(acc->y, v[q].y); acc->z = __fadd_rn(acc->z, v[q].z); acc->w = __fadd_rn(acc->w, v[q].w);
        }
    }
}

// level 1: warp per window; a window leads a group when its piece has offset 0 (the segment starts here and goes on: slot 1)
// or an offset that is a multiple of kGroup (slot 0)
__global__ void __launch_bounds__(256) lg_window_group_kernel(int dim, const int32_t* __restrict__ seg_of, const int32_t* __restrict__ seg_start,
                                                              const int* __restrict__ n_valid_p, float* __restrict__ part, const int* __restrict__ pflag) {
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_valid = *n_valid_p;
    const int64_t w_beg = w * kWin;
    if (w_beg >= n_valid) return;
    const int flags = pflag[w];
    for (int role = 0; role < 2; ++role) {
        int seg;
        int slot;
        if (role == 0) {          // the segment that starts in this window and continues
            if (!(flags & 2)) continue;
            seg = seg_of[w_beg + kWin - 1];
            slot = 1;
        } else {                  // the segment that came from an earlier window
            if (!(flags & 1)) continue;
            seg = seg_of[w_beg];
            const int64_t w_first = seg_start[seg] / kWin;
            if ((w - w_first) % kGroup != 0) continue;
            slot = 0;
        }
        const int64_t w_last = ((int64_t)seg_start[seg + 1] - 1) / kWin;
        const int64_t x_last = w + kGroup - 1 < w_last ? w + kGroup - 1 : w_last;
        if (x_last <= w) continue;
        for (int cc = 0; cc < dim; cc += 128) {
            const int c = cc + lane * 4;
            if (c >= dim) continue;
            float* mine = part + ((int64_t)w * 2 + slot) * dim + c;
            float4 acc = *reinterpret_cast<const float4*>(mine);
            lg_sum_pieces(part, dim, lane, w + 1, x_last, 1, &acc, cc);
            *reinterpret_cast<float4*>(mine) = acc;
        }
    }
}

// level 2: warp per window in which a spanning segment starts: the group sums in order
__global__ void __launch_bounds__(256) lg_window_merge_kernel(int dim, const int32_t* __restrict__ seg_of, const int32_t* __restrict__ seg_start,
                                                              const int* __restrict__ n_valid_p, float* __restrict__ uniq_rows,
                                                              const float* __restrict__ part, const int* __restrict__ pflag) {
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_valid = *n_valid_p;
    const int64_t w_beg = w * kWin;
    if (w_beg >= n_valid || !(pflag[w] & 2)) return;
    const int seg = seg_of[w_beg + kWin - 1];            // < n_valid because the last segment of the window goes on
    const int64_t w_last = ((int64_t)seg_start[seg + 1] - 1) / kWin;
    for (int cc = 0; cc < dim; cc += 128) {
        const int c = cc + lane * 4;
        if (c >= dim) continue;
        float4 acc = *reinterpret_cast<const float4*>(part + ((int64_t)w * 2 + 1) * dim + c);
        lg_sum_pieces(part, dim, lane, w + kGroup, w_last, kGroup, &acc, cc);
        *reinterpret_cast<float4*>(uniq_rows + (int64_t)seg * dim + c) = acc;
    }
}

struct LargeWs { size_t keys[2], hist, dtot, blk, scal, perm, seg_start, seg_of, part, pflag, total; };
inline LargeWs large_ws(int64_t n, int dim) {
    LargeWs w;
    size_t off = 0;
    auto take = [&](size_t b) { size_t o = off; off += al(b); return o; };
    const int64_t n_tiles = (n + kTile - 1) / kTile, n_blk = (n + 1023) / 1024, n_win = (n + kWin - 1) / kWin;
    w.keys[0] = take((size_t)n * 8); w.keys[1] = take((size_t)n * 8);
    w.hist = take((size_t)256 * n_tiles * 4);
    w.dtot = take(256 * 4);
    w.blk = take((size_t)n_blk * 4);
    w.scal = take(64);
    w.perm = take((size_t)n * 4); w.seg_start = take((size_t)(n + 1) * 4); w.seg_of = take((size_t)n * 4);
    w.part = take((size_t)n_win * 2 * dim * 4); w.pflag = take((size_t)n_win * 4);
    w.total = off;
    return w;
}

}  // namespace

extern "C" size_t mamdr_scatter_large_workspace_bytes(int64_t n, int32_t dim) {
    if (n < 0 || dim <= 0) return 0;
    return large_ws(n, dim).total;
}

extern "C" int mamdr_scatter_dedup_large_f32(mamdr_ctx* ctx, const int32_t* ids, const float* grad_rows, int64_t grad_stride, int64_t n,
                                             int32_t dim, int32_t* uniq_ids, float* uniq_rows, int32_t* n_uniq, void* ws_, size_t ws_bytes,
                                             mamdr_stream stream) {
    MAMDR_REQUIRE(ctx, ctx != nullptr, MAMDR_E_INVALID, "ctx is NULL");
    MAMDR_REQUIRE(ctx, n >= 1 && n < (1ll << 31), MAMDR_E_UNSUPPORTED, "n=%lld outside 1..2^31-1", (long long)n);
    MAMDR_REQUIRE(ctx, ids && grad_rows && uniq_ids && uniq_rows && n_uniq && ws_, MAMDR_E_INVALID, "NULL pointer");
    MAMDR_REQUIRE(ctx, dim > 0 && dim % 4 == 0 && grad_stride >= dim && grad_stride % 4 == 0, MAMDR_E_INVALID, "bad dim/stride");
    MAMDR_REQUIRE(ctx, aligned16(grad_rows) && aligned16(uniq_rows) && aligned16(ws_), MAMDR_E_INVALID, "misaligned pointer");
    const LargeWs w = large_ws(n, dim);
    MAMDR_REQUIRE(ctx, ws_bytes >= w.total, MAMDR_E_WORKSPACE, "workspace too small: %zu < %zu", ws_bytes, w.total);
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* ws = (unsigned char*)ws_;
    unsigned long long* keys[2] = {(unsigned long long*)(ws + w.keys[0]), (unsigned long long*)(ws + w.keys[1])};
    int* hist = (int*)(ws + w.hist);
    int* dtot = (int*)(ws + w.dtot);
    int* blk = (int*)(ws + w.blk);
    int* scal = (int*)(ws + w.scal);   // [0] = number of unique ids, [1] = number of real (non-padding) entries
    int32_t* perm = (int32_t*)(ws + w.perm);
    int32_t* seg_start = (int32_t*)(ws + w.seg_start);
    int32_t* seg_of = (int32_t*)(ws + w.seg_of);
    float* part = (float*)(ws + w.part);
    int* pflag = (int*)(ws + w.pflag);
    const int n_tiles = (int)((n + kTile - 1) / kTile), n_blk = (int)((n + 1023) / 1024), n_win = (int)((n + kWin - 1) / kWin);
    MAMDR_CUDA_OK(ctx, cudaMemsetAsync(scal, 0, 64, st));
    lg_key_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ids, keys[0], n);
    MAMDR_LAUNCH_OK(ctx);
    int cur = 0;
    for (int pass = 0; pass < 4; ++pass) {   // stable LSD passes over the id (bits 32..63); padding keys (all ones) end up last
        const int shift = 32 + 8 * pass;
        lg_hist_kernel<<<n_tiles, kRadixThreads, 0, st>>>(keys[cur], n, shift, hist, n_tiles);
        MAMDR_LAUNCH_OK(ctx);
        lg_digit_scan_kernel<<<256, 256, 0, st>>>(hist, n_tiles, dtot);
        MAMDR_LAUNCH_OK(ctx);
        lg_digit_base_kernel<<<256, 256, 0, st>>>(hist, n_tiles, dtot);
        MAMDR_LAUNCH_OK(ctx);
        lg_scatter_kernel<<<n_tiles, kRadixThreads, 0, st>>>(keys[cur], keys[cur ^ 1], n, shift, hist, n_tiles);
        MAMDR_LAUNCH_OK(ctx);
        cur ^= 1;
    }
    lg_head_count_kernel<<<n_blk, 1024, 0, st>>>(keys[cur], n, blk, scal + 1);
    MAMDR_LAUNCH_OK(ctx);
    lg_scan_kernel<<<1, 1024, 0, st>>>(blk, n_blk, scal);
    MAMDR_LAUNCH_OK(ctx);
    lg_head_write_kernel<<<n_blk, 1024, 0, st>>>(keys[cur], n, blk, uniq_ids, seg_start, perm, seg_of);
    MAMDR_LAUNCH_OK(ctx);
    lg_finish_kernel<<<1, 1, 0, st>>>(scal, scal + 1, n_uniq, seg_start);
    MAMDR_LAUNCH_OK(ctx);
    const unsigned wblocks = (unsigned)((n_win + 7) / 8);
    lg_window_sum_kernel<<<wblocks, 256, 0, st>>>(grad_rows, grad_stride, dim, perm, seg_of, seg_start, scal + 1, uniq_rows, part, pflag);
    MAMDR_LAUNCH_OK(ctx);
    lg_window_group_kernel<<<wblocks, 256, 0, st>>>(dim, seg_of, seg_start, scal + 1, part, pflag);
    MAMDR_LAUNCH_OK(ctx);
    lg_window_merge_kernel<<<wblocks, 256, 0, st>>>(dim, seg_of, seg_start, scal + 1, uniq_rows, part, pflag);
    MAMDR_LAUNCH_OK(ctx);
    return MAMDR_OK;
}

namespace {
}  // namespace

int mamdr_scatter_init_kernels(mamdr_ctx* ctx) {
    MAMDR_CUDA_OK(ctx, cudaFuncSetAttribute(sort_unique_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            kMaxN * (int)sizeof(unsigned long long)));
    return MAMDR_OK;
}

extern "C" int64_t mamdr_scatter_max_n(void) { return kMaxN; }

extern "C" size_t mamdr_scatter_workspace_bytes(int64_t n) {
    if (n < 0) return 0;
    return al((size_t)n * 4) + al((size_t)(n + 1) * 4);
}

// de-duplicate `n_jobs` (<= 2: the user and the item table of one mini-batch) id lists of n entries in two launches
int mamdr_scatter_dedup_jobs(mamdr_ctx* ctx, const DedupJob* jobs, int n_jobs, int n, cudaStream_t st) {
    MAMDR_REQUIRE(ctx, ctx != nullptr && jobs != nullptr, MAMDR_E_INVALID, "ctx / jobs is NULL");
    MAMDR_REQUIRE(ctx, n_jobs >= 1 && n_jobs <= 2, MAMDR_E_INVALID, "n_jobs must be 1 or 2");
    MAMDR_REQUIRE(ctx, n >= 1 && n <= kMaxN, MAMDR_E_UNSUPPORTED, "n=%d outside 1..%d", n, kMaxN);
    DedupArgs a;
    memset(&a, 0, sizeof(a));
    for (int q = 0; q < n_jobs; ++q) {
        const DedupJob& J = jobs[q];
        MAMDR_REQUIRE(ctx, J.ids && J.grad_rows && J.uniq_ids && J.uniq_rows && J.n_uniq && J.perm && J.seg_start, MAMDR_E_INVALID, "NULL pointer");
        MAMDR_REQUIRE(ctx, J.dim > 0 && J.dim % 4 == 0 && J.grad_stride >= J.dim && J.grad_stride % 4 == 0, MAMDR_E_INVALID, "bad dim/stride");
        MAMDR_REQUIRE(ctx, aligned16(J.grad_rows) && aligned16(J.uniq_rows), MAMDR_E_INVALID, "misaligned pointer");
        a.job[q] = J;
    }
    a.n = n;
    a.npow2 = 1;
    while (a.npow2 < n) a.npow2 <<= 1;
    sort_unique_kernel<<<n_jobs, kSortThreads, (size_t)a.npow2 * sizeof(unsigned long long), st>>>(a);
    MAMDR_LAUNCH_OK(ctx);
    segment_sum_kernel<<<dim3((n + 7) / 8, n_jobs), 256, 0, st>>>(a);   // n bounds the number of unique ids
    MAMDR_LAUNCH_OK(ctx);
    return MAMDR_OK;
}

void mamdr_scatter_job_ws(DedupJob* job, void* ws, int64_t n) {
    job->perm = (int32_t*)ws;
    job->seg_start = (int32_t*)((unsigned char*)ws + al((size_t)n * 4));
}

extern "C" int mamdr_scatter_dedup_f32(mamdr_ctx* ctx, const int32_t* ids, const float* grad_rows, int64_t grad_stride,
                                       int64_t n, int32_t dim, int32_t* uniq_ids, float* uniq_rows, int32_t* n_uniq,
                                       void* ws, size_t ws_bytes, mamdr_stream stream) {
    MAMDR_REQUIRE(ctx, ctx != nullptr, MAMDR_E_INVALID, "ctx is NULL");
    MAMDR_REQUIRE(ctx, n >= 0 && n <= kMaxN, MAMDR_E_UNSUPPORTED, "n=%lld exceeds %d", (long long)n, kMaxN);
    MAMDR_REQUIRE(ctx, n_uniq != nullptr, MAMDR_E_INVALID, "n_uniq_dev is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        MAMDR_CUDA_OK(ctx, cudaMemsetAsync(n_uniq, 0, 4, st));
        return MAMDR_OK;
    }
    MAMDR_REQUIRE(ctx, ws && aligned16(ws), MAMDR_E_INVALID, "workspace NULL or misaligned");
    MAMDR_REQUIRE(ctx, ws_bytes >= mamdr_scatter_workspace_bytes(n), MAMDR_E_WORKSPACE, "workspace too small");
    DedupJob J{ids, grad_rows, grad_stride, dim, uniq_ids, uniq_rows, n_uniq, nullptr, nullptr};
    mamdr_scatter_job_ws(&J, ws, n);
    return mamdr_scatter_dedup_jobs(ctx, &J, 1, (int)n, st);
}
