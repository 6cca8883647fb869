// Routing of one mini-batch slice for ROW-SHARDED embedding tables (north_star: "embedding tables that exceed a single GPU are
// row-sharded with NCCL all-to-all"; table row r lives on rank r % world at local index r / world).  The reference keeps whole
// tables in one TF variable (model_zoo/DeepCTR/deepctr.py:105-126); these kernels are the device side of the data-parallel
// restatement of its train step (mamdr_b200/sharded.py): they replace ~40 small tensor ops per id column and step.
//
//   route_plan : for every local id i: owner = id % world, pos = number of EARLIER local ids with the same owner (stable),
//                slot[i] = owner * cap + pos;  send[slot[i]] = id / world, every other entry of send[world * cap] = -1
//                (fixed-capacity blocks: all all-to-all splits are static, -1 entries are padding the owners skip).
//   pack_rows  : dst[slot[i], :] = src[i, :] * scale   (gradient rows into the exchange buffer; padding rows are never read).
#include "common.cuh"

namespace {

constexpr int kPlanThreads = 1024;
constexpr int kMaxWorld = 64;

struct PlanArgs {
    const int32_t* ids[2];
    int32_t*       slot[2];
    int32_t*       send[2];
    int            n, world, cap, block;   // block = entries per owner block of the exchange buffer (>= cap)
};

__global__ void __launch_bounds__(kPlanThreads)
route_plan_kernel(const PlanArgs a) {
    __shared__ int warp_cnt[kPlanThreads / 32][kMaxWorld];
    __shared__ int base[kMaxWorld];
    const int32_t* ids = a.ids[blockIdx.x];
    int32_t* slot = a.slot[blockIdx.x];
    int32_t* send = a.send[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < a.world * a.cap; i += kPlanThreads) send[(i / a.cap) * a.block + (i % a.cap)] = -1;
    if (tid < a.world) base[tid] = 0;
    __syncthreads();
    for (int i0 = 0; i0 < a.n; i0 += kPlanThreads) {
        for (int w = lane; w < a.world; w += 32) warp_cnt[warp][w] = 0;
        __syncwarp();
        const int i = i0 + tid;
        const bool live = i < a.n;
        const int32_t id = live ? __ldg(ids + i) : 0;
        const int o = live ? (int)(id % a.world) : -1;
        const unsigned peers = __match_any_sync(0xffffffffu, o);
        const int in_warp = __popc(peers & ((1u << lane) - 1u));
        if (live && in_warp == 0) warp_cnt[warp][o] = __popc(peers);
        __syncthreads();
        if (live) {
            int pos = base[o] + in_warp;
            for (int w = 0; w < warp; ++w) pos += warp_cnt[w][o];
            const int s = o * a.block + pos;
            slot[i] = s + (a.block > a.cap ? (int)blockIdx.x * a.cap : 0);   // shared buffer: slots index the buffer that starts at send_a
            send[s] = id / a.world;
        }
        __syncthreads();
        if (tid < a.world) {
            int t = 0;
            for (int w = 0; w < kPlanThreads / 32; ++w) t += warp_cnt[w][tid];
            base[tid] += t;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
route_pack_rows_kernel(const float* __restrict__ src, const int64_t src_stride, const int32_t* __restrict__ slot, const int n,
                       const int dv, const float scale, float* __restrict__ dst) {
    const int64_t total = (int64_t)n * dv;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(v / dv), c = (int)(v - (int64_t)r * dv);
        float4 x = ldg_f4(src + r * src_stride + c * 4);
        x.x = __fmul_rn(x.x, scale); x.y = __fmul_rn(x.y, scale); x.z = __fmul_rn(x.z, scale); x.w = __fmul_rn(x.w, scale);
        *reinterpret_cast<float4*>(dst + (int64_t)__ldg(slot + r) * dv * 4 + c * 4) = x;
    }
}

// The owners' side of a FUSED exchange (both tables in one buffer: block r = [cap user entries | cap item entries]):
// out[e, :] = (column of e ? table_b : table_a)[recv_idx[e], :] for recv_idx[e] >= 0 (padding rows untouched), and the received
// id list split per table for the de-duplication: ids_a[e] = recv_idx[e] if e is a user entry else -1; ids_b likewise.
__global__ void __launch_bounds__(256)
route_gather2_kernel(const float* __restrict__ ta, const float* __restrict__ tb, const int32_t* __restrict__ recv_idx, const int n_entries,
                     const int cap, const int dv, float* __restrict__ out, int32_t* __restrict__ ids_a, int32_t* __restrict__ ids_b) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int e = warp; e < n_entries; e += nwarps) {
        const int32_t id = __ldg(recv_idx + e);
        const int col = (e / cap) & 1;
        if (lane == 0) {
            ids_a[e] = col == 0 ? id : -1;
            ids_b[e] = col == 1 ? id : -1;
        }
        if (id < 0) continue;
        const float* src = (col ? tb : ta) + (int64_t)id * dv * 4;
        float* dst = out + (int64_t)e * dv * 4;
        for (int c = lane; c < dv; c += 32) st_stream_f4(dst + c * 4, ldg_f4(src + c * 4));
    }
}

}  // namespace

extern "C" int mamdr_route_gather2(mamdr_ctx* ctx, const float* table_a, const float* table_b, const int32_t* recv_idx, int32_t world,
                                   int32_t cap, int32_t dim, float* out, int32_t* ids_a, int32_t* ids_b, mamdr_stream stream) {
    MAMDR_REQUIRE(ctx, ctx != nullptr, MAMDR_E_INVALID, "ctx is NULL");
    MAMDR_REQUIRE(ctx, recv_idx && out && ids_a && ids_b, MAMDR_E_INVALID, "NULL pointer");
    MAMDR_REQUIRE(ctx, world >= 1 && cap >= 1 && dim > 0 && dim % 4 == 0, MAMDR_E_INVALID, "bad world / cap / dim");
    MAMDR_REQUIRE(ctx, aligned16(out) && (!table_a || aligned16(table_a)) && (!table_b || aligned16(table_b)), MAMDR_E_INVALID, "misaligned table / out");
    const int n_entries = world * 2 * cap;
    const int want = (n_entries + 7) / 8, capg = ctx->sm_count * 8;
    route_gather2_kernel<<<want < capg ? want : capg, 256, 0, (cudaStream_t)stream>>>(table_a, table_b, recv_idx, n_entries, cap, dim / 4, out, ids_a, ids_b);
    MAMDR_LAUNCH_OK(ctx);
    return MAMDR_OK;
}

extern "C" int mamdr_route_plan(mamdr_ctx* ctx, const int32_t* ids_a, const int32_t* ids_b, int32_t n, int32_t world, int32_t cap,
                                int32_t block, int32_t* slot_a, int32_t* slot_b, int32_t* send_a, int32_t* send_b, mamdr_stream stream) {
    MAMDR_REQUIRE(ctx, ctx != nullptr, MAMDR_E_INVALID, "ctx is NULL");
    MAMDR_REQUIRE(ctx, world >= 1 && world <= kMaxWorld, MAMDR_E_INVALID, "world must be in [1, %d]", kMaxWorld);
    MAMDR_REQUIRE(ctx, n >= 0 && cap >= 1 && n <= cap, MAMDR_E_INVALID, "need 0 <= n <= cap (every owner block can take the whole slice)");
    MAMDR_REQUIRE(ctx, send_a && (n == 0 || (ids_a && slot_a)), MAMDR_E_INVALID, "NULL pointer");
    const bool two = ids_b != nullptr || send_b != nullptr;
    if (two) MAMDR_REQUIRE(ctx, send_b && (n == 0 || (ids_b && slot_b)), MAMDR_E_INVALID, "NULL pointer (second column)");
    PlanArgs a;
    a.ids[0] = ids_a; a.slot[0] = slot_a; a.send[0] = send_a;
    a.ids[1] = ids_b; a.slot[1] = slot_b; a.send[1] = send_b;
    MAMDR_REQUIRE(ctx, block == cap || (block == 2 * cap && two && send_b == send_a + cap), MAMDR_E_INVALID,
                  "block must be cap (separate buffers) or 2 * cap with send_b == send_a + cap (one shared buffer)");
    a.n = n; a.world = world; a.cap = cap; a.block = block;
    route_plan_kernel<<<two ? 2 : 1, kPlanThreads, 0, (cudaStream_t)stream>>>(a);
    MAMDR_LAUNCH_OK(ctx);
    return MAMDR_OK;
}

extern "C" int mamdr_route_pack_rows(mamdr_ctx* ctx, const float* src, int64_t src_stride, const int32_t* slot, int32_t n, int32_t dim,
                                     float scale, float* dst, mamdr_stream stream) {
    MAMDR_REQUIRE(ctx, ctx != nullptr, MAMDR_E_INVALID, "ctx is NULL");
    MAMDR_REQUIRE(ctx, n >= 0, MAMDR_E_INVALID, "negative size");
    if (n == 0) return MAMDR_OK;
    MAMDR_REQUIRE(ctx, src && slot && dst, MAMDR_E_INVALID, "NULL pointer");
    MAMDR_REQUIRE(ctx, dim > 0 && dim % 4 == 0 && src_stride >= dim && src_stride % 4 == 0, MAMDR_E_INVALID, "dim / src_stride must be multiples of 4");
    MAMDR_REQUIRE(ctx, aligned16(src) && aligned16(dst), MAMDR_E_INVALID, "src/dst must be 16-byte aligned");
    const int dv = dim / 4;
    const int64_t want = ceil_div64((int64_t)n * dv, 256);
    const int64_t capg = (int64_t)ctx->sm_count * 8;
    route_pack_rows_kernel<<<(unsigned)(want < capg ? want : capg), 256, 0, (cudaStream_t)stream>>>(src, src_stride, slot, n, dv, scale, dst);
    MAMDR_LAUNCH_OK(ctx);
    return MAMDR_OK;
}
