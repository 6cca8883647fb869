// Tensor-map (CUtensorMap) cache shared by the tcgen05 paths (mlp_tc.cuh per-step kernels, mlp_pass.cu pass kernel).
#pragma once
#include <unordered_map>

#include "common.cuh"
#include "tc_common.cuh"

namespace mlptc {

// ---- tensor-map cache ---------------------------------------------------------------------------------------
struct TmapKey {
    const void* p;
    uint64_t    rows, cols, ld, zstride;   // zstride = 0: plain 2-D map, else pair array [2][rows][cols]
    uint32_t    br, bc, swz;
    uint32_t    bp;                        // planes per box of a pair map (1 or 2)
    bool operator==(const TmapKey& o) const {
        return p == o.p && rows == o.rows && cols == o.cols && ld == o.ld && zstride == o.zstride && br == o.br && bc == o.bc && swz == o.swz && bp == o.bp;
    }
};
struct TmapHash {
    size_t operator()(const TmapKey& k) const {
        size_t h = (size_t)k.p;
        auto mix = [&](uint64_t v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
        mix(k.rows); mix(k.cols); mix(k.ld); mix(k.zstride); mix(k.br); mix(k.bc); mix(k.swz); mix(k.bp);
        return h;
    }
};
typedef std::unordered_map<TmapKey, CUtensorMap, TmapHash> TmapCache;

inline bool get_tmap(mamdr_ctx* ctx, CUtensorMap* out, const float* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t br,
                     uint32_t bc, CUtensorMapSwizzle swz, uint64_t zstride = 0, uint32_t box_planes = 1) {
    if (!ctx->tmap_cache) ctx->tmap_cache = new TmapCache();
    TmapCache& c = *static_cast<TmapCache*>(ctx->tmap_cache);
    TmapKey k{base, rows, cols, ld, zstride, br, bc, (uint32_t)swz, box_planes};
    auto it = c.find(k);
    if (it != c.end()) {
        *out = it->second;
        return true;
    }
    if (c.size() > 65536) c.clear();
    if (zstride ? !tc::make_tmap_pair_f32(out, base, rows, cols, ld, zstride, br, bc, swz, box_planes) : !tc::make_tmap_2d_f32(out, base, rows, cols, ld, br, bc, swz))
        return false;
    c.emplace(k, *out);
    return true;
}
inline void free_tmap_cache(mamdr_ctx* ctx) {
    delete static_cast<TmapCache*>(ctx->tmap_cache);
    ctx->tmap_cache = nullptr;
}

// K-major operand [rows, K]: box = [box_rows, 32];   MN-major operand [K, mn]: box = [32 k, 32 mn]
inline bool kmajor_map(mamdr_ctx* ctx, CUtensorMap* m, const float* p, uint64_t rows, uint64_t K, uint32_t box_rows) {
    return get_tmap(ctx, m, p, rows, K, K, box_rows, 32, CU_TENSOR_MAP_SWIZZLE_128B);
}
inline bool mnmajor_map(mamdr_ctx* ctx, CUtensorMap* m, const float* p, uint64_t K, uint64_t mn) {
    return get_tmap(ctx, m, p, K, mn, mn, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
}

// the same two views over a pair array [2][rows][cols] (hi plane, lo plane `zstride` floats later)
inline bool pair_kmajor_map(mamdr_ctx* ctx, CUtensorMap* m, const float* p, uint64_t rows, uint64_t K, uint32_t box_rows, uint64_t zstride,
                            uint32_t box_planes = 1) {
    return get_tmap(ctx, m, p, rows, K, K, box_rows, 32, CU_TENSOR_MAP_SWIZZLE_128B, zstride, box_planes);
}
inline bool pair_mnmajor_map(mamdr_ctx* ctx, CUtensorMap* m, const float* p, uint64_t K, uint64_t mn, uint64_t zstride) {
    return get_tmap(ctx, m, p, K, mn, mn, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, zstride);
}

}  // namespace mlptc
