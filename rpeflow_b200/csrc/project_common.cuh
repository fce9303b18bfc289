// project_common.cuh — pieces of a8 (project_feat_with_nn_corr) shared by the two-pass kernels of gather.cu and the
// tiled route of project_tile.cu.
#pragma once
#include "common.cuh"

namespace b200 {

__host__ __device__ __forceinline__ int round4(int c) { return (c + 3) & ~3; }

constexpr int PN_SLAB = 32;

// out[b, 3 + k, p] = T[b, nn[b,p], k] for one slab of 32 feat3d channels and one pixel: 128-bit loads from the nearest
// point's row (16-byte aligned: row strides are multiples of 4 floats), coalesced streaming stores over the pixels of a warp.
__device__ __forceinline__ void project_slab_copy(const float* __restrict__ g, float* __restrict__ o, int HW, int k0, int k1) {
    int k = k0;
    for (; k + 16 <= k1; k += 16) {                                     // 4 row loads in flight, 16 plane stores
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = __ldg(reinterpret_cast<const float4*>(g + k) + u);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            __stcs(o + (size_t)(3 + k + 4 * u + 0) * HW, v[u].x);
            __stcs(o + (size_t)(3 + k + 4 * u + 1) * HW, v[u].y);
            __stcs(o + (size_t)(3 + k + 4 * u + 2) * HW, v[u].z);
            __stcs(o + (size_t)(3 + k + 4 * u + 3) * HW, v[u].w);
        }
    }
    for (; k + 4 <= k1; k += 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(g + k));
        __stcs(o + (size_t)(3 + k + 0) * HW, v.x);
        __stcs(o + (size_t)(3 + k + 1) * HW, v.y);
        __stcs(o + (size_t)(3 + k + 2) * HW, v.z);
        __stcs(o + (size_t)(3 + k + 3) * HW, v.w);
    }
    for (; k < k1; ++k) __stcs(o + (size_t)(3 + k) * HW, __ldg(g + k));
}

}  // namespace b200
