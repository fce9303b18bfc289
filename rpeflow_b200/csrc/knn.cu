// knn.cu — a4: exact k-nearest-neighbour search, D in {2,3}, k <= 32.
//
// Replaces models/csrc/k_nearest_neighbor/k_nearest_neighbor_kernel.cu:8-112 (one thread per query, insertion
// sort in local memory, every thread re-reading the same input point from global memory).
//
// Two kernels:
//   knn1_kernel        k == 1 (the 2-D "nearest projected point per pixel" calls of RPEFlow_core.py:329-330, which are
//                      60 % of all pairs): one thread per query, inputs broadcast from shared memory, running
//                      (best, index) in registers — no sort at all.
//   knn_select_kernel  2 <= k <= 32: one WARP per query (QPW queries register-blocked per warp).  The 32 lanes
//                      test 32 different inputs per step against the warp-uniform k-th best; survivors are found
//                      with a ballot and inserted into a sorted list that lives one slot per lane (shuffle-up
//                      insertion).  After the list warms up almost every step is 8 FP ops + 1 vote per lane.
//
// Result order: (distance ascending, index ascending); distance = ((dx*dx+dy*dy)+dz*dz) with
// every operation rounded separately (sqdist*_rule).  Inputs are visited in increasing index order and a
// candidate only enters on strict '<', which is exactly that order.
#include <math_constants.h>

#include "common.cuh"

namespace b200 {

constexpr int KNN_TILE = 1024;   // inputs staged per shared-memory tile (SoA, 12 KB)
constexpr int KNN_WARPS = 8;

template <int D, int QPW>
__global__ void __launch_bounds__(KNN_WARPS * 32)
knn_select_kernel(const float* __restrict__ input, const float* __restrict__ query, int64_t* __restrict__ out,
                  int M, int Q, int k) {
    __shared__ float sx[KNN_TILE], sy[KNN_TILE], sz[D == 3 ? KNN_TILE : 1];
    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = (blockIdx.x * KNN_WARPS + warp) * QPW;
    input += (size_t)b * M * D;
    query += (size_t)b * Q * D;

    float qx[QPW], qy[QPW], qz[QPW], thr[QPW], ld[QPW];
    int li[QPW];
#pragma unroll
    for (int i = 0; i < QPW; ++i) {
        const int q = min(q0 + i, Q - 1);            // out-of-range warps recompute the last query, never store
        qx[i] = __ldg(query + (size_t)q * D);
        qy[i] = __ldg(query + (size_t)q * D + 1);
        qz[i] = D == 3 ? __ldg(query + (size_t)q * D + 2) : 0.0f;
        thr[i] = CUDART_INF_F;                       // k-th best so far (warp-uniform)
        ld[i] = CUDART_INF_F;                        // lane s holds slot s of the sorted list
        li[i] = 0;                                   // k_nearest_neighbor.cpp:16 zero-initialises the indices
    }

    for (int base = 0; base < M; base += KNN_TILE) {
        const int cnt = min(KNN_TILE, M - base);
        __syncthreads();
        for (int e = threadIdx.x; e < cnt * D; e += KNN_WARPS * 32) {   // coalesced AoS read -> SoA tile
            const float v = __ldg(input + (size_t)base * D + e);
            const int p = e / D, c = e - p * D;
            if (c == 0) sx[p] = v;
            else if (c == 1) sy[p] = v;
            else sz[p] = v;
        }
        __syncthreads();

        for (int j0 = 0; j0 < cnt; j0 += 32) {
            const int j = j0 + lane;
            const bool ok = j < cnt;
            const float px = sx[ok ? j : 0], py = sy[ok ? j : 0], pz = D == 3 ? sz[ok ? j : 0] : 0.0f;
#pragma unroll
            for (int i = 0; i < QPW; ++i) {
                float d = D == 3 ? sqdist3_rule(qx[i], qy[i], qz[i], px, py, pz) : sqdist2_rule(qx[i], qy[i], px, py);
                if (!ok) d = CUDART_INF_F;
                unsigned m = __ballot_sync(FULL, d < thr[i]);
                while (m) {                          // rare once the list is warm: ~k*ln(M/k) times per query
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const float cd = __shfl_sync(FULL, d, src);
                    if (cd < thr[i]) {               // re-test: the threshold may have dropped in this very step
                        const int ci = base + j0 + src;
                        const int pos = __popc(__ballot_sync(FULL, ld[i] <= cd));   // equal distance: older (lower) index first
                        const float up_d = __shfl_up_sync(FULL, ld[i], 1);
                        const int up_i = __shfl_up_sync(FULL, li[i], 1);
                        if (lane > pos) { ld[i] = up_d; li[i] = up_i; }
                        if (lane == pos) { ld[i] = cd; li[i] = ci; }
                        thr[i] = __shfl_sync(FULL, ld[i], k - 1);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < QPW; ++i)
        if (q0 + i < Q && lane < k) out[((size_t)b * Q + q0 + i) * k + lane] = li[i];
}

constexpr int KNN1_THREADS = 256;
constexpr int KNN1_TILE = 2048;

template <int D, int QPT>
__global__ void __launch_bounds__(KNN1_THREADS)
knn1_kernel(const float* __restrict__ input, const float* __restrict__ query, int64_t* __restrict__ out, int M, int Q) {
    __shared__ float4 sp4[D == 3 ? KNN1_TILE : 1];
    __shared__ float2 sp2[D == 2 ? KNN1_TILE : 1];
    const int b = blockIdx.y;
    input += (size_t)b * M * D;
    query += (size_t)b * Q * D;
    const int qbase = blockIdx.x * KNN1_THREADS * QPT + threadIdx.x;

    float qx[QPT], qy[QPT], qz[QPT], best[QPT];
    int bi[QPT];
#pragma unroll
    for (int i = 0; i < QPT; ++i) {
        const int q = min(qbase + i * KNN1_THREADS, Q - 1);
        qx[i] = __ldg(query + (size_t)q * D);
        qy[i] = __ldg(query + (size_t)q * D + 1);
        qz[i] = D == 3 ? __ldg(query + (size_t)q * D + 2) : 0.0f;
        best[i] = CUDART_INF_F;
        bi[i] = 0;
    }
    for (int base = 0; base < M; base += KNN1_TILE) {
        const int cnt = min(KNN1_TILE, M - base);
        __syncthreads();
        for (int p = threadIdx.x; p < cnt; p += KNN1_THREADS) {
            if (D == 2) {
                const float* s = input + (size_t)(base + p) * 2;      // scalar loads: no alignment demand on the caller
                sp2[p] = make_float2(__ldg(s), __ldg(s + 1));
            } else {
                const float* s = input + (size_t)(base + p) * 3;
                sp4[p] = make_float4(__ldg(s), __ldg(s + 1), __ldg(s + 2), 0.0f);
            }
        }
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < cnt; ++j) {              // every lane reads the same point: one broadcast LDS
            float px, py, pz = 0.0f;
            if (D == 2) { const float2 p = sp2[j]; px = p.x; py = p.y; }
            else        { const float4 p = sp4[j]; px = p.x; py = p.y; pz = p.z; }
#pragma unroll
            for (int i = 0; i < QPT; ++i) {
                const float d = D == 3 ? sqdist3_rule(qx[i], qy[i], qz[i], px, py, pz) : sqdist2_rule(qx[i], qy[i], px, py);
                if (d < best[i]) { best[i] = d; bi[i] = base + j; }      // strict '<': lowest index wins ties
            }
        }
    }
#pragma unroll
    for (int i = 0; i < QPT; ++i) {
        const int q = qbase + i * KNN1_THREADS;
        if (q < Q) out[(size_t)b * Q + q] = bi[i];
    }
}

template <int D>
static void launch_knn(const float* input, const float* query, int64_t* idx, int B, int M, int Q, int k, cudaStream_t st) {
    const int64_t total_q = (int64_t)B * Q;
    const int64_t fill = (int64_t)sm_count() * 8;    // CTAs wanted before register-blocking queries
    if (k == 1) {
        if (total_q >= fill * KNN1_THREADS * 2) {
            dim3 grid(ceil_div(Q, KNN1_THREADS * 2), B);
            knn1_kernel<D, 2><<<grid, KNN1_THREADS, 0, st>>>(input, query, idx, M, Q);
        } else {
            dim3 grid(ceil_div(Q, KNN1_THREADS), B);
            knn1_kernel<D, 1><<<grid, KNN1_THREADS, 0, st>>>(input, query, idx, M, Q);
        }
        return;
    }
    if (total_q >= fill * KNN_WARPS * 4) {
        dim3 grid(ceil_div(Q, KNN_WARPS * 4), B);
        knn_select_kernel<D, 4><<<grid, KNN_WARPS * 32, 0, st>>>(input, query, idx, M, Q, k);
    } else if (total_q >= fill * KNN_WARPS * 2) {
        dim3 grid(ceil_div(Q, KNN_WARPS * 2), B);
        knn_select_kernel<D, 2><<<grid, KNN_WARPS * 32, 0, st>>>(input, query, idx, M, Q, k);
    } else {
        dim3 grid(ceil_div(Q, KNN_WARPS), B);
        knn_select_kernel<D, 1><<<grid, KNN_WARPS * 32, 0, st>>>(input, query, idx, M, Q, k);
    }
}

}  // namespace b200

extern "C" int b200_knn(const float* input_xyz, const float* query_xyz, int64_t* idx,
                        int B, int M, int Q, int D, int k, b200_stream_t stream) {
    using namespace b200;
    B200_REQUIRE(input_xyz && query_xyz && idx, "b200_knn: null pointer");
    B200_REQUIRE(D == 2 || D == 3, "b200_knn: D must be 2 or 3 (got %d)", D);
    B200_REQUIRE(k >= 1 && k <= 32, "b200_knn: k must be in [1,32] (got %d); the reference kernel has 32 slots", k);
    B200_REQUIRE(B >= 0 && M >= 1 && Q >= 0, "b200_knn: bad sizes B=%d M=%d Q=%d", B, M, Q);
    B200_REQUIRE(B <= 65535, "b200_knn: B=%d exceeds gridDim.y", B);
    if (B == 0 || Q == 0) return B200_OK;
    if (D == 2) launch_knn<2>(input_xyz, query_xyz, idx, B, M, Q, k, as_stream(stream));
    else        launch_knn<3>(input_xyz, query_xyz, idx, B, M, Q, k, as_stream(stream));
    B200_LAUNCH_CHECK("b200_knn");
    return B200_OK;
}
