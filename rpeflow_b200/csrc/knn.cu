// knn.cu — a4: exact k-nearest-neighbour search, D in {2,3}, k <= 32.
//
// Replaces models/csrc/k_nearest_neighbor/k_nearest_neighbor_kernel.cu:8-112 (one thread per query, insertion
// sort in local memory, every thread re-reading the same input point from global memory).
//
// Two kernels:
//   knn1_kernel        k == 1 (the 2-D "nearest projected point per pixel" calls of RPEFlow_core.py:329-330, which are
//                      60 % of all pairs): one thread per QPT queries, inputs broadcast from shared memory, running
//                      (best, index) in registers — no sort at all.
//   knn_select_kernel  2 <= k <= 32: one WARP per 4 queries.  The 32 lanes test 32 different inputs per step against
//                      the warp-uniform k-th best of each query; survivors are found with a ballot and inserted into
//                      a sorted list that lives one slot per lane (shuffle-up insertion).
//
// Both evaluate two (query, point) pairs per instruction with Blackwell's packed fp32x2 pipe (FADD2 / FFMA2):
// the input tile is stored in shared memory with every coordinate duplicated, so one LDS.64 yields the
// broadcast operand (x_j, x_j) for a pair of queries (q0.x, q1.x).
//
// Result order: (distance ascending, index ascending); distance = ((dx*dx+dy*dy)+dz*dz) with every
// operation rounded separately.  ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (observed with CUDA 12.9,
// even under --fmad=false), which would break that rule, so each square is computed as fma(d, d, -0.0) with the
// -0.0 passed as a kernel argument: one rounding of the product, and nothing left to contract.  Inputs are
// visited in increasing index order and a candidate only enters on strict '<', which gives the index tie rule.
#include <math_constants.h>

#include "common.cuh"

namespace b200 {

constexpr int KNN_TILE = 1024;   // inputs staged per shared-memory tile
constexpr int KNN_WARPS = 8;
constexpr int KNN_QPW = 4;       // queries per warp = 2 packed pairs

// one insertion round for one query of the warp: candidates = lanes whose distance beats the current k-th best
__device__ __forceinline__ void knn_insert(float d, int cand_base, float& ld, int& li, float& thr, int k, int lane) {
    unsigned m = __ballot_sync(FULL, d < thr);
    while (m) {                              // rare once the list is warm: ~k*ln(M/k) times per query
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const float cd = __shfl_sync(FULL, d, src);
        if (cd < thr) {                      // re-test: the threshold may have dropped in this very step
            const int pos = __popc(__ballot_sync(FULL, ld <= cd));       // equal distance: older (lower) index first
            const float up_d = __shfl_up_sync(FULL, ld, 1);
            const int up_i = __shfl_up_sync(FULL, li, 1);
            if (lane > pos) { ld = up_d; li = up_i; }
            if (lane == pos) { ld = cd; li = cand_base + src; }
            thr = __shfl_sync(FULL, ld, k - 1);
        }
    }
}

template <int D>
__global__ void __launch_bounds__(KNN_WARPS * 32)
knn_select_kernel(const float* __restrict__ input, const float* __restrict__ query, int64_t* __restrict__ out,
                  int M, int Q, int k, float negzero) {
    __shared__ float2 sx[KNN_TILE], sy[KNN_TILE], sz[D == 3 ? KNN_TILE : 1];      // (v,v) duplicated coordinates
    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = (blockIdx.x * KNN_WARPS + warp) * KNN_QPW;
    input += (size_t)b * M * D;
    query += (size_t)b * Q * D;
    const u64 nz = pack2(negzero, negzero);

    float qc[KNN_QPW][3];
    float thr[KNN_QPW], ld[KNN_QPW];
    int li[KNN_QPW];
#pragma unroll
    for (int i = 0; i < KNN_QPW; ++i) {
        const int q = min(q0 + i, Q - 1);            // out-of-range slots recompute the last query, never store
        qc[i][0] = __ldg(query + (size_t)q * D);
        qc[i][1] = __ldg(query + (size_t)q * D + 1);
        qc[i][2] = D == 3 ? __ldg(query + (size_t)q * D + 2) : 0.0f;
        thr[i] = CUDART_INF_F;                       // k-th best so far (warp-uniform)
        ld[i] = CUDART_INF_F;                        // lane s holds slot s of the sorted list
        li[i] = 0;                                   // k_nearest_neighbor.cpp:16 zero-initialises the indices
    }
    const u64 qxA = pack2(qc[0][0], qc[1][0]), qyA = pack2(qc[0][1], qc[1][1]), qzA = pack2(qc[0][2], qc[1][2]);
    const u64 qxB = pack2(qc[2][0], qc[3][0]), qyB = pack2(qc[2][1], qc[3][1]), qzB = pack2(qc[2][2], qc[3][2]);

    for (int base = 0; base < M; base += KNN_TILE) {
        const int cnt = min(KNN_TILE, M - base);
        const int cnt32 = (cnt + 31) & ~31;
        __syncthreads();
        for (int e = threadIdx.x; e < cnt32 * D; e += KNN_WARPS * 32) {   // coalesced AoS read -> duplicated SoA tile
            const int p = e / D, c = e - p * D;
            const float v = p < cnt ? __ldg(input + (size_t)base * D + e) : CUDART_INF_F;   // +inf padding never qualifies
            const float2 vv = make_float2(v, v);
            if (c == 0) sx[p] = vv;
            else if (c == 1) sy[p] = vv;
            else sz[p] = vv;
        }
        __syncthreads();

        const u64* px = reinterpret_cast<const u64*>(sx) + lane;
        const u64* py = reinterpret_cast<const u64*>(sy) + lane;
        const u64* pz = reinterpret_cast<const u64*>(sz) + (D == 3 ? lane : 0);
#pragma unroll 2
        for (int j0 = 0; j0 < cnt32; j0 += 32) {
            const u64 x = px[j0], y = py[j0], z = D == 3 ? pz[j0] : 0ull;
            float d0, d1, d2, d3;
            unpack2(sqdist_pair<D>(qxA, qyA, qzA, x, y, z, nz), d0, d1);
            unpack2(sqdist_pair<D>(qxB, qyB, qzB, x, y, z, nz), d2, d3);
            const bool hit = (d0 < thr[0]) | (d1 < thr[1]) | (d2 < thr[2]) | (d3 < thr[3]);
            if (__any_sync(FULL, hit)) {
                const int cb = base + j0;
                knn_insert(d0, cb, ld[0], li[0], thr[0], k, lane);
                knn_insert(d1, cb, ld[1], li[1], thr[1], k, lane);
                knn_insert(d2, cb, ld[2], li[2], thr[2], k, lane);
                knn_insert(d3, cb, ld[3], li[3], thr[3], k, lane);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < KNN_QPW; ++i)
        if (q0 + i < Q && lane < k) out[((size_t)b * Q + q0 + i) * k + lane] = li[i];
}

constexpr int KNN1_THREADS = 256;
constexpr int KNN1_TILE = 2048;

template <int D, int QPT>     // QPT = 2 or 4 queries per thread (1 or 2 packed pairs)
__global__ void __launch_bounds__(KNN1_THREADS)
knn1_kernel(const float* __restrict__ input, const float* __restrict__ query, int64_t* __restrict__ out, int M, int Q,
            float negzero) {
    constexpr int NP = QPT / 2;
    __shared__ float4 sxy[KNN1_TILE];                      // (x,x,y,y)
    __shared__ float2 szz[D == 3 ? KNN1_TILE : 1];         // (z,z)
    const int b = blockIdx.y;
    input += (size_t)b * M * D;
    query += (size_t)b * Q * D;
    const int qbase = blockIdx.x * KNN1_THREADS * QPT + threadIdx.x;
    const u64 nz = pack2(negzero, negzero);

    u64 qx[NP], qy[NP], qz[NP];
    float best[QPT];
    int bi[QPT];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        float c[2][3];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int q = min(qbase + (2 * p + h) * KNN1_THREADS, Q - 1);
            c[h][0] = __ldg(query + (size_t)q * D);
            c[h][1] = __ldg(query + (size_t)q * D + 1);
            c[h][2] = D == 3 ? __ldg(query + (size_t)q * D + 2) : 0.0f;
            best[2 * p + h] = CUDART_INF_F;
            bi[2 * p + h] = 0;
        }
        qx[p] = pack2(c[0][0], c[1][0]);
        qy[p] = pack2(c[0][1], c[1][1]);
        qz[p] = pack2(c[0][2], c[1][2]);
    }
    for (int base = 0; base < M; base += KNN1_TILE) {
        const int cnt = min(KNN1_TILE, M - base);
        __syncthreads();
        for (int p = threadIdx.x; p < cnt; p += KNN1_THREADS) {
            const float* s = input + (size_t)(base + p) * D;      // scalar loads: no alignment demand on the caller
            const float x = __ldg(s), y = __ldg(s + 1);
            sxy[p] = make_float4(x, x, y, y);
            if (D == 3) { const float z = __ldg(s + 2); szz[p] = make_float2(z, z); }
        }
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < cnt; ++j) {                  // every lane reads the same point: broadcast LDS
            const float4 pxy = sxy[j];
            const u64 px = pack2(pxy.x, pxy.y), py = pack2(pxy.z, pxy.w);
            u64 pz = 0ull;
            if (D == 3) pz = *reinterpret_cast<const u64*>(&szz[j]);
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                float d0, d1;
                unpack2(sqdist_pair<D>(qx[p], qy[p], qz[p], px, py, pz, nz), d0, d1);
                if (d0 < best[2 * p]) { best[2 * p] = d0; bi[2 * p] = base + j; }            // strict '<': lowest index wins ties
                if (d1 < best[2 * p + 1]) { best[2 * p + 1] = d1; bi[2 * p + 1] = base + j; }
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
    const float negzero = -0.0f;
    if (k == 1) {
        const int64_t total_q = (int64_t)B * Q;
        if (total_q >= (int64_t)sm_count() * 8 * KNN1_THREADS * 4) {
            dim3 grid(ceil_div(Q, KNN1_THREADS * 4), B);
            knn1_kernel<D, 4><<<grid, KNN1_THREADS, 0, st>>>(input, query, idx, M, Q, negzero);
        } else {
            dim3 grid(ceil_div(Q, KNN1_THREADS * 2), B);
            knn1_kernel<D, 2><<<grid, KNN1_THREADS, 0, st>>>(input, query, idx, M, Q, negzero);
        }
        return;
    }
    dim3 grid(ceil_div(Q, KNN_WARPS * KNN_QPW), B);
    knn_select_kernel<D><<<grid, KNN_WARPS * 32, 0, st>>>(input, query, idx, M, Q, k, negzero);
}

}  // namespace b200

extern "C" int b200_knn(const float* input_xyz, const float* query_xyz, int64_t* idx,
                        int B, int M, int Q, int D, int k, b200_stream_t stream) {
    using namespace b200;
    B200_REQUIRE((B == 0 || Q == 0) || (input_xyz && query_xyz && idx), "b200_knn: null pointer");   // empty calls carry null pointers
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
