// fps.cu — a3: furthest point sampling.
//
// Replaces models/csrc/furthest_point_sampling/furthest_point_sampling_kernel.cu:34-85 (one 1024-thread CTA
// per cloud; every iteration re-loads all points from global memory, read-modify-writes a global distance
// array, then does a 10-level shared-memory tree argmax with five __syncthreads).
//
// Here the points AND their running distances live in registers for the whole kernel, so an iteration is
// pure ALU work followed by a two-level warp argmax:
//   per thread : PPT points: d = ((dx*dx+dy*dy)+dz*dz) (non-fused), dist = min(dist, d), running (max, index)
//   per warp   : redux.sync.max on the distance bits, then redux.sync.min on the indices that hold that max
//   per CTA    : lane 0 of each warp -> shared slot; ONE __syncthreads; every warp repeats the two redux on
//                the <=32 slots, so every thread knows the winner without a second barrier (slots are
//                double-buffered by iteration parity).
// Clouds larger than 8192 points (config 5: 32768) are split over a thread-block cluster of CS CTAs; each CTA
// publishes {key, x, y, z} of its local winner into every peer's shared memory (DSMEM) and one cluster
// barrier per iteration replaces the global round trip.
//
// Exactness: start at index 0, distances start at 1e10, ties -> LOWEST index.  Equal to the
// reference torch fallback models/csrc/wrapper.py:83-96 bit for bit.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace b200 {

struct FpsSlot {               // what a CTA publishes to its cluster peers each iteration
    unsigned long long key;    // (distance bits << 32) | (0xFFFFFFFF - index): max key = max distance, then min index
    float x, y, z, pad;
};

template <int PPT, int T, int CS>
__global__ void __launch_bounds__(T, 1)
fps_kernel(const float* __restrict__ xyz, int64_t* __restrict__ out, int N, int S, float negzero) {
    constexpr int NW = T / 32;
    __shared__ unsigned s_d[2][32], s_i[2][32];
    __shared__ __align__(16) FpsSlot s_slot[2][CS > 1 ? CS : 1];
    extern __shared__ float s_pts[];                     // CS > 1: xyz of this CTA's chunk

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int cloud = blockIdx.x, rank = 0;
    if (CS > 1) {
        rank = (int)cg::this_cluster().block_rank();
        cloud = blockIdx.x / CS;
    }
    const int chunk = (N + CS - 1) / CS;
    const int lo = rank * chunk, hi = min(N, lo + chunk);
    const float* P = xyz + (size_t)cloud * N * 3;

    float px[PPT], py[PPT], pz[PPT], dist[PPT];
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
        const int gi = lo + i * T + tid;
        const bool ok = gi < hi;
        px[i] = ok ? __ldg(P + (size_t)gi * 3 + 0) : 0.0f;
        py[i] = ok ? __ldg(P + (size_t)gi * 3 + 1) : 0.0f;
        pz[i] = ok ? __ldg(P + (size_t)gi * 3 + 2) : 0.0f;
        dist[i] = ok ? 1e10f : 0.0f;                     // padding never beats a real point (index tie-break below)
        if (CS > 1 && ok) {
            s_pts[(i * T + tid) * 3 + 0] = px[i];
            s_pts[(i * T + tid) * 3 + 1] = py[i];
            s_pts[(i * T + tid) * 3 + 2] = pz[i];
        }
    }
    if (CS > 1) cg::this_cluster().sync();

    unsigned cur = 0;
    float cx = __ldg(P + 0), cy = __ldg(P + 1), cz = __ldg(P + 2);
    int64_t* o = out + (size_t)cloud * S;

    for (int s = 0; s < S; ++s) {
        if (rank == 0 && tid == 0) o[s] = (int64_t)cur;
        if (s == S - 1) break;

        // Running argmax over this thread's points.  Only the slot number is tracked in the loop (SEL with an
        // immediate); padding slots hold distance 0 and an index >= hi, so they lose every tie to a real point.
        float bd = -1.0f;
        int bslot = 0;
        if (PPT >= 2) {                                  // two points per FADD2/FFMA2 (non-fused rule kept, see sqr2)
            const u64 nz = pack2(negzero, negzero);
            const u64 c2x = pack2(cx, cx), c2y = pack2(cy, cy), c2z = pack2(cz, cz);
#pragma unroll
            for (int i = 0; i + 1 < PPT; i += 2) {
                float d0, d1;
                unpack2(sqdist_pair<3>(pack2(px[i], px[i + 1]), pack2(py[i], py[i + 1]), pack2(pz[i], pz[i + 1]),
                                       c2x, c2y, c2z, nz), d0, d1);
                dist[i] = fminf(dist[i], d0);
                dist[i + 1] = fminf(dist[i + 1], d1);
                if (dist[i] > bd) { bd = dist[i]; bslot = i; }                  // increasing index, strict '>': lowest index wins
                if (dist[i + 1] > bd) { bd = dist[i + 1]; bslot = i + 1; }
            }
        } else {
#pragma unroll
            for (int i = 0; i < PPT; ++i) {
                const float d = sqdist3_rule(px[i], py[i], pz[i], cx, cy, cz);
                dist[i] = fminf(dist[i], d);
                if (dist[i] > bd) { bd = dist[i]; bslot = i; }
            }
        }
        const unsigned bi = (unsigned)(lo + bslot * T + tid);
        const unsigned db = __float_as_uint(bd);         // distances are >= +0: the bit pattern orders like the value
        const unsigned wmax = __reduce_max_sync(FULL, db);
        const unsigned widx = __reduce_min_sync(FULL, db == wmax ? bi : 0xFFFFFFFFu);
        const int par = s & 1;
        if (lane == 0) { s_d[par][warp] = wmax; s_i[par][warp] = widx; }
        __syncthreads();

        if (CS == 1) {
            const unsigned vd = lane < NW ? s_d[par][lane] : 0u;
            const unsigned vi = lane < NW ? s_i[par][lane] : 0xFFFFFFFFu;
            const unsigned gmax = __reduce_max_sync(FULL, vd);
            cur = __reduce_min_sync(FULL, vd == gmax ? vi : 0xFFFFFFFFu);
            cx = __ldg(P + (size_t)cur * 3 + 0);         // L1-resident after the first touch
            cy = __ldg(P + (size_t)cur * 3 + 1);
            cz = __ldg(P + (size_t)cur * 3 + 2);
        } else {
            cg::cluster_group cluster = cg::this_cluster();
            if (warp == 0) {
                const unsigned vd = lane < NW ? s_d[par][lane] : 0u;
                const unsigned vi = lane < NW ? s_i[par][lane] : 0xFFFFFFFFu;
                const unsigned gmax = __reduce_max_sync(FULL, vd);
                const unsigned gidx = __reduce_min_sync(FULL, vd == gmax ? vi : 0xFFFFFFFFu);
                if (lane < CS) {                         // lane r writes this CTA's winner into peer r's slot
                    FpsSlot v;
                    v.key = ((unsigned long long)gmax << 32) | (unsigned long long)(0xFFFFFFFFu - gidx);
                    const int li = (int)gidx - lo;          // < PPT*T even for a padding slot (never the global winner)
                    v.x = s_pts[li * 3 + 0]; v.y = s_pts[li * 3 + 1]; v.z = s_pts[li * 3 + 2]; v.pad = 0.0f;
                    FpsSlot* dst = cluster.map_shared_rank(&s_slot[par][rank], lane);
                    *dst = v;
                }
            }
            cluster.sync();                              // release our DSMEM stores / acquire the peers'
            FpsSlot w = s_slot[par][0];
#pragma unroll
            for (int r = 1; r < CS; ++r) {
                const FpsSlot c = s_slot[par][r];
                if (c.key > w.key) w = c;
            }
            cur = 0xFFFFFFFFu - (unsigned)(w.key & 0xFFFFFFFFull);
            cx = w.x; cy = w.y; cz = w.z;
        }
    }
    if (CS > 1) cg::this_cluster().sync();               // no CTA may exit while peers can still write its smem
}

template <int PPT, int T, int CS>
static cudaError_t launch_fps(const float* xyz, int64_t* idx, int B, int N, int S, cudaStream_t st) {
    auto kern = fps_kernel<PPT, T, CS>;
    if (CS == 1) {
        kern<<<B, T, 0, st>>>(xyz, idx, N, S, -0.0f);
        return cudaGetLastError();
    }
    const size_t smem = (size_t)PPT * T * 3 * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(B * CS);
    cfg.blockDim = dim3(T);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, xyz, idx, N, S, -0.0f);
}

template <int CS>
static cudaError_t dispatch_ppt(const float* xyz, int64_t* idx, int B, int N, int S, cudaStream_t st) {
    const int per_cta = (N + CS - 1) / CS;
    if (per_cta <= 512) return launch_fps<1, 512, CS>(xyz, idx, B, N, S, st);
    if (per_cta <= 1024) return launch_fps<2, 512, CS>(xyz, idx, B, N, S, st);
    if (per_cta <= 2048) return launch_fps<4, 512, CS>(xyz, idx, B, N, S, st);
    if (per_cta <= 4096) return launch_fps<8, 512, CS>(xyz, idx, B, N, S, st);
    return launch_fps<16, 512, CS>(xyz, idx, B, N, S, st);
}

}  // namespace b200

extern "C" int b200_fps(const float* xyz, int64_t* idx, int B, int N, int n_samples, b200_stream_t stream) {
    using namespace b200;
    B200_REQUIRE(xyz && idx, "b200_fps: null pointer");
    B200_REQUIRE(B >= 0 && n_samples >= 1, "b200_fps: bad sizes B=%d n_samples=%d", B, n_samples);
    B200_REQUIRE(N > n_samples, "b200_fps: need N > n_samples (N=%d, n_samples=%d) as models/csrc/wrapper.py:98", N, n_samples);
    if (N > 8 * 8192) {
        set_error("b200_fps: N=%d exceeds 65536 points per cloud (8-CTA cluster x 8192)", N);
        return B200_ENOSUP;
    }
    if (B == 0) return B200_OK;
    // Cluster size: the smallest that fits the cloud in registers; B200_FPS_CLUSTER=2|4|8 forces a larger one
    // (shorter scan per iteration, one cluster barrier more) for small batches.
    int cs = N <= 8192 ? 1 : N <= 16384 ? 2 : N <= 32768 ? 4 : 8;
    if (const char* env = getenv("B200_FPS_CLUSTER")) {
        const int want = atoi(env);
        if ((want == 1 || want == 2 || want == 4 || want == 8) && want > cs) cs = want;
    }
    cudaStream_t st = as_stream(stream);
    cudaError_t e;
    switch (cs) {
        case 1: e = dispatch_ppt<1>(xyz, idx, B, N, n_samples, st); break;
        case 2: e = dispatch_ppt<2>(xyz, idx, B, N, n_samples, st); break;
        case 4: e = dispatch_ppt<4>(xyz, idx, B, N, n_samples, st); break;
        default: e = dispatch_ppt<8>(xyz, idx, B, N, n_samples, st); break;
    }
    if (e != cudaSuccess) return cuda_fail(e, "b200_fps");
    return B200_OK;
}
