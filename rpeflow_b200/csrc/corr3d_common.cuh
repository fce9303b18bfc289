// corr3d_common.cuh — pieces shared by the fp32 (corr3d.cu) and tensor-core (corr3d_tc.cu) paths of a5.
#pragma once
#include "common.cuh"

namespace b200 {

static inline int64_t align4(int64_t v) { return (v + 3) & ~int64_t(3); }

struct Corr3dScratch {
    float *A1, *G2, *P, *W2T, *W1cT, *n1WcT, *n2WcT;
    int64_t total;
};
static inline Corr3dScratch carve(float* base, int B, int Cout, int N1, int N2) {
    Corr3dScratch s;
    int64_t off = 0;
    auto take = [&](int64_t n) { float* p = base ? base + off : nullptr; off += align4(n); return p; };
    s.A1 = take((int64_t)B * N1 * Cout);
    s.G2 = take((int64_t)B * N2 * Cout);
    s.P = take((int64_t)B * N1 * Cout);
    s.W2T = take((int64_t)Cout * Cout);
    s.W1cT = take(3ll * Cout);
    s.n1WcT = take(8ll * Cout);
    s.n2WcT = take(8ll * Cout);
    s.total = off;
    return s;
}

// ---- the PointConv-style weight net: hidden 8-vector of relu(Wb.relu(Wa.d+ba)+bb) -------------------------------
__device__ __forceinline__ void weight_net_hidden(const float* __restrict__ Wa, const float* __restrict__ ba,
                                                  const float* __restrict__ Wb, const float* __restrict__ bb,
                                                  float dx, float dy, float dz, float* hid) {
    float h1[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) {
        float s = __ldg(ba + o);
        s = fmaf(__ldg(Wa + o * 3 + 0), dx, s);
        s = fmaf(__ldg(Wa + o * 3 + 1), dy, s);
        s = fmaf(__ldg(Wa + o * 3 + 2), dz, s);
        h1[o] = fmaxf(s, 0.0f);
    }
#pragma unroll
    for (int o = 0; o < 8; ++o) {
        float s = __ldg(bb + o);
#pragma unroll
        for (int i = 0; i < 8; ++i) s = fmaf(__ldg(Wb + o * 8 + i), h1[i], s);
        hid[o] = fmaxf(s, 0.0f);
    }
}


// corr3d_tc.cu: pass 2 on tcgen05 tensor cores (precision 1 = TF32, 2 = 3xTF32); false -> caller uses the fp32 kernel
bool corr3d_stage1_tc_eligible(int Cout, int k, int precision);
cudaError_t corr3d_stage1_tc(const float* xyz1, const float* xyz2, const int64_t* knn12, const Corr3dScratch& s,
                             const b200_corr3d_weights* w, int B, int Cout, int N1, int N2, int k, int precision,
                             cudaStream_t st);

}  // namespace b200
