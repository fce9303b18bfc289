// corr2d_small.cu — a1 on the coarse pyramid levels: NCHW maps whose rows are NOT a multiple of 16 bytes (W = 30, 15 at
// 960x540: levels 4 and 5), which TMA cannot address (global strides must be 16-byte multiples), so the TMA kernels of
// corr2d_nchw.cu / corr2d_diag.cu do not take them.  Until r2 these maps went the reference's way — two torch permutes to
// NHWC plus the NHWC kernel (wrapper.py:68-70) — and ran at 5-7 % of the HBM roofline: 0.17 ms per step for 0.07 GB.
//
// The maps are tiny (<= 540 pixels), so the design is plain: CTA = (sample, band of rows); the band of in1 and the halo
// band of in2 are staged in shared memory 8 channels at a time with coalesced row loads (zero fill = the reference's
// padding); warp w owns row shift dy = w - md; a lane owns up to four PAIRS of horizontally adjacent pixels and all
// 2md+1 column shifts of both: 18 accumulators per pair, each in2 window value feeds two pixels.
#include <algorithm>

#include "common.cuh"

namespace b200 {

constexpr int CS_CC = 8;                 // channels per shared-memory stage
constexpr int CS_IPT = 4;                // pixel pairs per lane
constexpr int CS_MAXPAIRS = 32 * CS_IPT;

template <int MD>
__global__ void __launch_bounds__((2 * MD + 1) * 32)
corr2d_fwd_small_kernel(const float* __restrict__ in1, const float* __restrict__ in2, float* __restrict__ out, int C, int H, int W,
                        int TR, int bands, float inv_c, float slope) {
    constexpr int ND = 2 * MD + 1, THREADS = ND * 32;
    extern __shared__ float cs_smem[];
    const int b = blockIdx.x / bands, band = blockIdx.x % bands;
    const int y0 = band * TR, rows = min(TR, H - y0);
    const int PW = (W + 1) / 2;                          // pixel pairs per row
    const int W1 = 2 * PW + 1;                           // in1 row pitch (odd: pairs of lanes hit different banks)
    const int W2 = W + 2 * MD + 2;                       // in2 halo row pitch (+1 column read by the second pixel of the last pair)
    float* s1 = cs_smem;                                 // [CS_CC][TR][W1]
    float* s2 = cs_smem + CS_CC * TR * W1;               // [CS_CC][TR + 2 MD][W2]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int npairs = rows * PW;

    float acc[CS_IPT][2][ND];
#pragma unroll
    for (int i = 0; i < CS_IPT; ++i)
#pragma unroll
        for (int d = 0; d < ND; ++d) acc[i][0][d] = acc[i][1][d] = 0.0f;
    int py[CS_IPT], px[CS_IPT];
#pragma unroll
    for (int i = 0; i < CS_IPT; ++i) {
        const int p = min(lane + 32 * i, max(npairs - 1, 0));
        py[i] = p / PW;
        px[i] = 2 * (p - py[i] * PW);
    }

    const size_t plane = (size_t)H * W;
    for (int c0 = 0; c0 < C; c0 += CS_CC) {
        __syncthreads();
        for (int e = threadIdx.x; e < CS_CC * TR * W1; e += THREADS) {          // in1 band, zero beyond the map / channel count
            const int x = e % W1, r = (e / W1) % TR, c = e / (W1 * TR);
            const int gy = y0 + r;
            s1[e] = (c0 + c < C && gy < H && x < W) ? __ldg(in1 + ((size_t)b * C + c0 + c) * plane + (size_t)gy * W + x) : 0.0f;
        }
        for (int e = threadIdx.x; e < CS_CC * (TR + 2 * MD) * W2; e += THREADS) {   // in2 halo band
            const int x = e % W2, r = (e / W2) % (TR + 2 * MD), c = e / (W2 * (TR + 2 * MD));
            const int gy = y0 + r - MD, gx = x - MD;
            s2[e] = (c0 + c < C && gy >= 0 && gy < H && gx >= 0 && gx < W)
                        ? __ldg(in2 + ((size_t)b * C + c0 + c) * plane + (size_t)gy * W + gx) : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < CS_IPT; ++i) {
            if (lane + 32 * i < npairs) {
                const float* a = s1 + py[i] * W1 + px[i];
                const float* w = s2 + (py[i] + warp) * W2 + px[i];               // halo row y + dy + MD, dy = warp - MD
#pragma unroll
                for (int c = 0; c < CS_CC; ++c) {
                    const float a0 = a[c * TR * W1], a1 = a[c * TR * W1 + 1];
                    float v[ND + 1];
#pragma unroll
                    for (int d = 0; d <= ND; ++d) v[d] = w[c * (TR + 2 * MD) * W2 + d];
#pragma unroll
                    for (int d = 0; d < ND; ++d) {
                        acc[i][0][d] = fmaf(a0, v[d], acc[i][0][d]);
                        acc[i][1][d] = fmaf(a1, v[d + 1], acc[i][1][d]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < CS_IPT; ++i) {
        if (lane + 32 * i < npairs) {
            const int y = y0 + py[i], x = px[i];
            float* o = out + ((size_t)b * ND * ND + (size_t)warp * ND) * plane + (size_t)y * W + x;
#pragma unroll
            for (int d = 0; d < ND; ++d) {
                float r0 = acc[i][0][d] * inv_c, r1 = acc[i][1][d] * inv_c;
                r0 = r0 > 0.0f ? r0 : r0 * slope;                                // slope = 1: plain correlation
                r1 = r1 > 0.0f ? r1 : r1 * slope;
                o[(size_t)d * plane] = r0;
                if (x + 1 < W) o[(size_t)d * plane + 1] = r1;
            }
        }
    }
}

bool corr2d_small_eligible(int B, int C, int H, int W, int md) {
    (void)C;
    return md >= 1 && md <= 4 && (W + 1) / 2 <= CS_MAXPAIRS && B > 0 && H > 0 && (int64_t)B * H < (1 << 30);
}

cudaError_t corr2d_fwd_small(const float* in1, const float* in2, float* out, int B, int C, int H, int W, int md, float slope,
                             cudaStream_t st) {
    const int PW = (W + 1) / 2;
    int TR = std::max(1, std::min(H, CS_MAXPAIRS / PW));
    // more, shorter bands when the launch would not even give every SM one CTA
    while (TR > 2 && (int64_t)B * ceil_div(H, TR) < sm_count()) TR = (TR + 1) / 2;
    const int bands = ceil_div(H, TR);
    const int W1 = 2 * PW + 1, W2 = W + 2 * md + 2;
    const size_t smem = (size_t)CS_CC * (TR * W1 + (TR + 2 * md) * W2) * sizeof(float);
    const float inv_c = 1.0f / (float)C;
    cudaError_t e = cudaSuccess;
#define CS_LAUNCH(MD)                                                                                                        \
    do {                                                                                                                     \
        e = cudaFuncSetAttribute(corr2d_fwd_small_kernel<MD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);        \
        if (e != cudaSuccess) return e;                                                                                      \
        corr2d_fwd_small_kernel<MD><<<B * bands, (2 * MD + 1) * 32, smem, st>>>(in1, in2, out, C, H, W, TR, bands, inv_c, slope); \
    } while (0)
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    if (md == 4) CS_LAUNCH(4);
    else if (md == 3) CS_LAUNCH(3);
    else if (md == 2) CS_LAUNCH(2);
    else CS_LAUNCH(1);
#undef CS_LAUNCH
    return cudaGetLastError();
}

}  // namespace b200
