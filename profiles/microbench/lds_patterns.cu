// Micro-benchmark: shared-memory LDS.128 cost vs address pattern (do lanes that read the same 16 bytes share a
// wavefront?), and FFMA2 : LDS.128 co-issue at several ratios.  sm_100a (B200).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o lds_patterns lds_patterns.cu && ./lds_patterns
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define ITERS 2048
#define NLD 8

__device__ __forceinline__ int pattern_addr(int pat, int lane) {   // in units of 16 bytes
    switch (pat) {
        case 0: return lane;                                   // 32 distinct, 512 B contiguous
        case 1: return lane & 7;                               // 8 distinct (every quarter-warp reads all 8)
        case 2: return lane >> 2;                              // 8 distinct (each quarter-warp reads 2)
        case 3: return 0;                                      // full broadcast
        case 4: return lane & 15;                              // 16 distinct
        case 5: return (lane & 7) * 17;                        // 8 distinct rows, stride 17*16 B (bank rotation)
        case 6: return ((lane & 7) + (lane >> 3)) * 17;        // 11 distinct rows (y + dy)
        case 7: return (lane & 7) * 17 + (lane >> 3) * 4;      // 32 distinct, 8 rows x 4 strips of 64 B
        case 8: return (lane & 3) * 17 + (lane >> 2) * 0;      // 4 distinct rows
        default: return lane * 17;                             // 32 distinct rows, rotating banks
    }
}

template <int RATIO>   // FFMA2 per LDS.128 (0 = loads only)
__global__ void __launch_bounds__(256) bench(float* out, int pat, float a) {
    extern __shared__ float4 sm[];
    for (int i = threadIdx.x; i < 3072; i += blockDim.x) sm[i] = make_float4(i, 1.f, 2.f, 3.f);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    unsigned base = (unsigned)__cvta_generic_to_shared(sm) + pattern_addr(pat, lane) * 16;
    u64 acc[8];
    for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1,%2};" : "=l"(acc[i]) : "f"(a * i), "f"(a));
    float s = 0.f;
    for (int it = 0; it < ITERS; ++it) {
        const unsigned rot = (unsigned)(it & 7) * 5120u;               // loop-variant address: ptxas must not hoist the loads
#pragma unroll
        for (int j = 0; j < NLD; ++j) {
            float x, y, z, w;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(base + rot + j * 16 * 40) : "memory");
            {
                u64 lo, hi;
                asm("mov.b64 %0, {%1,%2};" : "=l"(lo) : "f"(x), "f"(y));
                asm("mov.b64 %0, {%1,%2};" : "=l"(hi) : "f"(z), "f"(w));
#pragma unroll
                for (int r = 0; r < RATIO; ++r)
                    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[r & 7]) : "l"(r & 1 ? hi : lo), "l"(lo));
            }
        }
    }
    for (int i = 0; i < 8; ++i) { float lo, hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i])); s += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int RATIO> void run(int pat, const char* name) {
    float* out; cudaMalloc(&out, 148 * 4 * 256 * 4);
    cudaFuncSetAttribute(bench<RATIO>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3072 * 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench<RATIO><<<148 * 4, 256, 3072 * 16>>>(out, pat, 1.0001f);
    cudaEventRecord(e0); bench<RATIO><<<148 * 4, 256, 3072 * 16>>>(out, pat, 1.0001f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double lds = 4.0 * 8 * ITERS * NLD;                              // warp-level LDS.128 per SM
    double cyc = ms * 1e-3 * clk * 1e3;
    printf("ratio %d  pattern %-34s %7.3f ms  %5.2f clk/LDS.128/SM   %5.2f FFMA2/clk/SM\n", RATIO, name, ms, cyc / lds,
           lds * RATIO / cyc);
    cudaFree(out);
}
int main() {
    const char* names[] = {"32 distinct contiguous", "8 distinct (lane&7)", "8 distinct (lane>>2)", "broadcast", "16 distinct",
                           "8 rows stride 17", "11 rows (y+dy) stride 17", "8 rows x 4 strips (32 distinct)", "4 rows", "32 rows stride 17"};
    for (int p : {0, 3, 5, 6, 7, 9}) { run<1>(p, names[p]); run<2>(p, names[p]); run<3>(p, names[p]); run<4>(p, names[p]); run<6>(p, names[p]); }
    return 0;
}
