// Micro-benchmark: issue throughput of scalar vs packed fp32 instructions on sm_100a (B200).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp32_pipes fp32_pipes.cu && ./fp32_pipes
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define ITERS 4096
#define NACC 8
template <int OP>
__global__ void __launch_bounds__(256) bench(float* out, float a, float b) {
    float x[NACC]; u64 p[NACC];
    for (int i = 0; i < NACC; ++i) { x[i] = a * (threadIdx.x + i); asm("mov.b64 %0, {%1,%2};" : "=l"(p[i]) : "f"(x[i]), "f"(x[i] + 1.f)); }
    u64 pb; asm("mov.b64 %0, {%1,%2};" : "=l"(pb) : "f"(b), "f"(b));
    u64 pa; asm("mov.b64 %0, {%1,%2};" : "=l"(pa) : "f"(a), "f"(a));
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));
            if (OP == 1) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(b));
            if (OP == 2) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(a));
            if (OP == 3) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb));
            if (OP == 4) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
            if (OP == 5) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pa));
            if (OP == 6) asm volatile("min.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(b));
            if (OP == 7) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b)); asm volatile("min.f32 %0, %0, %1;" : "+f"(x[(i + 4) % NACC]) : "f"(b)); }
            if (OP == 8) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb)); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b)); }
        }
    }
    float s = 0; for (int i = 0; i < NACC; ++i) { float lo, hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i])); s += x[i] + lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP> void run(const char* name, int instr_per_iter, int flops_per_instr) {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench<OP><<<148 * 8, 256>>>(out, 1.0001f, 0.5f);
    cudaEventRecord(e0); bench<OP><<<148 * 8, 256>>>(out, 1.0001f, 0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double warp_instr = 148.0 * 8 * 8 * ITERS * NACC * instr_per_iter;      // warps * iters * NACC
    double per_sm_clk = warp_instr / 148 / (ms * 1e-3 * clk * 1e3);
    printf("%-22s %8.3f ms  %6.3f warp-instr/clk/SM  (%5.1f lane-flop/clk/SM at nominal %d MHz)\n", name, ms, per_sm_clk,
           per_sm_clk * 32 * flops_per_instr, clk / 1000);
    cudaFree(out);
}
int main() {
    run<0>("FFMA", 1, 2); run<1>("FADD", 1, 1); run<2>("FMUL", 1, 1); run<3>("FFMA2", 1, 4); run<4>("FADD2", 1, 2);
    run<5>("FMUL2", 1, 2); run<6>("FMNMX", 1, 1); run<7>("FFMA+FMNMX", 2, 1); run<8>("FFMA2+FFMA", 2, 3);
    return 0;
}
