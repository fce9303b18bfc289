// pointconv_v2.cu — f1, second generation of the fused PointConv body (pointconv.cu is the first).
//
// Same contract as pointconv.cu (models/pointconv.py:33-61 and :90-122 after the k_nearest_neighbor call): gathers, weight
// net (3 -> 8 -> 16), the [16 x k].[k x (C+3)] product per point, nn.Linear(16*(C+3), out) and LeakyReLU(0.1) in one pass.
// What changed, and why (profiles/r2_notes.md):
//   * CTA = 120 sampled points on the M = 128 rows of a cta_group::1 tcgen05.mma (the first kernel filled 64): 15 worker
//     warps + the MMA warp = 512 threads, the largest block that still gets 128 registers per thread.
//   * warp-specialised: 15 worker warps (4 threads per point; a thread keeps the gathered 16 neighbours x 4 channels slab
//     of a 16-channel block in 64 registers and forms one 16-byte chunk of an A row per weight-net output w) and ONE
//     MMA warp that only waits on "full" mbarriers, issues tcgen05.mma and commits to "empty" mbarriers: a 2-3 deep ring
//     instead of one __syncthreads per K block with the issuing thread also producing.
//   * K block = 16 floats (one w, 16 channels): rows of 64 bytes, SWIZZLE_64B operand tiles, which is what lets the
//     128-point weight table (133 KB), the rings and the index table share 227 KB.
//   * 3xTF32 with the two halves of the B operand stacked along N: D[:, 0:N) += Ah.Bh + Al.Bh, D[:, N:2N) += Ah.Bl is two
//     MMAs per k-step instead of three (out_channels <= 128; above that three MMAs); the halves are summed in the epilogue.
//   * the Linear weight is re-laid once per call as one contiguous [hi | lo] image per K block, already in the swizzled
//     operand layout; one thread of warp 0 brings it into the stage with cp.async.bulk (expect_tx on the stage's "full"
//     barrier) as soon as the stage is free; the weight-net outputs of the next K block are fetched under this block's FMAs.
// precision 1 = TF32 operands, 2 (and 0) = 3xTF32.
#include "umma_common.cuh"

namespace b200 {

constexpr int P2_PTS = 120, P2_ROWS = 128, P2_WORKERS = 4 * P2_PTS, P2_THREADS = P2_WORKERS + 32, P2_K = 16, P2_NW = 16, P2_CB = 16;
static_assert(P2_THREADS == 512, "512 threads: 128 registers each (the register file is allocated per 4 warps)");
constexpr int P2_WT_STRIDE = P2_NW * P2_K + 4;           // floats per point in the weight table (+4 rotates the bank group per point)
constexpr int P2_A_TILE = P2_ROWS * 64;                  // one 128-row x 16-float operand tile: 8 KB (rows 120..127 idle)
constexpr int P2_SMEM_MAX = 232448;

__device__ __forceinline__ float p2_lrelu(float v) { return v > 0.0f ? v : 0.1f * v; }

// K-major operand, rows of 64 bytes, 64-byte swizzle (16-byte chunk index ^= address bits 7-8), 8-row groups 512 bytes apart
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)(512u >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}
__device__ __forceinline__ uint32_t p2_sw64(uint32_t r, uint32_t q) { return r * 64u + ((q ^ ((r >> 1) & 3u)) << 4); }

__device__ __forceinline__ void p2_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t p2_mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done;
}
// bounded wait: a protocol bug traps (launch error) instead of hanging the device; the fast path is one try_wait
__device__ __noinline__ void p2_mbar_wait_slow(uint32_t bar, uint32_t parity) {
    for (uint32_t spin = 0; !p2_mbar_try(bar, parity); ++spin)
        if (spin > (1u << 26)) __trap();
}
__device__ __forceinline__ void p2_mbar_wait(uint32_t bar, uint32_t parity) {
    if (!p2_mbar_try(bar, parity)) p2_mbar_wait_slow(bar, parity);
}
__device__ __forceinline__ void p2_tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
}

// ---- pass 0 ---------------------------------------------------------------------------------------------------
// grid (ceil(N/32), B): Fpm[b][n][0..Cp) = [xyz ; feat ; 0] point-major (Cp = roundup16(C+3)); on blockIdx.y == 0 also the
// Linear weight as Limg[kb = cb*16 + w][hi|lo][o][c'] = L[o][w*(C+3) + cb*16 + c'] (zero beyond C+3 / out_channels),
// split into TF32 hi / lo parts (3xTF32) and stored in the swizzled layout of the B operand tile: one bulk copy per K block.
__global__ void __launch_bounds__(256)
pointconv_v2_prep_kernel(const float* __restrict__ xyz, const float* __restrict__ feat, const float* __restrict__ L,
                         float* __restrict__ Fpm, float* __restrict__ Limg, int C, int Cp, int N, int Cout, int Npad, int split) {
    __shared__ float tile[32][33];
    const int b = blockIdx.y, n0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int Cf = C + 3;
    for (int c0 = 0; c0 < Cp; c0 += 32) {
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int c = c0 + ty * 4 + u, n = n0 + tx;
            float v = 0.0f;
            if (n < N && c < Cf) v = c < 3 ? __ldg(xyz + ((size_t)b * 3 + c) * N + n) : __ldg(feat + ((size_t)b * C + (c - 3)) * N + n);
            tile[ty * 4 + u][tx] = v;
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int n = n0 + ty * 4 + u;
            if (n < N && c0 + tx < Cp) Fpm[((size_t)b * N + n) * Cp + c0 + tx] = tile[tx][ty * 4 + u];
        }
    }
    if (b == 0) {
        // K block kb = cb*16 + w: [hi | lo] tiles of Npad rows x 64 bytes, already in the swizzled operand layout
        const int total = P2_NW * Cp * Npad;                                  // (Cp/16 * 16 K blocks) * Npad * 16
        for (int e = blockIdx.x * 256 + threadIdx.x; e < total; e += gridDim.x * 256) {
            const int cc = e & 15, o = (e >> 4) % Npad, kb = (e >> 4) / Npad;
            const int c = (kb >> 4) * P2_CB + cc, w = kb & 15;
            const float v = (o < Cout && c < Cf) ? __ldg(L + (size_t)o * P2_NW * Cf + (size_t)w * Cf + c) : 0.0f;
            const uint32_t off = (p2_sw64((uint32_t)o, (uint32_t)(cc >> 2)) >> 2) + (uint32_t)(cc & 3);
            float* tile = Limg + (size_t)kb * (split ? 2 : 1) * Npad * 16;
            if (split) {
                const float hi = tf32_hi(v);
                tile[off] = hi;
                tile[Npad * 16 + off] = v - hi;
            } else {
                tile[off] = v;
            }
        }
    }
}

// ---- pass 1 ---------------------------------------------------------------------------------------------------
struct P2Smem {
    int ring, a_bytes, b_bytes, stage, wt, jj, small, bars, total;
};
__host__ __device__ inline P2Smem p2_layout(int Npad, int split, int depth) {
    P2Smem S;
    S.a_bytes = (split ? 2 : 1) * P2_A_TILE;
    S.b_bytes = (split ? 2 : 1) * Npad * 64;
    S.stage = S.a_bytes + S.b_bytes;
    int off = 0;                                                               // fixed-offset tables first, the ring after them
    S.wt = off;    off += P2_PTS * P2_WT_STRIDE * 4;
    S.jj = off;    off += P2_PTS * P2_K * 2;                                   // uint16 neighbour indices
    off = (off + 1023) & ~1023;
    S.ring = off;  off += depth * S.stage;
    S.small = off; off += (24 + 8 + 128 + 16 + Npad) * 4;                      // Wa, ba, Wb, bb, bias
    off = (off + 15) & ~15;
    S.bars = off;  off += 128;                                                 // full[4], empty[4], accf, tmem slot
    S.total = off;
    return S;
}

__device__ __forceinline__ void p2_mbar_arrive_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void p2_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// MODE 0: TF32 (one MMA per k-step); 1: 3xTF32, B halves stacked along N (two MMAs, Npad <= 128); 2: 3xTF32, three MMAs.
template <int MODE>
__global__ void __launch_bounds__(P2_THREADS, 1)
pointconv_v2_kernel(const float* __restrict__ xyz, const float* __restrict__ sampled, const int64_t* __restrict__ knn,
                    const float* __restrict__ Fpm, const float* __restrict__ Limg, const float* __restrict__ Wa,
                    const float* __restrict__ ba, const float* __restrict__ Wb, const float* __restrict__ bb,
                    const float* __restrict__ bias, float* __restrict__ out, int Cp, int N, int S, int Cout, int Npad,
                    uint32_t tmem_cols, int depth) {
    constexpr int SPLIT = MODE != 0;
    extern __shared__ uint8_t p2_smem_raw[];
    const uint32_t sbase = (tc_smem_u32(p2_smem_raw) + 1023u) & ~1023u;
    uint8_t* g = p2_smem_raw + (sbase - tc_smem_u32(p2_smem_raw));
    const P2Smem Ls = p2_layout(Npad, SPLIT, depth);
    float* s_wt = reinterpret_cast<float*>(g + Ls.wt);
    uint16_t* s_j = reinterpret_cast<uint16_t*>(g + Ls.jj);
    float* s_small = reinterpret_cast<float*>(g + Ls.small);      // Wa[24] ba[8] Wb[128] bb[16] bias[Npad]
    const uint32_t bars = sbase + Ls.bars;                        // full[s] = bars + 8s, empty[s] = bars + 32 + 8s
    const uint32_t bar_accf = bars + 64u, tmem_slot = bars + 72u;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(g + Ls.bars + 72);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y, s0 = blockIdx.x * P2_PTS;
    const int ncb = Cp / P2_CB, nkb = ncb * P2_NW;
    const uint32_t stage_bytes = (uint32_t)Ls.stage, b_bytes = (uint32_t)Ls.b_bytes;

    if (tid == 0) {
        for (uint32_t s = 0; s < 4; ++s) {
            tc_mbar_init(bars + 8u * s, P2_WORKERS / 32 + 1);      // one arrival per worker warp + the bulk copy's expect_tx arrival
            tc_mbar_init(bars + 32u + 8u * s, 1);                  // stage free: tcgen05.commit
        }
        tc_mbar_init(bar_accf, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == P2_WORKERS / 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int e = tid; e < 176 + Npad; e += P2_THREADS) {
        float v;
        if (e < 24) v = __ldg(Wa + e);
        else if (e < 32) v = __ldg(ba + e - 24);
        else if (e < 160) v = __ldg(Wb + e - 32);
        else if (e < 176) v = __ldg(bb + e - 160);
        else v = e - 176 < Cout ? __ldg(bias + e - 176) : 0.0f;
        s_small[e] = v;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot_ptr;

    if (warp == P2_WORKERS / 32) {
        // ================================================ MMA warp ================================================
        if (lane == 0) {
            const uint32_t idesc1 = umma_idesc_tf32(128, Npad), idesc2 = umma_idesc_tf32(128, 2 * Npad);
            const uint32_t a_bytes = (uint32_t)Ls.a_bytes, blo = (uint32_t)Npad * 64u;
            uint32_t s = 0, ph = 0;
            for (int kbi = 0; kbi < nkb; ++kbi) {
                p2_mbar_wait(bars + 8u * s, ph & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_hi = sbase + Ls.ring + s * stage_bytes, a_lo = a_hi + P2_A_TILE;
                const uint32_t b_hi = a_hi + a_bytes, b_lo = b_hi + blo;
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    const uint64_t ah = umma_desc_sw64(a_hi + ks * 32), bh = umma_desc_sw64(b_hi + ks * 32);
                    const uint32_t first = (kbi | ks) != 0;
                    if (MODE == 0) {
                        umma_tf32(tmem, ah, bh, idesc1, first);
                    } else if (MODE == 1) {
                        umma_tf32(tmem, ah, bh, idesc2, first);                                   // [Ah.Bh | Ah.Bl]
                        umma_tf32(tmem, umma_desc_sw64(a_lo + ks * 32), bh, idesc1, 1u);         // + Al.Bh
                    } else {
                        umma_tf32(tmem, ah, bh, idesc1, first);
                        umma_tf32(tmem, ah, umma_desc_sw64(b_lo + ks * 32), idesc1, 1u);
                        umma_tf32(tmem, umma_desc_sw64(a_lo + ks * 32), bh, idesc1, 1u);
                    }
                }
                umma_commit(bars + 32u + 8u * s);
                if (++s == (uint32_t)depth) { s = 0; ++ph; }
            }
            umma_commit(bar_accf);
        }
        __syncwarp();
    } else {
        // ================================================ workers =================================================
        const int p = tid >> 2, q = tid & 3;                    // 4 threads per point; q = neighbours 4q..4q+3 (meta) / channels 4q..4q+3
        {   // ---- meta: neighbour indices + the weight net of the point's 16 neighbours -> shared memory
            const int i = min(s0 + p, S - 1);
            float sx[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) sx[a] = __ldg(sampled + ((size_t)b * 3 + a) * S + i);
            const longlong2* kp = reinterpret_cast<const longlong2*>(knn + ((size_t)b * S + i) * P2_K + 4 * q);
            const longlong2 k01 = __ldg(kp), k23 = __ldg(kp + 1);
            const int64_t jraw[4] = {k01.x, k01.y, k23.x, k23.y};
            float wv[P2_NW][4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                int64_t j = jraw[u];
                if (j < 0) j += N;
                j = j < 0 ? 0 : (j >= N ? N - 1 : j);
                s_j[p * P2_K + 4 * q + u] = (uint16_t)j;
                float d[3], h1[8];
#pragma unroll
                for (int a = 0; a < 3; ++a) d[a] = __ldg(xyz + ((size_t)b * 3 + a) * N + j) - sx[a];
#pragma unroll
                for (int o = 0; o < 8; ++o) {
                    float v = s_small[24 + o];
#pragma unroll
                    for (int a = 0; a < 3; ++a) v = fmaf(s_small[o * 3 + a], d[a], v);
                    h1[o] = p2_lrelu(v);
                }
#pragma unroll
                for (int o = 0; o < P2_NW; ++o) {
                    float v = s_small[160 + o];
#pragma unroll
                    for (int m = 0; m < 8; ++m) v = fmaf(s_small[32 + o * 8 + m], h1[m], v);
                    wv[o][u] = p2_lrelu(v);
                }
            }
#pragma unroll
            for (int o = 0; o < P2_NW; ++o)
                *reinterpret_cast<float4*>(s_wt + p * P2_WT_STRIDE + o * P2_K + 4 * q) = make_float4(wv[o][0], wv[o][1], wv[o][2], wv[o][3]);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(P2_WORKERS) : "memory");

        // ---- K loop: K block kbi = (cb, w).  Per block: wait for the stage, 64 FMAs with the weights fetched one block ahead, one
        // 16-byte chunk of the A row (hi, lo) into the stage, one arrival per warp.
        const float* frow = Fpm + (size_t)b * N * Cp + 4 * q;
        const float4* wt4 = reinterpret_cast<const float4*>(s_wt + p * P2_WT_STRIDE);
        uint8_t* const st0 = g + Ls.ring + p2_sw64((uint32_t)p, (uint32_t)q);
        const uint8_t* bsrc = reinterpret_cast<const uint8_t*>(Limg);      // this K block's B tile in global memory
        const uint32_t bdst0 = sbase + Ls.ring + (uint32_t)Ls.a_bytes;
        uint8_t* st = st0;                                      // this thread's chunk of the A row in the current stage
        uint32_t bdst = bdst0, bar_f = bars, s = 0, par_e = 1;  // parity 1 on a fresh mbarrier passes at once (first round)
        float4 wc[4];
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) wc[k4] = wt4[k4];
        for (int cb = 0; cb < ncb; ++cb) {
            float f[P2_K][4];
            {
                const uint4 j0 = *reinterpret_cast<const uint4*>(s_j + p * P2_K), j1 = *reinterpret_cast<const uint4*>(s_j + p * P2_K + 8);
                const uint32_t jw[8] = {j0.x, j0.y, j0.z, j0.w, j1.x, j1.y, j1.z, j1.w};
#pragma unroll
                for (int kk = 0; kk < P2_K; ++kk) {
                    const uint32_t j = (jw[kk >> 1] >> ((kk & 1) * 16)) & 0xffffu;
                    const float4 v = __ldg(reinterpret_cast<const float4*>(frow + (size_t)j * Cp + cb * P2_CB));
                    f[kk][0] = v.x; f[kk][1] = v.y; f[kk][2] = v.z; f[kk][3] = v.w;
                }
            }
#pragma unroll 1
            for (int w = 1; w <= P2_NW; ++w) {
                if (warp == 0) {                                // warp 0 takes the stage first: one thread starts the bulk copy of the B tile
                    p2_mbar_wait(bar_f + 32u, par_e);           // the MMAs that read this stage are done
                    if (lane == 0) {
                        p2_mbar_arrive_tx(bar_f, b_bytes);
                        p2_bulk_g2s(bdst, bsrc, b_bytes, bar_f);
                    }
                    bsrc += b_bytes;
                }
                const float4* wnx = wt4 + (w & 15) * 4;         // next block's weights: each quarter re-loaded right after its use
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                    const float ws[4] = {wc[k4].x, wc[k4].y, wc[k4].z, wc[k4].w};
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                        for (int c = 0; c < 4; ++c) acc[c] = fmaf(ws[u], f[k4 * 4 + u][c], acc[c]);
                    wc[k4] = wnx[k4];
                }
                if (warp != 0) p2_mbar_wait(bar_f + 32u, par_e);    // the other warps need the stage only now, after their FMAs
                if (SPLIT) {
                    const float4 hi = make_float4(tf32_hi(acc[0]), tf32_hi(acc[1]), tf32_hi(acc[2]), tf32_hi(acc[3]));
                    *reinterpret_cast<float4*>(st) = hi;
                    *reinterpret_cast<float4*>(st + P2_A_TILE) = make_float4(acc[0] - hi.x, acc[1] - hi.y, acc[2] - hi.z, acc[3] - hi.w);
                } else {
                    *reinterpret_cast<float4*>(st) = make_float4(acc[0], acc[1], acc[2], acc[3]);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) p2_mbar_arrive(bar_f);
                if (++s == (uint32_t)depth) { s = 0; par_e ^= 1u; st = st0; bdst = bdst0; bar_f = bars; }
                else { st += stage_bytes; bdst += stage_bytes; bar_f += 8u; }
            }
        }

    }

    // ---- epilogue, all 16 warps: warp -> TMEM lane quarter (warp & 3) = 32 points, column group (warp >> 2) = Npad/4 channels
    {
        p2_mbar_wait(bar_accf, 0u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int qd = warp & 3, cg = warp >> 2, ncol = Npad >> 2;
        const int row = qd * 32 + lane, sp = s0 + row;
        const bool live = row < P2_PTS && sp < S;
        const uint32_t tbase = tmem + ((uint32_t)(qd * 32) << 16);
        for (int c0 = cg * ncol; c0 < (cg + 1) * ncol; c0 += 8) {
            uint32_t r0[8], r1[8];
            p2_tmem_ld8(tbase + (uint32_t)c0, r0);
            if (MODE == 1) p2_tmem_ld8(tbase + (uint32_t)(Npad + c0), r1);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int o = c0 + c;
                float v = __uint_as_float(r0[c]);
                if (MODE == 1) v += __uint_as_float(r1[c]);
                if (o < Cout && live) out[((size_t)b * Cout + o) * S + sp] = p2_lrelu(v + s_small[176 + o]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == P2_WORKERS / 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols) : "memory");
}

int p2_cp(int C) { return (C + 3 + 15) / 16 * 16; }
int p2_npad(int Cout) { return (Cout + 31) / 32 * 32; }

template <int MODE>
static int p2_launch(const float* xyz, const float* sampled, const int64_t* knn, const float* Fpm, const float* Limg,
                     const b200_pointconv_weights* w, float* out, int B, int Cp, int N, int S, int Cout, int Npad, cudaStream_t st) {
    const int split = MODE != 0;
    int depth = 4;
    while (depth > 1 && p2_layout(Npad, split, depth).total + 1024 > P2_SMEM_MAX) --depth;
    const size_t smem = (size_t)p2_layout(Npad, split, depth).total + 1024;
    if (smem > (size_t)P2_SMEM_MAX) {
        set_error("b200_pointconv_fwd: out_channels = %d does not fit the shared-memory budget", Cout);
        return B200_ENOSUP;
    }
    const int ncols = MODE == 1 ? 2 * Npad : Npad;
    uint32_t cols = 32;
    while ((int)cols < ncols) cols <<= 1;
    cudaError_t e = cudaFuncSetAttribute(pointconv_v2_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "pointconv_v2_kernel(attr)");
    pointconv_v2_kernel<MODE><<<dim3(ceil_div(S, P2_PTS), B), P2_THREADS, smem, st>>>(
        xyz, sampled, knn, Fpm, Limg, w->Wa, w->ba, w->Wb, w->bb, w->bias, out, Cp, N, S, Cout, Npad, cols, depth);
    B200_LAUNCH_CHECK("pointconv_v2_kernel");
    return B200_OK;
}

// floats of scratch the second-generation route needs (b200_pointconv_scratch_floats takes the larger of the two routes)
int64_t pointconv_v2_scratch_floats(int B, int C, int Cout, int N) {
    const int64_t cp = p2_cp(C);
    return (((int64_t)B * N * cp + 63) & ~int64_t(63)) + 2 * (int64_t)p2_npad(Cout) * P2_NW * cp;
}

// called by b200_pointconv_fwd (pointconv.cu) after its argument checks; scratch sized by b200_pointconv_scratch_floats
int pointconv_v2_run(const float* xyz, const float* feat, const float* sampled_xyz, const int64_t* knn,
                     const b200_pointconv_weights* w, float* out, float* scratch, int B, int C, int Cout, int N, int S,
                     int precision, cudaStream_t st) {
    const int Cp = p2_cp(C), Npad = p2_npad(Cout);
    float* Fpm = scratch;
    float* Limg = scratch + (((int64_t)B * N * Cp + 63) & ~int64_t(63));
    pointconv_v2_prep_kernel<<<dim3(ceil_div(N, 32), B), 256, 0, st>>>(xyz, feat, w->L, Fpm, Limg, C, Cp, N, Cout, Npad, precision != 1);
    B200_LAUNCH_CHECK("pointconv_v2_prep_kernel");
    if (precision == 1) return p2_launch<0>(xyz, sampled_xyz, knn, Fpm, Limg, w, out, B, Cp, N, S, Cout, Npad, st);
    return Npad <= 128 ? p2_launch<1>(xyz, sampled_xyz, knn, Fpm, Limg, w, out, B, Cp, N, S, Cout, Npad, st)
                       : p2_launch<2>(xyz, sampled_xyz, knn, Fpm, Limg, w, out, B, Cp, N, S, Cout, Npad, st);
}

}  // namespace b200
