/*
 * oracle/ref_shim.cu — extern "C" doorway onto the UNMODIFIED reference CUDA kernels.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/README.md).  oracle/Makefile compiles the reference's own kernel
 * sources where they lie under /root/reference/models/csrc/ (they only need cuda_runtime.h) together with
 * this file into oracle/_ref/libref_kernels.so.  No reference source is copied into this repo; the
 * declarations below are the host launchers those files define (C++ linkage):
 *   correlation_forward_kernel.cu:51, correlation_backward_kernel.cu:76,
 *   furthest_point_sampling_kernel.cu:81, k_nearest_neighbor_kernel.cu:96,105.
 * The reference launches on the legacy default stream and never checks errors, so each doorway
 * synchronises and returns cudaGetLastError().
 */
#include <cuda_runtime.h>
#include <stdint.h>

void correlation_forward_kernel_wrapper(float* output, const float* input1, const float* input2,
                                        int n_batches, int in_channels, int height, int width, int max_displacement);
void correlation_backward_kernel_wrapper(const float* grad_output, float* grad_input1, float* grad_input2,
                                         const float* input1, const float* input2,
                                         int n_batches, int in_channels, int height, int width, int max_displacement);
void furthest_point_sampling_kernel_wrapper(float* batched_points_xyz, float* batched_dists_temp, int n_batch,
                                            int n_points, int n_samples, int64_t* batched_furthest_indices);
void k_nearest_neighbor_2d_kernel_wrapper(int b, int n, int m, int k, const float* query_xyz,
                                          const float* input_xyz, int64_t* indices);
void k_nearest_neighbor_3d_kernel_wrapper(int b, int n, int m, int k, const float* query_xyz,
                                          const float* input_xyz, int64_t* indices);

static int finish(int sync) {
    if (sync) cudaDeviceSynchronize();
    return (int)cudaGetLastError();
}

extern "C" {
/* out must be zero-filled by the caller (correlation.cpp:17 uses torch::zeros). */
int ref_corr2d_fwd(const float* in1, const float* in2, float* out, int B, int C, int H, int W, int md, int sync) {
    correlation_forward_kernel_wrapper(out, in1, in2, B, C, H, W, md);
    return finish(sync);
}
int ref_corr2d_bwd(const float* gout, const float* in1, const float* in2, float* g1, float* g2,
                   int B, int C, int H, int W, int md, int sync) {
    correlation_backward_kernel_wrapper(gout, g1, g2, in1, in2, B, C, H, W, md);
    return finish(sync);
}
/* dists_temp [B,N] must be filled with 1e10 by the caller (furthest_point_sampling.cpp:12). */
int ref_fps(float* xyz, float* dists_temp, int64_t* idx, int B, int N, int S, int sync) {
    furthest_point_sampling_kernel_wrapper(xyz, dists_temp, B, N, S, idx);
    return finish(sync);
}
/* idx must be zero-filled by the caller (k_nearest_neighbor.cpp:16). */
int ref_knn(const float* input, const float* query, int64_t* idx, int B, int M, int Q, int D, int k, int sync) {
    if (D == 2) k_nearest_neighbor_2d_kernel_wrapper(B, Q, M, k, query, input, idx);
    else        k_nearest_neighbor_3d_kernel_wrapper(B, Q, M, k, query, input, idx);
    return finish(sync);
}
}
