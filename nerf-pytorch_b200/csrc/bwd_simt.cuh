// bwd_simt.cuh -- exact-mode (fp32 CUDA-core) backward of one network pass: dL/dtheta given dL/drgb_map.
//
// Replaces autograd's backward through run_nerf.py:381-386 / :397-403: the pass is
// recomputed in fp32 with saved activations (mlp_simt_kernel, save mode), the compositing adjoint
//  gives dL/draw, and the MLP is walked backwards with two generic fp32 GEMMs
// (dgrad: C = A B, wgrad: C += A^T B reduced over all sample rows), a ReLU-mask and a column-sum
// kernel.  No gradient flows to rays / z (run_nerf.py:394).
// This is the first correct backward (round 1); the tcgen05 dgrad/wgrad kernels replace it in round 2.
#pragma once
#include "common.cuh"

namespace nb {

constexpr int GT = 64, GK = 16;   // GEMM tile: 64 x 64 outputs, K step 16, 256 threads x (4 x 4)

// C[M,N] (ldc) = (beta ? C : 0) + A[M,K] (lda) * B[K,N] (ldb)        -- dgrad
__global__ void __launch_bounds__(256) sgemm_nn_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                                                      float* __restrict__ C, int ldc, long long M, int N, int K, int beta) {
  __shared__ float sA[GK][GT + 4], sB[GK][GT + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const long long m0 = (long long)blockIdx.x * GT;
  const int n0 = blockIdx.y * GT;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += GK) {
    for (int i = threadIdx.x; i < GT * GK; i += 256) {
      int r = i / GK, k = i % GK;                     // A tile: [64 rows][16 k], k contiguous in memory
      long long m = m0 + r;
      sA[k][r] = (m < M && k0 + k < K) ? A[m * lda + k0 + k] : 0.0f;
      int kk = i / GT, c = i % GT;                    // B tile: [16 k][64 n], n contiguous
      sB[kk][c] = (k0 + kk < K && n0 + c < N) ? B[(size_t)(k0 + kk) * ldb + n0 + c] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = sA[k][ty * 4 + i]; b[i] = sB[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n < N) { float* c = C + m * ldc + n; *c = (beta ? *c : 0.0f) + acc[i][j]; }
    }
  }
}

// C[K1,N] (ldc) += A[M,K1]^T (lda) * B[M,N] (ldb), reduction over the M sample rows, split across
// blockIdx.z slabs of `rows_per_slab` rows; partial tiles are combined with fp32 atomics.   -- wgrad
__global__ void __launch_bounds__(256) sgemm_tn_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                                                      float* __restrict__ C, int ldc, long long M, int K1, int N, int rows_per_slab) {
  __shared__ float sA[GK][GT + 4], sB[GK][GT + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int i0 = blockIdx.x * GT, n0 = blockIdx.y * GT;
  const long long mb = (long long)blockIdx.z * rows_per_slab;
  const long long me = (mb + rows_per_slab < M) ? mb + rows_per_slab : M;
  float acc[4][4] = {};
  for (long long m0 = mb; m0 < me; m0 += GK) {
    for (int i = threadIdx.x; i < GT * GK; i += 256) {
      int mm = i / GT, c = i % GT;                    // both tiles: [16 rows][64 cols], cols contiguous
      long long m = m0 + mm;
      sA[mm][c] = (m < me && i0 + c < K1) ? A[m * lda + i0 + c] : 0.0f;
      sB[mm][c] = (m < me && n0 + c < N) ? B[m * ldb + n0 + c] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = sA[k][ty * 4 + i]; b[i] = sB[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int r = i0 + ty * 4 + i;
    if (r >= K1) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n < N) atomicAdd(C + (size_t)r * ldc + n, acc[i][j]);
    }
  }
}

// d[m, c] *= (h[m, c] > 0)   (ReLU backward, in place), and colsum[c] += sum_m d[m, c]  (bias gradient)
__global__ void relu_mask_colsum_kernel(float* __restrict__ d, int ldd, const float* __restrict__ h, int ldh,
                                        long long M, int C, float* __restrict__ colsum, int rows_per_block) {
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const long long m0 = (long long)blockIdx.x * rows_per_block;
  const long long m1 = (m0 + rows_per_block < M) ? m0 + rows_per_block : M;
  float s = 0.0f;
  for (long long m = m0; m < m1; ++m) {
    float v = d[m * ldd + c];
    if (h != nullptr) { v = (h[m * ldh + c] > 0.0f) ? v : 0.0f; d[m * ldd + c] = v; }
    s += v;
  }
  if (colsum) atomicAdd(colsum + c, s);
}

// activations saved by the exact forward (row-major fp32), all optional
struct SimtSave {
  float* enc;    // [M, IC]
  float* encv;   // [M, ICV]  (per-row copy of the ray's view encoding)
  float* h;      // [D][M, W] post-ReLU outputs of pts_linears
  float* feat;   // [M, W]    feature_linear output
  float* hv;     // [M, W/2]  post-ReLU views layer
};

static inline size_t bwd_floats_per_row(const NerfNetParams& n) {
  // enc + encv + h[D] + feat + hv + raw + d_raw + dh ping/pong + d_hv
  return (size_t)n.input_ch + n.input_ch_views + (size_t)n.D * n.W + n.W + n.W / 2 + 4 + 4 + 2 * (size_t)n.W + n.W / 2;
}

}  // namespace nb
