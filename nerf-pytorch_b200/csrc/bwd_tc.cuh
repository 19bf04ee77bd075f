// bwd_tc.cuh -- EXPERIMENTAL building blocks of the tensor-core backward.  Not on any
// default path: reachable only through the nerf_b200_exp_* entry points, exercised by tests/test_gpu_exp_bwd.py
// when NERF_B200_EXPERIMENTAL=1.  Written at the end of round 1 (GPU budget spent), to be validated first thing
// in round 2.  Everything here reuses operand encodings that ARE validated on B200:
//   * K-major SWIZZLE_128B A tiles written in the activation layout (march kernels, selftest_gemm),
//   * MN-major SWIZZLE_128B operands read from the same layout (selftest_gemm_tn: LBO = stride between 64-column
//     groups, SBO = 1024 = stride between 8-row groups, one K=16 step = +2048 bytes).
//
// "Tile image": a [128 rows x C columns] fp16 tile (C in {64,128,256}) stored in GLOBAL memory byte-for-byte as the
// shared-memory activation tile: C/64 K-blocks of 16 KB; inside a K-block row r sits at (r>>3)*1024 + (r&7)*128 and
// its 16-byte chunk c at ((c ^ (r&7)) << 4).  A [M x C] matrix = ceil(M/128) images.  One cp.async.bulk moves an
// image (or a 64-row half of a K-block: 8 KB) into shared memory ready for the MMA, no tensor map needed.
//
//   wgrad_tiles_kernel : dW[Mc x Nc] += sum over tiles of  X_tile^T (128 x Mc)  Y_tile (128 x Nc)      (dA^T H)
//   dgrad_tiles_kernel : OUT_tile (128 x 256) = X_tile (128 x Kc) W (Kc x 256), optional ReLU mask from H, fp16 image out
//   tile_pack / tile_unpack / tile_colsum : fp32 row-major <-> images, column sums (bias gradients)
#pragma once
#include "fused_tc.cuh"

namespace nb {

__host__ __device__ __forceinline__ uint32_t img_off(int r, int col) {
  return (uint32_t)((col >> 6) * 16384 + (r >> 3) * 1024 + (r & 7) * 128 + ((((col & 63) >> 3) ^ (r & 7)) << 4) + (col & 7) * 2);
}
__host__ __device__ __forceinline__ size_t img_bytes(int cols) { return (size_t)(cols >> 6) * 16384; }

// fp32 row-major [M, ncols] (row stride ld) -> tile images of C >= ncols columns (zero padded), value * scale rounded
// to fp16; rows >= M are zero
__global__ void tile_pack_kernel(const float* __restrict__ src, int ld, long long M, int ncols, int C, float scale, uint8_t* __restrict__ img) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;          // one thread per 8-column chunk
  const int cpr = C >> 3;
  const long long Mpad = ((M + 127) / 128) * 128;
  if (i >= Mpad * cpr) return;
  const long long row = i / cpr;
  const int c8 = (int)(i - row * cpr);
  float x[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = (row < M && c8 * 8 + j < ncols) ? src[row * ld + c8 * 8 + j] * scale : 0.f;
  uint8_t* dst = img + (size_t)(row >> 7) * img_bytes(C) + img_off((int)(row & 127), c8 * 8);
  *reinterpret_cast<uint4*>(dst) = make_uint4(ptx::cvt_f16x2(x[0], x[1]), ptx::cvt_f16x2(x[2], x[3]), ptx::cvt_f16x2(x[4], x[5]), ptx::cvt_f16x2(x[6], x[7]));
}

// tile images -> fp32 row-major [M, C] (row stride ld), value * scale
__global__ void tile_unpack_kernel(const uint8_t* __restrict__ img, long long M, int C, float scale, float* __restrict__ dst, int ld) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int cpr = C >> 3;
  if (i >= M * cpr) return;
  const long long row = i / cpr;
  const int c8 = (int)(i - row * cpr);
  const uint4 v = *reinterpret_cast<const uint4*>(img + (size_t)(row >> 7) * img_bytes(C) + img_off((int)(row & 127), c8 * 8));
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  float* d = dst + row * ld + c8 * 8;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __half2 hh = *reinterpret_cast<const __half2*>(&w[j]);
    d[2 * j] = __low2float(hh) * scale; d[2 * j + 1] = __high2float(hh) * scale;
  }
}

// column sums of a tile-image matrix (bias gradient = sum over sample rows of dA): colsum[c] += scale * sum_r X[r][c]
__global__ void tile_colsum_kernel(const uint8_t* __restrict__ img, long long n_tiles, int C, float scale, float* __restrict__ colsum) {
  // block = one group of tiles; thread = (row group of 8 rows) x (column chunk): 16 x cpr threads cooperate per tile
  extern __shared__ float s_sum[];                     // [C]
  for (int c = threadIdx.x; c < C; c += blockDim.x) s_sum[c] = 0.f;
  __syncthreads();
  const int cpr = C >> 3;
  for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const uint8_t* base = img + (size_t)t * img_bytes(C);
    for (int i = threadIdx.x; i < 128 * cpr; i += blockDim.x) {
      const int row = i / cpr, c8 = i - row * cpr;
      const uint4 v = *reinterpret_cast<const uint4*>(base + img_off(row, c8 * 8));
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __half2 hh = *reinterpret_cast<const __half2*>(&w[j]);
        atomicAdd(&s_sum[c8 * 8 + 2 * j], __low2float(hh));
        atomicAdd(&s_sum[c8 * 8 + 2 * j + 1], __high2float(hh));
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(colsum + c, s_sum[c] * scale);
}

// ---------------------------------------------------------------------------------------------------------------
// wgrad: dW[Mc, n_valid <= Nc] (fp32, row-major, ld = ldw) += scale * sum_t X_t^T Y_t over this CTA's tiles (grid-stride).
// X images have Mc columns (Mc in {128, 256}), Y images Nc columns (Nc in {64, 128, 256}).  Both operands MN-major.
// Ring: 2 stages of one 64-row half tile of X and of Y (<= 2 x 64 KB); accumulators: TMEM columns [0, Nc) for output
// rows 0-127 and [256, 256 + Nc) for rows 128-255, kept across all tiles; one fp32 atomic flush per CTA.
// 192 threads: warps 0-3 epilogue (TMEM lane quadrants), warp 4 producer, warp 5 issuer.
// ---------------------------------------------------------------------------------------------------------------
constexpr int WG_THREADS = 192;
constexpr uint32_t WG_STAGE = 65536, WG_BARS = 2 * WG_STAGE, WG_TOTAL = WG_BARS + 128;

__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_tiles_kernel(const uint8_t* __restrict__ ximg, const uint8_t* __restrict__ yimg,
                                                                   long long n_tiles, int Mc, int Nc, float scale,
                                                                   float* __restrict__ dW, int ldw, int n_valid) {
  uint8_t* smem = tc_smem;
  const uint32_t sb = ptx::smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_full = sb + WG_BARS, bar_empty = sb + WG_BARS + 16, bar_done = sb + WG_BARS + 32;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + WG_BARS + 64);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(bar_full + 8 * i, 1); ptx::mbar_init(bar_empty + 8 * i, 1); }
    ptx::mbar_init(bar_done, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) { ptx::tmem_alloc(ptx::smem_u32(s_tmem), 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *s_tmem;
  const int xkb = Mc >> 6, ykb = Nc >> 6;                      // K-blocks (64-column groups) per image
  const uint32_t ybase = (uint32_t)xkb * 8192u;                // Y half tile follows the X half tile inside a stage
  long long my_tiles = 0;
  for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) ++my_tiles;
  const long long my_halves = 2 * my_tiles;

  if (warp == 4) {
    // producer: per half tile, one 8 KB bulk copy per K-block of X and of Y
    uint32_t st = 0, ph = 0;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      for (int h = 0; h < 2; ++h) {
        ptx::mbar_wait(bar_empty + 8 * st, ph ^ 1);
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(bar_full + 8 * st, (uint32_t)(xkb + ykb) * 8192u);
          const uint8_t* xs = ximg + (size_t)t * img_bytes(Mc) + h * 8192;
          const uint8_t* ys = yimg + (size_t)t * img_bytes(Nc) + h * 8192;
          for (int kb = 0; kb < xkb; ++kb) ptx::bulk_g2s(sb + st * WG_STAGE + kb * 8192, xs + (size_t)kb * 16384, 8192, bar_full + 8 * st);
          for (int kb = 0; kb < ykb; ++kb) ptx::bulk_g2s(sb + st * WG_STAGE + ybase + kb * 8192, ys + (size_t)kb * 16384, 8192, bar_full + 8 * st);
        }
        __syncwarp();
        if (++st == 2) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 5) {
    // issuer: 4 K-steps (16 sample rows each) per half tile, for each 128-row half of the output
    uint32_t st = 0, ph = 0;
    const uint32_t idesc = ptx::umma_idesc_f16_major(128, Nc, 1, 1);
    const int mhalves = Mc >> 7;
    for (long long i = 0; i < my_halves; ++i) {
      ptx::mbar_wait(bar_full + 8 * st, ph);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint32_t xb = sb + st * WG_STAGE, yb = xb + ybase;
        for (int mh = 0; mh < mhalves; ++mh)
          for (int k = 0; k < 4; ++k) {
            // MN-major SWIZZLE_128B: LBO = stride between 64-column groups (8 KB: half-tile K-blocks), SBO = 1 KB
            const uint64_t ad = ptx::umma_desc_full(xb + (uint32_t)(mh * 2) * 8192u + k * 2048, 8192, 1024, ptx::UMMA_SW128);
            const uint64_t bd = ptx::umma_desc_full(yb + k * 2048, 8192, 1024, ptx::UMMA_SW128);
            ptx::mma_f16_ss(tmem + mh * 256, ad, bd, idesc, (i > 0 || k > 0) ? 1u : 0u);
          }
        ptx::mma_commit(bar_empty + 8 * st);
        if (i == my_halves - 1) ptx::mma_commit(bar_done);
      }
      __syncwarp();
      if (++st == 2) { st = 0; ph ^= 1; }
    }
  }
  if (warp < 4 && my_halves > 0) {
    // epilogue: TMEM lane = output row inside the 128-row half; fp32 atomics into dW
    ptx::mbar_wait(bar_done, 0);
    ptx::tc_fence_after();
    const int r = 32 * warp + lane;
    for (int mh = 0; mh < (Mc >> 7); ++mh)
      for (int c0 = 0; c0 < Nc; c0 += 32) {
        uint32_t v[32];
        ptx::tmem_ld_x32(tmem + ((uint32_t)(32 * warp) << 16) + mh * 256 + c0, v);
        ptx::tmem_ld_wait();
        float* o = dW + (size_t)(mh * 128 + r) * ldw + c0;
#pragma unroll
        for (int j = 0; j < 32; ++j) if (c0 + j < n_valid) atomicAdd(o + j, __uint_as_float(v[j]) * scale);
      }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------------------------
// dgrad: OUT_t (128 x 256, fp16 image) = mask( X_t (128 x Kc) * W (Kc x 256) ) per tile (grid-stride); the ReLU
// mask is (H_t > 0) read from the H image (nullptr: none).  W is a [Kc x 256] fp16 matrix given as tile images
// (rows = K index), resident in shared memory for the whole launch and read as an MN-major B operand; X tiles are
// K-major A operands (the forward's layout).  Values stay in the caller's loss-scaled units.
// 160 threads: warps 0-3 epilogue, warp 4 loads + issues.  Serial per tile except that the next tile's load overlaps
// the epilogue (the HBM stream, not the tensor pipe, bounds this kernel: 64 KB in + 64 KB mask + 64 KB out per tile).
// ---------------------------------------------------------------------------------------------------------------
constexpr int DG_THREADS = 160;
constexpr uint32_t DG_W = 0, DG_X = 131072, DG_BARS = DG_X + 65536, DG_TOTAL = DG_BARS + 128;

__global__ void __launch_bounds__(DG_THREADS, 1) dgrad_tiles_kernel(const uint8_t* __restrict__ ximg, const uint8_t* __restrict__ wimg,
                                                                   const uint8_t* __restrict__ himg, long long n_tiles, int Kc,
                                                                   uint8_t* __restrict__ oimg) {
  uint8_t* smem = tc_smem;
  const uint32_t sb = ptx::smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_w = sb + DG_BARS, bar_x = sb + DG_BARS + 8, bar_d = sb + DG_BARS + 16, bar_e = sb + DG_BARS + 24;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + DG_BARS + 64);
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar_w, 1); ptx::mbar_init(bar_x, 1); ptx::mbar_init(bar_d, 1); ptx::mbar_init(bar_e, 4);
    ptx::fence_mbar_init();
  }
  if (warp == 0) { ptx::tmem_alloc(ptx::smem_u32(s_tmem), 256); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *s_tmem;
  const int xkb = Kc >> 6;                                     // K-blocks of an X image
  const uint32_t xbytes = (uint32_t)xkb * 16384u;

  if (warp == 4) {
    uint32_t phx = 0, phe = 0;
    if (ptx::elect_one()) {                                    // the whole weight matrix, once: Kc/128 images of 64 KB
      ptx::mbar_arrive_expect_tx(bar_w, (uint32_t)(Kc >> 7) * 65536u);
      for (int i = 0; i < (Kc >> 7); ++i) ptx::bulk_g2s(sb + DG_W + i * 65536, wimg + (size_t)i * 65536, 65536, bar_w);
    }
    __syncwarp();
    ptx::mbar_wait(bar_w, 0);
    const uint32_t idesc = ptx::umma_idesc_f16_major(128, 256, 0, 1);
    bool first = true;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      if (first && ptx::elect_one()) {                         // later tiles are prefetched right after the previous tile's MMAs
        ptx::mbar_arrive_expect_tx(bar_x, xbytes);
        ptx::bulk_g2s(sb + DG_X, ximg + (size_t)t * xbytes, xbytes, bar_x);
      }
      __syncwarp();
      ptx::mbar_wait(bar_x, phx); phx ^= 1;
      if (!first) { ptx::mbar_wait(bar_e, phe); phe ^= 1; }    // accumulator drained by the epilogue of the previous tile
      first = false;
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        for (int k = 0; k < (Kc >> 4); ++k) {
          // A: K-major SWIZZLE_128B activation tile, K-step = 32 bytes inside the 64-column K-block
          const uint64_t ad = ptx::umma_desc(sb + DG_X + (uint32_t)(k >> 2) * 16384u + (uint32_t)(k & 3) * 32u, 1024, ptx::UMMA_SW128);
          // B: W rows 16k..16k+15 (K) x 256 columns (N), MN-major: LBO = 16 KB (64-column groups), SBO = 1 KB (8-row groups)
          const uint64_t bd = ptx::umma_desc_full(sb + DG_W + (uint32_t)(k >> 3) * 65536u + (uint32_t)(k & 7) * 2048u, 16384, 1024, ptx::UMMA_SW128);
          ptx::mma_f16_ss(tmem, ad, bd, idesc, k > 0 ? 1u : 0u);
        }
        ptx::mma_commit(bar_d);
      }
      __syncwarp();
      // X buffer is free once the MMAs completed: wait for that, then prefetch the next tile under the epilogue
      const long long tn = t + gridDim.x;
      if (tn < n_tiles) {
        ptx::mbar_wait(bar_d, phx ^ 1);                        // bar_d completes once per tile; phx was toggled above
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(bar_x, xbytes);
          ptx::bulk_g2s(sb + DG_X, ximg + (size_t)tn * xbytes, xbytes, bar_x);
        }
        __syncwarp();
      }
    }
  } else {
    const int r = 32 * warp + lane;
    uint32_t phd = 0;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      ptx::mbar_wait(bar_d, phd); phd ^= 1;
      ptx::tc_fence_after();
      const uint8_t* hrow = himg ? himg + (size_t)t * 65536 : nullptr;
      uint8_t* orow = oimg + (size_t)t * 65536;
      for (int c0 = 0; c0 < 256; c0 += 32) {
        uint32_t v[32];
        ptx::tmem_ld_x32(tmem + ((uint32_t)(32 * warp) << 16) + c0, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint32_t off = img_off(r, c0 + 8 * g);
          float x[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) x[j] = __uint_as_float(v[8 * g + j]);
          if (hrow) {
            const uint4 hv = *reinterpret_cast<const uint4*>(hrow + off);
            const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const __half2 hh = *reinterpret_cast<const __half2*>(&hw[j]);
              if (!(__low2float(hh) > 0.f)) x[2 * j] = 0.f;
              if (!(__high2float(hh) > 0.f)) x[2 * j + 1] = 0.f;
            }
          }
          *reinterpret_cast<uint4*>(orow + off) = make_uint4(ptx::cvt_f16x2(x[0], x[1]), ptx::cvt_f16x2(x[2], x[3]),
                                                              ptx::cvt_f16x2(x[4], x[5]), ptx::cvt_f16x2(x[6], x[7]));
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_e);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 256);
}

}  // namespace nb
