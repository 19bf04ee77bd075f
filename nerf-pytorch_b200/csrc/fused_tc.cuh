// fused_tc.cuh -- the production path: one persistent, warp-specialised sm_100a kernel per network
// pass that computes  pts = o + d*z  ->  positional encoding  ->  8x256 MLP (+ heads)  ->
// alpha compositing, with the MLP as tcgen05.mma tiles (fp16 operands, fp32 accumulate in TMEM).
//
// Replaces run_nerf.py:381-386 / :397-403 (pts, network_query_fn, raw2outputs), i.e. run_network
// (:37-51), batchify (:27-34), Embedder.embed (run_nerf_helpers.py:36-45) and NeRF.forward
// (:96-119).  sigma/rgb never leave the SM unless `raw` is requested.
//
// CTA = 640 threads, 1 CTA / SM, persistent over a contiguous range of whole rays:
//   warp 0      : weight producer   -- streams the pre-swizzled fp16 weight chunks (K=32 x N) from
//                                      L2 into a 3-stage ring with cp.async.bulk (TMA engine)
//   warps 1-2   : MMA issuers       -- one thread per tile slot issues tcgen05.mma (M=128, N=256|128,
//                                      K=16); warp 2 also owns the TMEM allocation
//   warp 3      : sampler           -- o + d*z and sin/cos encoding for the NEXT 256-row super-tile
//   warps 4-19  : epilogue          -- 2 tile slots x 8 warps: tcgen05.ld -> +bias -> ReLU -> fp16 ->
//                                      st.shared into the next layer's A operand (128B swizzle);
//                                      heads (alpha, rgb) on CUDA cores; warp-scan compositing
// Two 128-row tiles (slots A, B) run in lock-step on the same weight chunk, each with its own
// 128x256 fp32 accumulator (2 x 256 TMEM columns): the tensor pipe works on one tile while the
// other tile's epilogue drains its accumulator.
#pragma once
#include "common.cuh"
#include "tc_ptx.cuh"

namespace nb {

constexpr int TC_THREADS = 640;              // producer + 2 MMA issuers + 1 sampler + 16 epilogue warps
constexpr int TC_SAMPLER_THREADS = 32;       // (register allocation is per 4 warps: 20 warps x 96 regs fit)
constexpr int TC_W = 256;                 // hidden width supported by the tensor-core path
constexpr int TC_MAXD = 8;                // pts layers supported (bias table lives in smem)
constexpr int TC_TILE = 128;              // rows per MMA tile
constexpr int TC_ST = 256;                // rows per super-tile (two tile slots)
constexpr int TC_NST = 3;                 // weight ring stages
constexpr int TC_CHUNK_K = 32;            // K per weight chunk
constexpr uint32_t TC_STAGE_BYTES = 256 * TC_CHUNK_K * 2;   // 16 KB

// shared-memory map (bytes, relative to a 1024-aligned base)
constexpr uint32_t SM_ACT = 0;                                  // 2 x 64 KB  A operand (4 K-blocks x 16 KB)
constexpr uint32_t SM_ENC = 131072;                             // 2 x 16 KB  encoded inputs (1 K-block)
constexpr uint32_t SM_WRING = 163840;                           // 3 x 16 KB  weight ring
constexpr uint32_t SM_ONES = 212992;                            // 256 B: ONE 8-row atom of the bias-selector A slab (K=16, SWIZZLE_32B,
                                                                //        SBO = 0: all 128 rows read the same atom)
constexpr uint32_t SM_BIASB = SM_ONES + 512;                    // 213504: 8 KB resident B operand [256 x K=16]: all layers' biases
                                                                //        (one 256-byte selector atom per tile slot before it)
constexpr uint32_t SM_HEADS = SM_BIASB + 8192;                  // 221440: head weights, 4128 B
constexpr uint32_t SM_PART = SM_HEADS + 4128;                   // 225568: 2 x 128 x float4 partials
constexpr uint32_t SM_BARS = SM_PART + 4096;                    // mbarriers
constexpr uint32_t SM_MISC = SM_BARS + 256;                     // 229920: tmem ptr, compositing carry
constexpr uint32_t SM_TOTAL = SM_MISC + 128;                    // 230048
constexpr uint32_t SM_ALLOC = 230400 > SM_TOTAL ? 230400 : SM_TOTAL;                           // >= SM_TOTAL and the pair kernel's P2_TOTAL; dynamic smem base is 1024-aligned (checked at run time)
static_assert(SM_TOTAL <= SM_ALLOC, "shared-memory map exceeds the allocation");
constexpr uint32_t TC_BIAS_CHUNK_BYTES = 256 * 32;              // [256 rows x K=16] fp16: column pair (2l, 2l+1) = (hi, lo) of layer l's bias

// K columns of the resident bias operand used by bias layer `li` of `nb`: (hi, lo) pairs while 16 columns
// suffice, hi only for the last 2*nb-16 layers.  Returns hi column; *lo = lo column or -1.
__host__ __device__ inline int tc_bias_cols(int li, int nb, int* lo) {
  const int n_single = (2 * nb > 16) ? 2 * nb - 16 : 0, n_pair = nb - n_single;
  if (li < n_pair) { *lo = 2 * li + 1; return 2 * li; }
  *lo = -1;
  return 2 * n_pair + (li - n_pair);
}

// heads region (floats): viewdirs: alpha_w[256] rgb_w[3][128] alpha_b rgb_b[3]; else output_w[4][256] output_b[4]
constexpr int HEADS_FLOATS = 1032;

// ---------------------------------------------------------------------------------------------
// packed weight buffer
// ---------------------------------------------------------------------------------------------
struct PackLayout {
  int D, skip, use_viewdirs, IC, ICV, NL;
  int n_chunks;
  size_t off_chunks, chunk_bytes, off_bias, off_heads, off_vdir, off_biasb, off_pair, total;
};

__host__ __device__ inline int tc_layer_chunks(int l, int D, int skip) {
  if (l == 0) return 2;
  if (l < D && skip >= 0 && l == skip + 1) return 10;
  return 8;
}
__host__ __device__ inline uint32_t tc_layer_chunk_bytes(int l, int D) { return (l == D + 1) ? TC_STAGE_BYTES / 2 : TC_STAGE_BYTES; }
// every layer but the view layer (whose bias is per ray) starts with a bias chunk: one K=16 MMA of a
// constant-one A slab against [bias_hi, bias_lo, 0...] initialises the accumulator with the bias
__host__ __device__ inline bool tc_layer_has_bias(int l, int D) { return l != D + 1; }

static inline PackLayout make_pack_layout(const NerfNetParams& n) {
  PackLayout L;
  L.D = n.D; L.skip = n.skip; L.use_viewdirs = n.use_viewdirs; L.IC = n.input_ch; L.ICV = n.input_ch_views;
  L.NL = n.D + (n.use_viewdirs ? 2 : 0);
  L.n_chunks = 0; L.chunk_bytes = 0;
  for (int l = 0; l < L.NL; ++l) {
    int c = tc_layer_chunks(l, n.D, n.skip);
    L.n_chunks += c;
    L.chunk_bytes += (size_t)c * tc_layer_chunk_bytes(l, n.D);
  }
  L.off_chunks = 1024;
  L.off_bias = L.off_chunks + L.chunk_bytes;
  L.off_heads = L.off_bias + (size_t)(TC_MAXD + 1) * 1024;
  L.off_vdir = L.off_heads + HEADS_FLOATS * 4;
  L.off_biasb = L.off_vdir + (size_t)(128 * (n.input_ch_views > 0 ? n.input_ch_views : 1) + 128) * 4;
  L.off_biasb = (L.off_biasb + 255) & ~(size_t)255;
  L.off_pair = L.off_biasb + TC_BIAS_CHUNK_BYTES;          // rank-split copy of the chunk stream for the CTA-pair kernel
  L.total = L.off_pair + L.chunk_bytes;
  return L;
}

struct PackChunk { const float* src; int ld, k0, kvalid, nrows; unsigned dst_off; };
constexpr int PACK_MAX_CHUNKS = 96;
struct PackJob { PackChunk c[PACK_MAX_CHUNKS]; int n; };

// one thread per 16-byte unit (8 fp16 of one row) of the swizzled chunk image
__global__ void pack_chunks_kernel(PackJob job, uint8_t* __restrict__ dst) {
  const int ci = blockIdx.y;
  if (ci >= job.n) return;
  const PackChunk c = job.c[ci];
  int u = blockIdx.x * blockDim.x + threadIdx.x;       // unit index: row * 4 + c16
  if (u >= c.nrows * 4) return;
  int row = u >> 2, c16 = u & 3;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    int k = c16 * 8 + j;
    v[j] = (k < c.kvalid) ? c.src[(size_t)row * c.ld + c.k0 + k] : 0.0f;
  }
  uint4 o;
  o.x = ptx::cvt_f16x2(v[0], v[1]); o.y = ptx::cvt_f16x2(v[2], v[3]);
  o.z = ptx::cvt_f16x2(v[4], v[5]); o.w = ptx::cvt_f16x2(v[6], v[7]);
  // SWIZZLE_64B K-major: 8-row atoms of 512 B, 16B-chunk index ^= (row%8)>>1
  unsigned off = c.dst_off + (row >> 3) * 512 + (row & 7) * 64 + ((c16 ^ ((row & 7) >> 1)) << 4);
  *reinterpret_cast<uint4*>(dst + off) = o;
}

// resident bias operand: [256 rows x 16 K] fp16, K-major SWIZZLE_32B (rows of 32 B, 8-row atoms of 256 B)
struct PackBiasJob { const float* src[TC_MAXD + 2]; int n; unsigned dst_off; };
__global__ void pack_bias_kernel(PackBiasJob job, uint8_t* __restrict__ dst) {
  const int row = threadIdx.x;                         // 256 threads = 256 output rows
  __half h[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) h[k] = __float2half_rn(0.0f);
  for (int li = 0; li < job.n; ++li) {
    const float b = job.src[li][row];
    int lo;
    const int hi = tc_bias_cols(li, job.n, &lo);
    const __half bh = __float2half_rn(b);
    h[hi] = bh;
    if (lo >= 0) h[lo] = __float2half_rn(b - __half2float(bh));
  }
  uint8_t* base = dst + job.dst_off + (row >> 3) * 256 + (row & 7) * 32;
  const int sw = ((row & 7) >> 2) & 1;                 // Swizzle<1,4,3>: 16-byte chunk index ^= bit 7 of the byte address
  uint4 c[2];
  for (int q = 0; q < 2; ++q) {
    const unsigned short* u = reinterpret_cast<const unsigned short*>(h + 8 * q);
    c[q] = make_uint4(u[0] | (u[1] << 16), u[2] | (u[3] << 16), u[4] | (u[5] << 16), u[6] | (u[7] << 16));
  }
  *reinterpret_cast<uint4*>(base + ((0 ^ sw) << 4)) = c[0];
  *reinterpret_cast<uint4*>(base + ((1 ^ sw) << 4)) = c[1];
}

struct PackTables { NerfNetParams net; size_t off_bias, off_heads, off_vdir; };
__global__ void pack_tables_kernel(PackTables t, uint8_t* __restrict__ dst) {
  const NerfNetParams& n = t.net;
  float* bias = reinterpret_cast<float*>(dst + t.off_bias);
  float* heads = reinterpret_cast<float*>(dst + t.off_heads);
  float* vdir = reinterpret_cast<float*>(dst + t.off_vdir);
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
  for (int i = tid; i < (TC_MAXD + 1) * 256; i += nt) {
    int l = i >> 8, c = i & 255;
    float v = 0.0f;
    if (l < n.D) v = n.pts_b[l][c];
    else if (l == n.D && n.use_viewdirs) v = n.feature_b[c];
    bias[i] = v;
  }
  if (n.use_viewdirs) {
    for (int i = tid; i < 256; i += nt) heads[i] = n.alpha_w[i];
    for (int i = tid; i < 384; i += nt) heads[256 + i] = n.rgb_w[i];
    if (tid == 0) { heads[640] = n.alpha_b[0]; heads[641] = n.rgb_b[0]; heads[642] = n.rgb_b[1]; heads[643] = n.rgb_b[2]; }
    const int ICV = n.input_ch_views;
    for (int i = tid; i < 128 * ICV; i += nt) { int j = i / ICV, c = i % ICV; vdir[i] = n.views_w[(size_t)j * (256 + ICV) + 256 + c]; }
    for (int i = tid; i < 128; i += nt) vdir[128 * ICV + i] = n.views_b[i];
  } else {
    for (int i = tid; i < 1024; i += nt) heads[i] = n.output_w[i];        // first 4 rows of output_linear
    for (int i = tid; i < 4; i += nt) heads[1024 + i] = n.output_b[i];
  }
}

// view-direction contribution of views_linears[0], once per ray (the reference recomputes it per
// sample through the expand at run_nerf.py:44-46):  vb[n][j] = b[j] + sum_c W[j][256+c] * enc(v_n)[c]
constexpr int VB_RAYS = 8;        // rays per block (512 blocks for a 4096-ray chunk: the kernel is latency-bound)
__global__ void __launch_bounds__(128) view_bias_kernel(const float* __restrict__ dirs, int dir_stride, long long N, int Lv, int ICV,
                                                     const float* __restrict__ vdir, float* __restrict__ vb) {
  __shared__ float s_enc[VB_RAYS][64];
  const long long n0 = (long long)blockIdx.x * VB_RAYS;
  const int j = threadIdx.x;
  for (int i = threadIdx.x; i < VB_RAYS * ICV; i += 128) {
    const int r = i / ICV, c = i - r * ICV;
    float v = 0.0f;
    if (n0 + r < N) {
      const float* d = dirs + (n0 + r) * dir_stride;
      if (c < 3) v = d[c];
      else { int f = (c - 3) / 6, q = (c - 3) % 6; float a = __fmul_rn(d[q % 3], exp2f((float)f)); v = (q < 3) ? sinf(a) : cosf(a); }
    }
    s_enc[r][c] = v;
  }
  float w[64];
#pragma unroll
  for (int c = 0; c < 64; ++c) w[c] = (c < ICV) ? vdir[j * ICV + c] : 0.0f;
  const float b = vdir[128 * ICV + j];
  __syncthreads();
  for (int r = 0; r < VB_RAYS && n0 + r < N; ++r) {
    float acc = b;
#pragma unroll
    for (int c = 0; c < 64; ++c) if (c < ICV) acc = fmaf(w[c], s_enc[r][c], acc);
    vb[(n0 + r) * 128 + j] = acc;
  }
}

// ---------------------------------------------------------------------------------------------
// fused march kernel
// ---------------------------------------------------------------------------------------------
struct MarchParams {
  const float* rays; int ray_stride;      // [N, ray_stride]  (o,d,near,far[,viewdir]) or NULL in pts mode
  const float* z_vals;                    // [N,S]
  const float* pts;                       // [N*S,3] (pts mode: run_network) or NULL
  const float* noise;                     // [N,S] or NULL
  const float* vb;                        // [N,128] view bias (use_viewdirs)
  long long N; int S; int rays_per_cta;
  const uint8_t* chunks; const float* bias; const float* heads; const uint8_t* biasb;
  int D, skip, use_viewdirs, L, IC;
  int white_bkgd, do_composite;
  NerfPassOut out;
  unsigned long long pair_half_bytes;     // pair kernel: bytes of one rank's half of the chunk stream
  long long* trace;                       // debug: clock64 timestamps of CTA 0, super-tile 1 (or NULL)
  int dbg;                                // debug: NERF_B200_DBG bit flags for A/B experiments (0 in production)
};

// compositing carry of the ray that is still open at a warp boundary (shared memory, 32 B)
constexpr uint32_t CARRY_T = 0, CARRY_R = 4, CARRY_G = 8, CARRY_B = 12, CARRY_D = 16, CARRY_A = 20, CARRY_TURN = 24;

__device__ __forceinline__ uint32_t act_row_off(int r) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128); }

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds32(uint32_t addr) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr)); return v; }
__device__ __forceinline__ void sts32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_shared(uint32_t addr) {
  uint32_t v; asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); return v;
}
__device__ __forceinline__ void st_release_shared(uint32_t addr, uint32_t v) {
  asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
// (a, b) += (c, d) as one packed FADD2
__device__ __forceinline__ void add2(float& a, float& b, float c, float d) {
  asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%0, %1};\n\tmov.b64 rb, {%2, %3};\n\tadd.rn.f32x2 rc, ra, rb;\n\tmov.b64 {%0, %1}, rc;\n\t}"
      : "+f"(a), "+f"(b) : "f"(c), "f"(d));
}

// v[32] (fp32 accumulators of one row, bias already added) -> (ReLU) -> fp16 -> 4 x 16-byte stores
// into the 128B-swizzled K-major A tile.
template <bool RELU>
__device__ __forceinline__ void store_act32(const float (&x)[32], uint32_t act_base, int r, int col0) {
  const uint32_t kb = (uint32_t)(col0 >> 6) * 16384u;
  const int c16_0 = (col0 & 63) >> 3;
  const uint32_t row = act_base + kb + act_row_off(r);
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint32_t h0, h1, h2, h3;
    if (RELU) {
      h0 = ptx::cvt_relu_f16x2(x[g * 8 + 0], x[g * 8 + 1]); h1 = ptx::cvt_relu_f16x2(x[g * 8 + 2], x[g * 8 + 3]);
      h2 = ptx::cvt_relu_f16x2(x[g * 8 + 4], x[g * 8 + 5]); h3 = ptx::cvt_relu_f16x2(x[g * 8 + 6], x[g * 8 + 7]);
    } else {
      h0 = ptx::cvt_f16x2(x[g * 8 + 0], x[g * 8 + 1]); h1 = ptx::cvt_f16x2(x[g * 8 + 2], x[g * 8 + 3]);
      h2 = ptx::cvt_f16x2(x[g * 8 + 4], x[g * 8 + 5]); h3 = ptx::cvt_f16x2(x[g * 8 + 6], x[g * 8 + 7]);
    }
    ptx::st_shared_v4(row + (uint32_t)(((c16_0 + g) ^ (r & 7)) << 4), h0, h1, h2, h3);
  }
}

// Same, with the 8 swizzled 16-byte-chunk addresses of this thread's row precomputed once
// (sw[c] = tile base + row offset + ((c ^ (r & 7)) << 4)): every store is [register + immediate].
template <bool RELU, int COL0>
__device__ __forceinline__ void store_act32_pre(const float (&x)[32], const uint32_t (&sw)[8]) {
  constexpr uint32_t kb = (uint32_t)(COL0 >> 6) * 16384u;
  constexpr int c16_0 = (COL0 & 63) >> 3;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint32_t h0, h1, h2, h3;
    if (RELU) {
      h0 = ptx::cvt_relu_f16x2(x[g * 8 + 0], x[g * 8 + 1]); h1 = ptx::cvt_relu_f16x2(x[g * 8 + 2], x[g * 8 + 3]);
      h2 = ptx::cvt_relu_f16x2(x[g * 8 + 4], x[g * 8 + 5]); h3 = ptx::cvt_relu_f16x2(x[g * 8 + 6], x[g * 8 + 7]);
    } else {
      h0 = ptx::cvt_f16x2(x[g * 8 + 0], x[g * 8 + 1]); h1 = ptx::cvt_f16x2(x[g * 8 + 2], x[g * 8 + 3]);
      h2 = ptx::cvt_f16x2(x[g * 8 + 4], x[g * 8 + 5]); h3 = ptx::cvt_f16x2(x[g * 8 + 6], x[g * 8 + 7]);
    }
    ptx::st_shared_v4(sw[c16_0 + g] + kb, h0, h1, h2, h3);
  }
}

// segmented (per-ray) inclusive scans over one warp; `s` = lane of the segment start at or before
// this lane inside the warp, or -1 when the segment began in an earlier warp.
__device__ __forceinline__ float seg_scan_mul(float v, int lane, int s) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { float t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o && lane - o >= s) v *= t; }
  return v;
}
__device__ __forceinline__ float seg_scan_add(float v, int lane, int s) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { float t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o && lane - o >= s) v += t; }
  return v;
}

// One 32-column batch of a hidden-layer epilogue: the bias is already in the accumulator (bias chunk)
__device__ __forceinline__ void as_float32(const uint32_t (&v)[32], float (&x)[32]) {
#pragma unroll
  for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
}

// Row `r` (0..7) of the 8-row bias-selector A atom for bias layer `li`: 1.0 in that layer's K columns.
__device__ __forceinline__ void write_bias_selector(uint32_t slab, int r, int li, int nb) {
  int lo;
  const int hi = tc_bias_cols(li, nb, &lo);
  uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};            // 16 halves as 8 words
  w[hi >> 1] |= (hi & 1) ? 0x3c000000u : 0x00003c00u;
  if (lo >= 0) w[lo >> 1] |= (lo & 1) ? 0x3c000000u : 0x00003c00u;
  const int sw = (r >> 2) & 1;
  const uint32_t a = slab + (uint32_t)(r * 32);
  ptx::st_shared_v4(a + ((0 ^ sw) << 4), w[0], w[1], w[2], w[3]);
  ptx::st_shared_v4(a + ((1 ^ sw) << 4), w[4], w[5], w[6], w[7]);
}

// Positional encoding of one point into its 128-byte K-block row of the A operand (64 fp16 channels:
// x y z, then per frequency sin(2^f xyz), cos(2^f xyz); channel 63 is zero padding), written as eight
// 16-byte swizzled chunks.  run_nerf_helpers.py:36-45 evaluates sin/cos of 2^f x for every f; here only
// f = 0 and f = 5 are evaluated with sincosf, the other octaves come from the double-angle identities
// sin 2a = 2 sin a cos a, cos 2a = (cos a - sin a)(cos a + sin a).  Four doublings amplify the fp32
// rounding of the anchor to <= 1.6e-6 absolute (numpy model over [-6, 6]: 0.07 % of the fp16 operand
// values move by one fp16 ulp), far below the fp16 quantisation of the operand itself, and cut the
// sampler warp's work ~4x: with 30 sincosf per row it took longer than the MMAs of half a super-tile and
// the issuers idled on enc_full (profiles/r01_summary.md, "sampler").
__device__ __forceinline__ void encode_row_store(uint32_t row, int tr_, float px, float py, float pz, int L) {
  uint32_t h[32];
  float s[3], c[3];
  const float q[3] = {px, py, pz};
  float carry = 0.f;                                   // pending even-indexed channel of the current pair
  int n = 0;                                           // channels emitted so far (compile-time after unrolling)
  auto emit = [&](float v) {
    if (n & 1) { __half2 hh = __floats2half2_rn(carry, v); h[n >> 1] = *reinterpret_cast<uint32_t*>(&hh); }
    else carry = v;
    ++n;
  };
  emit(px); emit(py); emit(pz);
#pragma unroll
  for (int f = 0; f < 10; ++f) {
    if (f == 0 || f == 5) {
#pragma unroll
      for (int j = 0; j < 3; ++j) sincosf(q[j] * (float)(1 << f), &s[j], &c[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float s2 = (s[j] + s[j]) * c[j], c2 = (c[j] - s[j]) * (c[j] + s[j]);
        s[j] = s2; c[j] = c2;
      }
    }
    const bool on = f < L;
#pragma unroll
    for (int j = 0; j < 3; ++j) emit(on ? s[j] : 0.f);
#pragma unroll
    for (int j = 0; j < 3; ++j) emit(on ? c[j] : 0.f);
  }
  emit(0.f);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const uint32_t a = row + (uint32_t)((k ^ (tr_ & 7)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(h[4 * k]), "r"(h[4 * k + 1]), "r"(h[4 * k + 2]), "r"(h[4 * k + 3]) : "memory");
  }
}

// raw (rgb, sigma) of one row per lane -> alpha compositing of whole rays (run_nerf.py:275-303) with the carry of the
// ray that is open at the warp boundary threaded warp -> warp through shared memory (ticket order: st, slot, q).
__device__ __forceinline__ void composite_rows(const MarchParams& p, const float4 raw4, const bool valid, const int lr, const int rl,
                                               const long long n_ray, const long long row_begin, const int st, const int X,
                                               const int q, const int lane, const uint32_t a_carry) {
  const long long m = row_begin + lr;
  if (valid && p.out.raw) reinterpret_cast<float4*>(p.out.raw)[m] = raw4;
  if (!p.do_composite) return;
  const int k = lr - rl * p.S;
  float alpha = 0.f, cr = 0.f, cg = 0.f, cb = 0.f, z = 0.f;
  if (valid) {
    const float* rd = p.rays + n_ray * p.ray_stride + 3;
    const float norm = sqrtf(rd[0] * rd[0] + rd[1] * rd[1] + rd[2] * rd[2]);         // run_nerf.py:280
    z = p.z_vals[m];
    float dist = (k == p.S - 1) ? 1e10f : __fsub_rn(p.z_vals[m + 1], z);             // :277-278
    dist = __fmul_rn(dist, norm);
    const float sg = raw4.w + (p.noise ? p.noise[m] : 0.0f);
    alpha = __fsub_rn(1.0f, expf(-fmaxf(sg, 0.0f) * dist));                           // :275
    cr = sigmoidf_acc(raw4.x); cg = sigmoidf_acc(raw4.y); cb = sigmoidf_acc(raw4.z); // :282
  }
  const bool seg_start = valid && (k == 0), seg_end = valid && (k == p.S - 1);
  const unsigned smask = __ballot_sync(0xffffffffu, seg_start);
  const unsigned below = smask & ((lane == 31) ? 0xffffffffu : ((2u << lane) - 1u));
  const int s = below ? (31 - __clz(below)) : -1;
  const float qv = valid ? __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f) : 1.0f;         // :295
  const float pv = seg_scan_mul(qv, lane, s);
  float ev = __shfl_up_sync(0xffffffffu, pv, 1);
  if (lane == 0 || s == lane) ev = 1.0f;
  const float wl = valid ? alpha * ev : 0.0f;
  float t_r = seg_scan_add(wl * cr, lane, s), t_g = seg_scan_add(wl * cg, lane, s), t_b = seg_scan_add(wl * cb, lane, s);
  float t_d = seg_scan_add(wl * z, lane, s), t_a = seg_scan_add(wl, lane, s);
  const uint32_t ticket = (uint32_t)((st * 2 + X) * 4 + q);
  if (lane == 0) { while (ld_acquire_shared(a_carry + CARRY_TURN) != ticket) { } }
  __syncwarp();
  const float Tin = lds32(a_carry + CARRY_T);
  const float c_r = lds32(a_carry + CARRY_R), c_g = lds32(a_carry + CARRY_G), c_b = lds32(a_carry + CARRY_B);
  const float c_d = lds32(a_carry + CARRY_D), c_a = lds32(a_carry + CARRY_A);
  __syncwarp();
  if (s < 0) { t_r = fmaf(Tin, t_r, c_r); t_g = fmaf(Tin, t_g, c_g); t_b = fmaf(Tin, t_b, c_b); t_d = fmaf(Tin, t_d, c_d); t_a = fmaf(Tin, t_a, c_a); }
  if (lane == 31) {
    if (seg_end) {
      sts32(a_carry + CARRY_T, 1.0f); sts32(a_carry + CARRY_R, 0.f); sts32(a_carry + CARRY_G, 0.f); sts32(a_carry + CARRY_B, 0.f);
      sts32(a_carry + CARRY_D, 0.f); sts32(a_carry + CARRY_A, 0.f);
    } else {
      sts32(a_carry + CARRY_T, (s >= 0) ? pv : Tin * pv);
      sts32(a_carry + CARRY_R, t_r); sts32(a_carry + CARRY_G, t_g); sts32(a_carry + CARRY_B, t_b);
      sts32(a_carry + CARRY_D, t_d); sts32(a_carry + CARRY_A, t_a);
    }
    st_release_shared(a_carry + CARRY_TURN, ticket + 1u);
  }
  if (valid && p.out.weights) p.out.weights[m] = (s < 0) ? Tin * wl : wl;
  if (seg_end) {
    float rr = t_r, gg = t_g, bb = t_b;
    if (p.white_bkgd) { const float bg = 1.0f - t_a; rr += bg; gg += bg; bb += bg; }  // :302-303
    if (p.out.rgb_map) { p.out.rgb_map[n_ray * 3] = rr; p.out.rgb_map[n_ray * 3 + 1] = gg; p.out.rgb_map[n_ray * 3 + 2] = bb; }
    if (p.out.disp_map) {
      const float ratio = t_d / t_a;
      const float mm = (ratio != ratio) ? ratio : fmaxf(1e-10f, ratio);              // :299
      p.out.disp_map[n_ray] = 1.0f / mm;
    }
    if (p.out.acc_map) p.out.acc_map[n_ray] = t_a;
    if (p.out.depth_map) p.out.depth_map[n_ray] = t_d;
  }
}

extern __shared__ __align__(1024) uint8_t tc_smem[];

__global__ void __launch_bounds__(TC_THREADS, 1) march_tc_kernel(const MarchParams p) {
  uint8_t* smem = tc_smem;
  const uint32_t sb = ptx::smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if ((sb & 1023u) != 0) __trap();                    // swizzle atoms need the 1024-byte alignment

  float* s_heads = reinterpret_cast<float*>(smem + SM_HEADS);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + SM_MISC);
  const uint32_t a_heads = sb + SM_HEADS, a_part = sb + SM_PART, a_carry = sb + SM_MISC + 16;

  // mbarriers
  const uint32_t bar_wfull = sb + SM_BARS;            // [3]
  const uint32_t bar_wempty = sb + SM_BARS + 24;      // [3]
  const uint32_t bar_dfull = sb + SM_BARS + 48;       // [2]
  const uint32_t bar_act = sb + SM_BARS + 64;         // [2]
  const uint32_t bar_encfull = sb + SM_BARS + 80;
  const uint32_t bar_encfree = sb + SM_BARS + 88;

  // this CTA's rays / rows
  const long long ray0 = (long long)blockIdx.x * p.rays_per_cta;
  const long long ray1 = (ray0 + p.rays_per_cta < p.N) ? ray0 + p.rays_per_cta : p.N;
  const int nrows = (ray1 > ray0) ? (int)(ray1 - ray0) * p.S : 0;
  const int nst = (nrows + TC_ST - 1) / TC_ST;
  const long long row_begin = ray0 * p.S;
  const int D = p.D, NL = p.D + (p.use_viewdirs ? 2 : 0);
  const int last_enc_layer = (p.skip >= 0 && p.skip + 1 < D) ? p.skip + 1 : 0;

  // ---- one-time setup ----
  const int n_bias = D + (p.use_viewdirs ? 1 : 0);      // layers whose bias rides in the GEMM (all but the view layer)
  if (threadIdx.x < 16) write_bias_selector(sb + SM_ONES + (threadIdx.x >> 3) * 256, threadIdx.x & 7, 0, n_bias);
  for (int i = threadIdx.x; i < (int)(TC_BIAS_CHUNK_BYTES / 16); i += TC_THREADS)
    reinterpret_cast<uint4*>(smem + SM_BIASB)[i] = reinterpret_cast<const uint4*>(p.biasb)[i];
  ptx::fence_proxy_async_smem();
  for (int i = threadIdx.x; i < HEADS_FLOATS; i += TC_THREADS) s_heads[i] = p.heads[i];
  if (threadIdx.x == 0) {
    sts32(a_carry + CARRY_T, 1.0f); sts32(a_carry + CARRY_R, 0.f); sts32(a_carry + CARRY_G, 0.f); sts32(a_carry + CARRY_B, 0.f);
    sts32(a_carry + CARRY_D, 0.f); sts32(a_carry + CARRY_A, 0.f); st_release_shared(a_carry + CARRY_TURN, 0u);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < TC_NST; ++i) { ptx::mbar_init(bar_wfull + 8 * i, 1); ptx::mbar_init(bar_wempty + 8 * i, 2); }
    for (int x = 0; x < 2; ++x) { ptx::mbar_init(bar_dfull + 8 * x, 1); ptx::mbar_init(bar_act + 8 * x, 256); }
    ptx::mbar_init(bar_encfull, TC_SAMPLER_THREADS);
    ptx::mbar_init(bar_encfree, 2);
    ptx::fence_mbar_init();
  }
  if (warp == 2) { ptx::tmem_alloc(ptx::smem_u32(s_tmem), 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *s_tmem;

  if (warp == 0) {
    // =========================== weight producer ===========================
    // (warp-uniform control flow; one elected lane issues the bulk copies)
    {
      uint32_t stage = 0, ph = 0;
      for (int st = 0; st < nst; ++st) {
        const uint8_t* src = p.chunks;
        for (int l = 0; l < NL; ++l) {
          const int nch = tc_layer_chunks(l, D, p.skip);
          const uint32_t cb = tc_layer_chunk_bytes(l, D);
          for (int c = 0; c < nch; ++c) {
            ptx::mbar_wait(bar_wempty + 8 * stage, ph ^ 1);
            if (ptx::elect_one()) {
              ptx::mbar_arrive_expect_tx(bar_wfull + 8 * stage, cb);
              ptx::bulk_g2s(sb + SM_WRING + stage * TC_STAGE_BYTES, src, cb, bar_wfull + 8 * stage);
            }
            __syncwarp();
            src += cb;
            if (++stage == TC_NST) { stage = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1 || warp == 2) {
    // =========================== MMA issuers (one per tile slot) ===========================
    // A warp cannot overlap its own bookkeeping (mbarrier waits, commits) with its tcgen05.mma issue
    // (measured: issue time is additive), so each tile slot has its own issuing warp: while one warp
    // waits / commits, the other warp's MMAs keep the tensor pipe busy.  Accumulation order inside a
    // slot is preserved because one thread issues all MMAs of that slot's accumulator.
    // Control flow is warp-uniform (all 32 lanes run the loops and the mbarrier waits); only the tcgen05
    // instructions are issued by one elected lane.  That keeps descriptors and addresses in uniform
    // registers -- issued from a divergent `if (lane == 0)` region every UTCHMMA is wrapped in an
    // ELECT / R2UR.BROADCAST waterfall loop that costs ~100 cycles per MMA (profiles/r01 issue probe).
    {
      const int X = warp - 1;
      uint32_t stage = 0, ph = 0, actph = 0;
      bool ready = false;                                // w_full of the current chunk already observed
      const uint64_t adesc0 = ptx::umma_desc(sb, 1024, ptx::UMMA_SW128);
      const uint64_t bdesc0 = ptx::umma_desc(sb, 512, ptx::UMMA_SW64);
      const uint64_t sel_desc = ptx::umma_desc(sb + SM_ONES + X * 256, 0, ptx::UMMA_SW32);   // SBO = 0: one atom for all rows
      const uint64_t bias_desc = ptx::umma_desc(sb + SM_BIASB, 256, ptx::UMMA_SW32);
      const uint32_t d_tmem = __shfl_sync(0xffffffffu, tmem, 0) + X * 256;
      for (int st = 0; st < nst; ++st) {
        const bool tr = p.trace && blockIdx.x == 0 && st == 1 && X == 0 && lane == 0;
        const bool trs = p.trace && blockIdx.x == 0 && X == 0 && lane == 0 && st < 48;
        if (trs) p.trace[2200 + 2 * st] = clock64();
        ptx::mbar_wait(bar_encfull, st & 1);
        if (trs) p.trace[2201 + 2 * st] = clock64();
        for (int l = 0; l < NL; ++l) {
          const int nch = tc_layer_chunks(l, D, p.skip);
          const bool skip_layer = (l < D && p.skip >= 0 && l == p.skip + 1);
          const bool has_bias = tc_layer_has_bias(l, D);
          const uint32_t idesc = ptx::umma_idesc_f16(128, (l == D + 1) ? 128 : 256);
          const bool last_layer = (st == nst - 1) && (l == NL - 1);
          for (int c = 0; c < nch; ++c) {
            long long* trp = p.trace + 4 * (l * 10 + c);
            if (tr) trp[0] = clock64();
            if (!ready) ptx::mbar_wait(bar_wfull + 8 * stage, ph);
            if (tr) trp[1] = clock64();
            if (c == 0) { ptx::mbar_wait(bar_act + 8 * X, actph); actph ^= 1; if (tr) trp[2] = clock64(); }
            ptx::tc_fence_after();
            const bool is_enc = (l == 0) || (skip_layer && c < 2);
            const int kc = skip_layer ? c - 2 : c;
            const uint64_t bd = bdesc0 + ((SM_WRING + stage * TC_STAGE_BYTES) >> 4);
            const uint32_t a_off = is_enc ? (SM_ENC + X * 16384 + c * 64) : (SM_ACT + X * 65536 + (kc >> 1) * 16384 + (kc & 1) * 64);
            const uint64_t ad = adesc0 + (a_off >> 4);
            const uint32_t nstage = (stage + 1 == TC_NST) ? 0u : stage + 1, nph = (stage + 1 == TC_NST) ? ph ^ 1u : ph;
            if (ptx::elect_one()) {
              // the layer's first MMA initialises the accumulator with the bias: D = selector(1.0 in this
              // layer's K columns) x resident bias operand
              if (c == 0 && has_bias) ptx::mma_f16_ss(d_tmem, sel_desc, bias_desc, idesc, 0u);
              ptx::mma_f16_ss(d_tmem, ad, bd, idesc, (c > 0 || has_bias) ? 1u : 0u);
              ptx::mma_f16_ss(d_tmem, ad + 2, bd + 2, idesc, 1u);
              if (c == nch - 1) ptx::mma_commit(bar_dfull + 8 * X);
              ptx::mma_commit(bar_wempty + 8 * stage);
              if (is_enc && c == 1 && l == last_enc_layer) ptx::mma_commit(bar_encfree);
            }
            __syncwarp();
            // probe the NEXT chunk's weights (usually already there: the ring runs two chunks ahead)
            ready = !(last_layer && c == nch - 1) && ptx::mbar_try_wait(bar_wfull + 8 * nstage, nph);
            ready = __all_sync(0xffffffffu, ready);
            stage = nstage; ph = nph;
            if (tr) trp[3] = clock64();
          }
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // =========================== epilogue ===========================
    // TMEM lane quadrant is fixed by (warp id % 4)
    const int X = (warp - 4) >> 3, e = (warp - 4) & 7, q = warp & 3, ch = e >> 2;
    const int r = 32 * q + lane;                                  // tile row == TMEM lane
    const uint32_t t_lane = tmem + ((uint32_t)(32 * q) << 16) + X * 256;
    const uint32_t act_base = sb + SM_ACT + X * 65536;
    ptx::mbar_arrive(bar_act + 8 * X);                            // accumulator initially free
    uint32_t dph = 0;
    // swizzled 16-byte-chunk addresses of this thread's row in the two K-blocks of its column half
    uint32_t swk[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) swk[c] = act_base + (uint32_t)(ch * 2) * 16384u + act_row_off(r) + (uint32_t)((c ^ (r & 7)) << 4);
    for (int st = 0; st < nst; ++st) {
      float hp0 = 0.f, hp1 = 0.f, hp2 = 0.f, hp3 = 0.f;           // head partial sums of this thread's columns
      const int lr = st * TC_ST + X * TC_TILE + r;                // row index inside this CTA's range
      const bool valid = lr < nrows;
      const int rl = (valid ? lr : nrows - 1) / p.S;              // local ray
      const long long n_ray = ray0 + rl;
      for (int l = 0; l < NL; ++l) {
        const bool tr = p.trace && blockIdx.x == 0 && st == 1 && e == 0 && lane == 0;
        long long* trp = p.trace + 2048 + 4 * (X * 16 + l);
        if (tr) trp[0] = clock64();
        ptx::mbar_wait(bar_dfull + 8 * X, dph);
        dph ^= 1;
        ptx::tc_fence_after();
        if (tr) trp[1] = clock64();
        if (l <= D) {
          // pts layer (ReLU) or feature layer (no activation): 128 columns per warp in 4 batches,
          // the TMEM load of batch b+1 in flight while batch b is converted and stored
          const bool last_pts = (l == D - 1);
          const bool write_act = !(last_pts && !p.use_viewdirs);
          const int colw = ch * 128;
          uint32_t va[32], vb[32];
          ptx::tmem_ld_x32(t_lane + colw, va);
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const int col0 = colw + b * 32;
            uint32_t (&v)[32] = (b & 1) ? vb : va;
            uint32_t (&vn)[32] = (b & 1) ? va : vb;
            ptx::tmem_ld_wait();
            if (b < 3) ptx::tmem_ld_x32(t_lane + col0 + 32, vn);
            float x[32];
            as_float32(v, x);
            if (last_pts) {
              if (p.use_viewdirs) {                               // alpha_linear (run_nerf_helpers.py:106)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float4 w = lds128(a_heads + (uint32_t)(col0 + 4 * j) * 4u);
                  hp3 = fmaf(fmaxf(x[4 * j + 0], 0.f), w.x, hp3); hp3 = fmaf(fmaxf(x[4 * j + 1], 0.f), w.y, hp3);
                  hp3 = fmaf(fmaxf(x[4 * j + 2], 0.f), w.z, hp3); hp3 = fmaf(fmaxf(x[4 * j + 3], 0.f), w.w, hp3);
                }
              } else {                                            // output_linear (:117)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float4 w0 = lds128(a_heads + (uint32_t)(col0 + 4 * j) * 4u), w1 = lds128(a_heads + (uint32_t)(256 + col0 + 4 * j) * 4u);
                  const float4 w2 = lds128(a_heads + (uint32_t)(512 + col0 + 4 * j) * 4u), w3 = lds128(a_heads + (uint32_t)(768 + col0 + 4 * j) * 4u);
                  const float h0 = fmaxf(x[4 * j + 0], 0.f), h1 = fmaxf(x[4 * j + 1], 0.f), h2 = fmaxf(x[4 * j + 2], 0.f), h3 = fmaxf(x[4 * j + 3], 0.f);
                  hp0 = fmaf(h0, w0.x, hp0); hp0 = fmaf(h1, w0.y, hp0); hp0 = fmaf(h2, w0.z, hp0); hp0 = fmaf(h3, w0.w, hp0);
                  hp1 = fmaf(h0, w1.x, hp1); hp1 = fmaf(h1, w1.y, hp1); hp1 = fmaf(h2, w1.z, hp1); hp1 = fmaf(h3, w1.w, hp1);
                  hp2 = fmaf(h0, w2.x, hp2); hp2 = fmaf(h1, w2.y, hp2); hp2 = fmaf(h2, w2.z, hp2); hp2 = fmaf(h3, w2.w, hp2);
                  hp3 = fmaf(h0, w3.x, hp3); hp3 = fmaf(h1, w3.y, hp3); hp3 = fmaf(h2, w3.z, hp3); hp3 = fmaf(h3, w3.w, hp3);
                }
              }
            }
            if (write_act) {
              // (b & 1) selects the K-block inside the column half; ch selects the half: the
              // immediate part of the address is compile-time, the row/swizzle part is in sw[]
              if (l < D) {
                if (b == 0) store_act32_pre<true, 0>(x, swk); else if (b == 1) store_act32_pre<true, 32>(x, swk);
                else if (b == 2) store_act32_pre<true, 64>(x, swk); else store_act32_pre<true, 96>(x, swk);
              } else {
                if (b == 0) store_act32_pre<false, 0>(x, swk); else if (b == 1) store_act32_pre<false, 32>(x, swk);
                else if (b == 2) store_act32_pre<false, 64>(x, swk); else store_act32_pre<false, 96>(x, swk);
              }
            }
          }
          // this slot's bias selector for the NEXT layer
          {
            const int nxt = (l + 1 < n_bias) ? l + 1 : ((l == NL - 1) ? 0 : -1);
            if (q == 0 && ch == 0 && lane < 8 && nxt >= 0) write_bias_selector(sb + SM_ONES + X * 256, lane, nxt, n_bias);
          }
          ptx::tc_fence_before();
          ptx::fence_proxy_async_smem();
          ptx::mbar_arrive(bar_act + 8 * X);
          if (tr) trp[2] = clock64();
        } else {
          // views_linears[0] (N=128): 64 columns per warp; + per-ray view bias, ReLU, rgb_linear
          const float* vbrow = p.vb + n_ray * 128;
          uint32_t va[32], vb[32];
          ptx::tmem_ld_x32(t_lane + ch * 64, va);
          ptx::tmem_ld_x32(t_lane + ch * 64 + 32, vb);
          ptx::tmem_ld_wait();
          if (q == 0 && ch == 0 && lane < 8) write_bias_selector(sb + SM_ONES + X * 256, lane, 0, n_bias);   // next super-tile, layer 0
          ptx::tc_fence_before();
          ptx::fence_proxy_async_smem();
          ptx::mbar_arrive(bar_act + 8 * X);                      // accumulator drained: next super-tile may start
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            const int col0 = ch * 64 + b * 32;
            const uint32_t (&v)[32] = b ? vb : va;
            const float4* vb4 = reinterpret_cast<const float4*>(vbrow + col0);
            const uint32_t w0 = a_heads + (uint32_t)(256 + col0) * 4u;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 bb = vb4[j];
              const float h0 = fmaxf(__uint_as_float(v[4 * j + 0]) + bb.x, 0.f), h1 = fmaxf(__uint_as_float(v[4 * j + 1]) + bb.y, 0.f);
              const float h2 = fmaxf(__uint_as_float(v[4 * j + 2]) + bb.z, 0.f), h3 = fmaxf(__uint_as_float(v[4 * j + 3]) + bb.w, 0.f);
              const float4 wr = lds128(w0 + 16 * j), wg = lds128(w0 + 512 + 16 * j), wb = lds128(w0 + 1024 + 16 * j);
              hp0 = fmaf(h0, wr.x, hp0); hp0 = fmaf(h1, wr.y, hp0); hp0 = fmaf(h2, wr.z, hp0); hp0 = fmaf(h3, wr.w, hp0);
              hp1 = fmaf(h0, wg.x, hp1); hp1 = fmaf(h1, wg.y, hp1); hp1 = fmaf(h2, wg.z, hp1); hp1 = fmaf(h3, wg.w, hp1);
              hp2 = fmaf(h0, wb.x, hp2); hp2 = fmaf(h1, wb.y, hp2); hp2 = fmaf(h2, wb.z, hp2); hp2 = fmaf(h3, wb.w, hp2);
            }
          }
        }
      }
      // ---- heads: combine the two column halves, then raw -> compositing (ch == 0 warps) ----
      const uint32_t part = a_part + (uint32_t)(X * 128 + r) * 16u;
      if (ch == 1) sts128(part, make_float4(hp0, hp1, hp2, hp3));
      ptx::named_bar_sync(1 + X, 256);
      if (ch == 0) {
        const float4 o = lds128(part);
        float4 raw4;
        if (p.use_viewdirs) raw4 = make_float4(hp0 + o.x + lds32(a_heads + 641 * 4), hp1 + o.y + lds32(a_heads + 642 * 4),
                                               hp2 + o.z + lds32(a_heads + 643 * 4), hp3 + o.w + lds32(a_heads + 640 * 4));
        else raw4 = make_float4(hp0 + o.x + lds32(a_heads + 1024 * 4), hp1 + o.y + lds32(a_heads + 1025 * 4),
                                hp2 + o.z + lds32(a_heads + 1026 * 4), hp3 + o.w + lds32(a_heads + 1027 * 4));
        const long long m = row_begin + lr;
        if (valid && p.out.raw) reinterpret_cast<float4*>(p.out.raw)[m] = raw4;
        if (p.do_composite) {
          const int k = lr - rl * p.S;
          float alpha = 0.f, cr = 0.f, cg = 0.f, cb = 0.f, z = 0.f;
          if (valid) {
            const float* rd = p.rays + n_ray * p.ray_stride + 3;
            const float norm = sqrtf(rd[0] * rd[0] + rd[1] * rd[1] + rd[2] * rd[2]);         // run_nerf.py:280
            z = p.z_vals[m];
            float dist = (k == p.S - 1) ? 1e10f : __fsub_rn(p.z_vals[m + 1], z);             // :277-278
            dist = __fmul_rn(dist, norm);
            const float sg = raw4.w + (p.noise ? p.noise[m] : 0.0f);
            alpha = __fsub_rn(1.0f, expf(-fmaxf(sg, 0.0f) * dist));                           // :275
            cr = sigmoidf_acc(raw4.x); cg = sigmoidf_acc(raw4.y); cb = sigmoidf_acc(raw4.z); // :282
          }
          // warp-local part (independent of the carry): transmittance / weights relative to
          // max(ray start, warp start) and their segmented sums
          const bool seg_start = valid && (k == 0), seg_end = valid && (k == p.S - 1);
          const unsigned smask = __ballot_sync(0xffffffffu, seg_start);
          const unsigned below = smask & ((lane == 31) ? 0xffffffffu : ((2u << lane) - 1u));
          const int s = below ? (31 - __clz(below)) : -1;
          const float qv = valid ? __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f) : 1.0f;         // :295
          const float pv = seg_scan_mul(qv, lane, s);
          float ev = __shfl_up_sync(0xffffffffu, pv, 1);
          if (lane == 0 || s == lane) ev = 1.0f;
          const float wl = valid ? alpha * ev : 0.0f;
          float t_r = seg_scan_add(wl * cr, lane, s), t_g = seg_scan_add(wl * cg, lane, s), t_b = seg_scan_add(wl * cb, lane, s);
          float t_d = seg_scan_add(wl * z, lane, s), t_a = seg_scan_add(wl, lane, s);
          // take the compositing turn: rows are consumed in order across warps / slots / super-tiles;
          // only the few flops that thread the carry through this warp sit on the serial chain
          const uint32_t ticket = (uint32_t)((st * 2 + X) * 4 + q);
          if (lane == 0) { while (ld_acquire_shared(a_carry + CARRY_TURN) != ticket) { } }
          __syncwarp();
          const float Tin = lds32(a_carry + CARRY_T);
          const float c_r = lds32(a_carry + CARRY_R), c_g = lds32(a_carry + CARRY_G), c_b = lds32(a_carry + CARRY_B);
          const float c_d = lds32(a_carry + CARRY_D), c_a = lds32(a_carry + CARRY_A);
          __syncwarp();
          if (s < 0) { t_r = fmaf(Tin, t_r, c_r); t_g = fmaf(Tin, t_g, c_g); t_b = fmaf(Tin, t_b, c_b); t_d = fmaf(Tin, t_d, c_d); t_a = fmaf(Tin, t_a, c_a); }
          if (lane == 31) {
            if (seg_end) {
              sts32(a_carry + CARRY_T, 1.0f); sts32(a_carry + CARRY_R, 0.f); sts32(a_carry + CARRY_G, 0.f); sts32(a_carry + CARRY_B, 0.f);
              sts32(a_carry + CARRY_D, 0.f); sts32(a_carry + CARRY_A, 0.f);
            } else {
              sts32(a_carry + CARRY_T, (s >= 0) ? pv : Tin * pv);
              sts32(a_carry + CARRY_R, t_r); sts32(a_carry + CARRY_G, t_g); sts32(a_carry + CARRY_B, t_b);
              sts32(a_carry + CARRY_D, t_d); sts32(a_carry + CARRY_A, t_a);
            }
            st_release_shared(a_carry + CARRY_TURN, ticket + 1u);
          }
          // off the chain: weights and per-ray outputs
          if (valid && p.out.weights) p.out.weights[m] = (s < 0) ? Tin * wl : wl;
          if (seg_end) {
            float rr = t_r, gg = t_g, bb = t_b;
            if (p.white_bkgd) { const float bg = 1.0f - t_a; rr += bg; gg += bg; bb += bg; }  // :302-303
            if (p.out.rgb_map) { p.out.rgb_map[n_ray * 3] = rr; p.out.rgb_map[n_ray * 3 + 1] = gg; p.out.rgb_map[n_ray * 3 + 2] = bb; }
            if (p.out.disp_map) {
              const float ratio = t_d / t_a;
              const float mm = (ratio != ratio) ? ratio : fmaxf(1e-10f, ratio);              // :299
              p.out.disp_map[n_ray] = 1.0f / mm;
            }
            if (p.out.acc_map) p.out.acc_map[n_ray] = t_a;
            if (p.out.depth_map) p.out.depth_map[n_ray] = t_d;
          }
        }
      }
    }
  } else {
    // =========================== sampler (warp 3) ===========================
    const int t = threadIdx.x - 96;                               // 0..31
    for (int st = 0; st < nst; ++st) {
      const bool trs = p.trace && blockIdx.x == 0 && t == 0 && st < 48;
      if (trs) p.trace[2300 + 2 * st] = clock64();
      ptx::mbar_wait(bar_encfree, (st & 1) ^ 1);
      if (trs) p.trace[2301 + 2 * st] = clock64();
#pragma unroll 1
      for (int i = 0; i < 8; ++i) {
        const int X = i >> 2, tr_ = t + 32 * (i & 3);             // tile slot, tile row
        const int lr = st * TC_ST + X * TC_TILE + tr_;
        float px = 0.f, py = 0.f, pz = 0.f;
        if (lr < nrows) {
          const long long m = row_begin + lr;
          if (p.pts) { px = p.pts[m * 3]; py = p.pts[m * 3 + 1]; pz = p.pts[m * 3 + 2]; }
          else {
            const float* ry = p.rays + (ray0 + lr / p.S) * p.ray_stride;
            const float z = p.z_vals[m];
            px = __fadd_rn(ry[0], __fmul_rn(ry[3], z));                                       // run_nerf.py:381
            py = __fadd_rn(ry[1], __fmul_rn(ry[4], z));
            pz = __fadd_rn(ry[2], __fmul_rn(ry[5], z));
          }
        }
        encode_row_store(sb + SM_ENC + X * 16384 + act_row_off(tr_), tr_, px, py, pz, p.L);
      }
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(bar_encfull);
      if (trs) p.trace[2400 + st] = clock64();
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------
// self-test GEMM: out[128,N] = A[128,K] * W[N,K]^T through exactly the operand layouts, descriptors,
// bulk copies and TMEM loads the march kernel uses (fp16 operands, fp32 accumulate).  1 CTA.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) selftest_gemm_kernel(const float* __restrict__ A, const uint8_t* __restrict__ chunks,
                                                             int K, int N, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sb = ptx::smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, r = threadIdx.x;
  const uint32_t ACT = 0, WST = 65536, BAR = 65536 + 16384, TPTR = BAR + 64;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + TPTR);
  if (threadIdx.x == 0) { ptx::mbar_init(sb + BAR, 1); ptx::mbar_init(sb + BAR + 8, 1); ptx::fence_mbar_init(); }
  if (warp == 0) { ptx::tmem_alloc(ptx::smem_u32(s_tmem), 256); ptx::tmem_relinquish(); }
  // A -> fp16, 128B-swizzled K-major (thread r owns row r), K padded to a multiple of 32 by the caller
  for (int c0 = 0; c0 < K; c0 += 32) {
    float x[32];
    for (int j = 0; j < 32; ++j) x[j] = A[(size_t)r * K + c0 + j];
    store_act32<false>(x, sb + ACT, r, c0);
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *s_tmem;
  if (threadIdx.x == 0) {
    const uint32_t cb = (uint32_t)N * 64;
    const uint32_t idesc = ptx::umma_idesc_f16(128, N);
    for (int c = 0; c < K / 32; ++c) {
      ptx::mbar_arrive_expect_tx(sb + BAR, cb);
      ptx::bulk_g2s(sb + WST, chunks + (size_t)c * cb, cb, sb + BAR);
      ptx::mbar_wait(sb + BAR, c & 1);
      ptx::tc_fence_after();
      const uint32_t a_base = sb + ACT + (c >> 1) * 16384 + (c & 1) * 64;
      for (int j = 0; j < 2; ++j)
        ptx::mma_f16_ss(tmem, ptx::umma_desc(a_base + j * 32, 1024, ptx::UMMA_SW128),
                        ptx::umma_desc(sb + WST + j * 32, 512, ptx::UMMA_SW64), idesc, (c > 0 || j > 0) ? 1u : 0u);
      ptx::mma_commit(sb + BAR + 8);
      ptx::mbar_wait(sb + BAR + 8, c & 1);       // serialise: the single weight stage is reused
    }
  }
  __syncthreads();
  ptx::tc_fence_after();
  for (int col0 = 0; col0 < N; col0 += 32) {
    uint32_t v[32];
    ptx::tmem_ld_x32(tmem + ((uint32_t)(32 * warp) << 16) + col0, v);
    ptx::tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(size_t)r * N + col0 + j] = __uint_as_float(v[j]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 256);
}

// ---------------------------------------------------------------------------------------------
// "TN" self-test for the weight-gradient GEMM of the backward (round 2): out[256,256] = X^T Y with
// X, Y [128 rows, 256] given as the forward's activation tiles (row = sample, K-blocks of 64 columns,
// SWIZZLE_128B) and read by the MMA as MN-MAJOR operands: A = X viewed [M = column, K = row],
// B = Y viewed [N = column, K = row].  The canonical MN-major SWIZZLE_128B atom is 64 contiguous
// MN elements x 8 K rows = the same physical 1 KB atom as the K-major one, so no re-layout is needed:
// LBO = stride between 64-column groups (16 KB: the K-block stride), SBO = stride between 8-row groups
// (1 KB).  lbo / sbo are arguments so that one GPU run can confirm the encoding.  1 CTA, 128 threads.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) selftest_gemm_tn_kernel(const float* __restrict__ X, const float* __restrict__ Y,
                                                                float* __restrict__ out, uint32_t lbo, uint32_t sbo) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sb = ptx::smem_u32(smem);
  const int warp = threadIdx.x >> 5, r = threadIdx.x;
  const uint32_t XT = 0, YT = 65536, BAR = 131072, TPTR = BAR + 64;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + TPTR);
  if (threadIdx.x == 0) { ptx::mbar_init(sb + BAR, 1); ptx::fence_mbar_init(); }
  if (warp == 0) { ptx::tmem_alloc(ptx::smem_u32(s_tmem), 512); ptx::tmem_relinquish(); }
  for (int c0 = 0; c0 < 256; c0 += 32) {
    float x[32], y[32];
    for (int j = 0; j < 32; ++j) { x[j] = X[(size_t)r * 256 + c0 + j]; y[j] = Y[(size_t)r * 256 + c0 + j]; }
    store_act32<false>(x, sb + XT, r, c0);
    store_act32<false>(y, sb + YT, r, c0);
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *s_tmem;
  if (threadIdx.x == 0) {
    const uint32_t idesc = ptx::umma_idesc_f16_major(128, 256, 1, 1);
    for (int mh = 0; mh < 2; ++mh)                      // output rows (X columns) 0-127, 128-255
      for (int k = 0; k < 8; ++k) {                     // 16 sample rows per MMA: two 8-row groups = 2 KB
        const uint64_t ad = ptx::umma_desc_full(sb + XT + mh * 2 * 16384 + k * 2048, lbo, sbo, ptx::UMMA_SW128);
        const uint64_t bd = ptx::umma_desc_full(sb + YT + k * 2048, lbo, sbo, ptx::UMMA_SW128);
        ptx::mma_f16_ss(tmem + mh * 256, ad, bd, idesc, k > 0 ? 1u : 0u);
      }
    ptx::mma_commit(sb + BAR);
  }
  ptx::mbar_wait(sb + BAR, 0);
  ptx::tc_fence_after();
  for (int mh = 0; mh < 2; ++mh)
    for (int col0 = 0; col0 < 256; col0 += 32) {
      uint32_t v[32];
      ptx::tmem_ld_x32(tmem + ((uint32_t)(32 * warp) << 16) + mh * 256 + col0, v);
      ptx::tmem_ld_wait();
      for (int j = 0; j < 32; ++j) out[(size_t)(mh * 128 + r) * 256 + col0 + j] = __uint_as_float(v[j]);
    }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 512);
}

// MMA issue-rate microbenchmark: `reps` x (M=128, N, K=16) tcgen05.mma on resident (garbage) operands,
// alternating between two accumulators; out[0] = cycles from first issue to completion of the last.
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int reps, int N, int b_sw64, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sb = ptx::smem_u32(smem);
  const int warp = threadIdx.x >> 5;
  const uint32_t BAR = 65536 + 32768, TPTR = BAR + 64;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + TPTR);
  for (int i = threadIdx.x; i < (65536 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 1.0
  if (threadIdx.x == 0) { ptx::mbar_init(sb + BAR, 1); ptx::fence_mbar_init(); }
  if (warp == 0) { ptx::tmem_alloc(ptx::smem_u32(s_tmem), 512); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *s_tmem;
  if (threadIdx.x == 0) {
    const uint32_t idesc = ptx::umma_idesc_f16(128, N);
    const long long t0 = clock64();
    for (int i = 0; i < reps; ++i) {
      const uint32_t a = sb + ((i >> 1) & 3) * 16384 + (i & 1) * 32;
      const uint64_t bd = b_sw64 ? ptx::umma_desc(sb + 65536 + (i & 1) * 32, 512, ptx::UMMA_SW64)
                                 : ptx::umma_desc(sb + 65536 + (i & 3) * 32, 1024, ptx::UMMA_SW128);
      ptx::mma_f16_ss(tmem + (i & 1) * 256, ptx::umma_desc(a, 1024, ptx::UMMA_SW128), bd, idesc, 1u);
    }
    const long long t1 = clock64();
    ptx::mma_commit(sb + BAR);
    ptx::mbar_wait(sb + BAR, 0);
    const long long t2 = clock64();
    out[0] = t2 - t0; out[1] = t1 - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 512);
}

// Epilogue-rate microbenchmark: 8 warps drain a 128x256 fp32 accumulator `reps` times the way the
// march kernel does (mode 0: tcgen05.ld only; 1: + bias/ReLU/convert; 2: + st.shared of the A tile),
// optionally while another warp keeps the tensor pipe busy on the other accumulator (mma != 0).
// out[0] = cycles for `reps` tile-layer epilogues (warp 4 lane 0).
__global__ void __launch_bounds__(384, 1) epi_rate_kernel(int reps, int mode, int mma, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sb = ptx::smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t ACT = 0, OPS = 65536, BIAS = 65536 + 49152, BAR = BIAS + 1024, TPTR = BAR + 64;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + TPTR);
  float* s_bias = reinterpret_cast<float*>(smem + BIAS);
  volatile int* s_stop = reinterpret_cast<volatile int*>(smem + TPTR + 16);
  for (int i = threadIdx.x; i < (65536 + 49152) / 4; i += 384) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x < 256) s_bias[threadIdx.x] = 0.5f;
  if (threadIdx.x == 0) { ptx::mbar_init(sb + BAR, 1); ptx::fence_mbar_init(); *s_stop = 0; }
  if (warp == 0) { ptx::tmem_alloc(ptx::smem_u32(s_tmem), 512); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *s_tmem;
  if (warp == 1) {
    if (lane == 0 && mma) {
      const uint32_t idesc = ptx::umma_idesc_f16(128, 256);
      int i = 0;
      while (!*s_stop) {
        for (int j = 0; j < 8; ++j, ++i)
          ptx::mma_f16_ss(tmem + 256, ptx::umma_desc(sb + OPS + (i & 1) * 32, 1024, ptx::UMMA_SW128),
                          ptx::umma_desc(sb + OPS + 16384 + (i & 1) * 32, 512, ptx::UMMA_SW64), idesc, 1u);
        ptx::mma_commit(sb + BAR);
        ptx::mbar_wait(sb + BAR, (i / 8 - 1) & 1);
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int e = warp - 4, q = e & 3, ch = e >> 2, r = 32 * q + lane;
    const uint32_t t_lane = tmem + ((uint32_t)(32 * q) << 16);
    float sink = 0.f;
    const int math = mode & 3;
    const bool no_pfence = mode & 4, no_tfence = mode & 8, no_bar = mode & 16, dual = mode & 32;
    long long ldw = 0;
    ptx::named_bar_sync(1, 256);
    const long long t0 = clock64();
    for (int it = 0; it < reps; ++it) {
      if (dual) {
#pragma unroll 1
        for (int b = 0; b < 2; ++b) {
          const int col0 = ch * 128 + b * 64;
          uint32_t v[32], w[32];
          const long long c0 = clock64();
          ptx::tmem_ld_x32(t_lane + col0, v);
          ptx::tmem_ld_x32(t_lane + col0 + 32, w);
          ptx::tmem_ld_wait();
          ldw += clock64() - c0;
          sink += __uint_as_float(v[it & 31]) + __uint_as_float(w[it & 31]);
        }
      } else {
#pragma unroll 1
        for (int b = 0; b < 4; ++b) {
          const int col0 = ch * 128 + b * 32;
          uint32_t v[32];
          const long long c0 = clock64();
          ptx::tmem_ld_x32(t_lane + col0, v);
          ptx::tmem_ld_wait();
          ldw += clock64() - c0;
          if (math == 0) { sink += __uint_as_float(v[it & 31]); continue; }
          float x[32];
          const float4* b4 = reinterpret_cast<const float4*>(s_bias + col0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 bb = b4[j];
            x[4 * j + 0] = __uint_as_float(v[4 * j + 0]) + bb.x; x[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + bb.y;
            x[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + bb.z; x[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + bb.w;
          }
          if (math == 1) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) sink += __uint_as_float(ptx::cvt_relu_f16x2(x[j], x[j + 1]));
          } else store_act32<true>(x, sb + ACT, r, col0);
        }
      }
      if (!no_tfence) ptx::tc_fence_before();
      if (!no_pfence) ptx::fence_proxy_async_smem();
      if (!no_bar) ptx::named_bar_sync(1, 256);
    }
    const long long t1 = clock64();
    if (e == 0 && lane == 0) { out[0] = t1 - t0; out[1] = ldw; *s_stop = 1; }
    if (sink == 123.456f) out[1] = 1;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 512);
}

// L2 -> shared-memory streaming probe: every CTA streams the same `buf_bytes` buffer `passes` times through a ring of
// `stages` x `chunk` bytes with cp.async.bulk (a consumer warp frees a stage as soon as it lands).
// out[blockIdx.x*2] = cycles, out[blockIdx.x*2+1] = summed issue->landed latency of warp 0's copies.
__global__ void __launch_bounds__(64, 1) l2_stream_probe_kernel(const uint8_t* __restrict__ buf, int buf_bytes, int chunk, int stages,
                                                              int passes, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sb = ptx::smem_u32(smem);
  const uint32_t BAR = (uint32_t)stages * (uint32_t)chunk;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { for (int i = 0; i < stages; ++i) { ptx::mbar_init(sb + BAR + 8 * i, 1); ptx::mbar_init(sb + BAR + 128 + 8 * i, 1); } ptx::fence_mbar_init(); }
  __syncthreads();
  const int n = (buf_bytes / chunk) * passes;
  const long long t0 = clock64();
  if (warp == 0) {
    uint32_t stage = 0, ph = 0;
    int off = 0;
    for (int i = 0; i < n; ++i) {
      ptx::mbar_wait(sb + BAR + 128 + 8 * stage, ph ^ 1);
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(sb + BAR + 8 * stage, chunk);
        ptx::bulk_g2s(sb + stage * chunk, buf + off, chunk, sb + BAR + 8 * stage);
      }
      __syncwarp();
      off += chunk; if (off + chunk > buf_bytes) off = 0;
      if (++stage == (uint32_t)stages) { stage = 0; ph ^= 1; }
    }
  } else {
    uint32_t stage = 0, ph = 0;
    for (int i = 0; i < n; ++i) {
      ptx::mbar_wait(sb + BAR + 8 * stage, ph);
      if (ptx::elect_one()) ptx::mbar_arrive(sb + BAR + 128 + 8 * stage);
      __syncwarp();
      if (++stage == (uint32_t)stages) { stage = 0; ph ^= 1; }
    }
    if (threadIdx.x == 32) out[blockIdx.x] = clock64() - t0;
  }
}

// Issue-overhead probe: one thread runs `reps` iterations of {optional mbarrier try_wait on a completed
// barrier; `nmma` x tcgen05.mma (N=256, K=16); optional tcgen05.commit to a scratch barrier} and reports the
// cycles per iteration (out[0] = to completion of everything, out[1] = issue loop only).
//   flags bit0: try_wait per iteration   bit1: one commit per iteration   bit2: two commits   bit3: tcgen05.fence::after
__global__ void __launch_bounds__(128, 1) issue_probe_kernel(int reps, int nmma, int flags, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sb = ptx::smem_u32(smem);
  const int warp = threadIdx.x >> 5;
  const uint32_t BAR = 65536 + 32768, TPTR = BAR + 64;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + TPTR);
  for (int i = threadIdx.x; i < (65536 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { ptx::mbar_init(sb + BAR, 1); ptx::mbar_init(sb + BAR + 8, 1); ptx::mbar_init(sb + BAR + 16, 1); ptx::fence_mbar_init(); }
  if (warp == 0) { ptx::tmem_alloc(ptx::smem_u32(s_tmem), 512); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *s_tmem;
  if (warp == 0 && (flags & 128)) {
    // warp-uniform control flow, only the tcgen05 instructions are issued by one elected lane
    const uint32_t idesc = ptx::umma_idesc_f16(128, 256);
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    if (threadIdx.x == 0) ptx::mbar_arrive(sb + BAR + 16);
    __syncwarp();
    const uint64_t ad = ptx::umma_desc(sb, 1024, ptx::UMMA_SW128), bd = ptx::umma_desc(sb + 65536, 512, ptx::UMMA_SW64);
    const long long t0 = clock64();
    for (int i = 0; i < reps; ++i) {
      if (flags & 1) ptx::mbar_wait(sb + BAR + 16, 0);
      if (flags & 8) ptx::tc_fence_after();
      if (ptx::elect_one()) {
        for (int j = 0; j < nmma; ++j) {
          const uint32_t dsel = (flags & 16) ? 0u : ((flags & 32) ? (uint32_t)(j & 1) : (uint32_t)(i & 1));
          const uint32_t ks = (flags & 64) ? (uint32_t)((i + j) & 3) : (uint32_t)(j & 1);
          ptx::mma_f16_ss(tm + dsel * 256, ad + 2 * ks, bd + 2 * (ks & 1), idesc, 1u);
        }
        if (flags & 2) ptx::mma_commit(sb + BAR + 8);
        if (flags & 4) ptx::mma_commit(sb + BAR + 8);
      }
      __syncwarp();
    }
    const long long t1 = clock64();
    if (ptx::elect_one()) ptx::mma_commit(sb + BAR);
    __syncwarp();
    ptx::mbar_wait(sb + BAR, 0);
    const long long t2 = clock64();
    if (threadIdx.x == 0) { out[0] = t2 - t0; out[1] = t1 - t0; }
  } else if (threadIdx.x == 0 && !(flags & 128)) {
    const uint32_t idesc = ptx::umma_idesc_f16(128, 256);
    ptx::mbar_arrive(sb + BAR + 16);                    // complete phase 0 of the "always ready" barrier
    const uint64_t ad = ptx::umma_desc(sb, 1024, ptx::UMMA_SW128), bd = ptx::umma_desc(sb + 65536, 512, ptx::UMMA_SW64);
    const long long t0 = clock64();
    for (int i = 0; i < reps; ++i) {
      if (flags & 1) ptx::mbar_wait(sb + BAR + 16, 0);
      if (flags & 8) ptx::tc_fence_after();
      // flags bit4: keep ONE accumulator (no alternation)   bit5: alternate the accumulator every MMA   bit6: vary operand k-step per iteration
      for (int j = 0; j < nmma; ++j) {
        const uint32_t dsel = (flags & 16) ? 0u : ((flags & 32) ? (uint32_t)(j & 1) : (uint32_t)(i & 1));
        const uint32_t ks = (flags & 64) ? (uint32_t)((i + j) & 3) : (uint32_t)(j & 1);
        ptx::mma_f16_ss(tmem + dsel * 256, ad + 2 * ks, bd + 2 * (ks & 1), idesc, 1u);
      }
      if (flags & 2) ptx::mma_commit(sb + BAR + 8);
      if (flags & 4) ptx::mma_commit(sb + BAR + 8);
    }
    const long long t1 = clock64();
    ptx::mma_commit(sb + BAR);
    ptx::mbar_wait(sb + BAR, 0);
    const long long t2 = clock64();
    out[0] = t2 - t0; out[1] = t1 - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 512);
}

// TMEM -> register load-rate probe.  `nwarps` warps (4 or 8) each drain their share of a 128-lane x 256-column
// fp32 accumulator `reps` times with the given tcgen05.ld shape; values are xor-folded so nothing spills.
//   shape 0: 32x32b.x32 (one load in flight)   1: 32x32b.x32, two loads per wait   2: 32x32b.x64
//   shape 3: 16x256b.x8 (two per 32-lane group) 4: 16x128b.x16                     5: 32x32b.x16, 4 per wait
#define NB_LDTM(SHAPE, NREG, ...) asm volatile("tcgen05.ld.sync.aligned." SHAPE ".b32 {" __VA_ARGS__ "}, [%" #NREG "];"
__device__ __forceinline__ uint32_t fold32(const uint32_t (&v)[32]) { uint32_t a = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) a ^= v[i]; return a; }
__device__ __forceinline__ void ldtm_16x256b_x8(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ldtm_16x128b_x16(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.16x128b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ldtm_32x32b_x16(uint32_t taddr, uint32_t (&v)[32], int o) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]), "=r"(v[o + 6]), "=r"(v[o + 7]),
        "=r"(v[o + 8]), "=r"(v[o + 9]), "=r"(v[o + 10]), "=r"(v[o + 11]), "=r"(v[o + 12]), "=r"(v[o + 13]), "=r"(v[o + 14]), "=r"(v[o + 15])
      : "r"(taddr) : "memory");
}

__global__ void __launch_bounds__(384, 1) ldtm_rate_kernel(int reps, int shape, int nwarps, int mma, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sb = ptx::smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t OPS = 0, BAR = 49152, TPTR = BAR + 64;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + TPTR);
  volatile int* s_stop = reinterpret_cast<volatile int*>(smem + TPTR + 16);
  for (int i = threadIdx.x; i < 49152 / 4; i += 384) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { ptx::mbar_init(sb + BAR, 1); ptx::fence_mbar_init(); *s_stop = 0; }
  if (warp == 0) { ptx::tmem_alloc(ptx::smem_u32(s_tmem), 512); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *s_tmem;
  if (warp == 1) {
    if (lane == 0 && mma) {
      const uint32_t idesc = ptx::umma_idesc_f16(128, 256);
      int i = 0;
      while (!*s_stop) {
        for (int j = 0; j < 8; ++j, ++i)
          ptx::mma_f16_ss(tmem + 256, ptx::umma_desc(sb + OPS + (i & 1) * 32, 1024, ptx::UMMA_SW128),
                          ptx::umma_desc(sb + OPS + 16384 + (i & 1) * 32, 512, ptx::UMMA_SW64), idesc, 1u);
        ptx::mma_commit(sb + BAR);
        ptx::mbar_wait(sb + BAR, (i / 8 - 1) & 1);
      }
    }
    __syncwarp();
  } else if (warp >= 4 && warp < 4 + nwarps) {
    const int e = warp - 4, q = warp & 3, ch = e >> 2;
    const int ncol = (nwarps == 8) ? 128 : 256, col_base = (nwarps == 8) ? ch * 128 : 0;
    const uint32_t t_lane = tmem + ((uint32_t)(32 * q) << 16) + col_base;
    uint32_t acc = 0;
    ptx::named_bar_sync(1, nwarps * 32);
    const long long t0 = clock64();
    for (int it = 0; it < reps; ++it) {
      if (shape == 0) {
#pragma unroll 1
        for (int c = 0; c < ncol; c += 32) { uint32_t v[32]; ptx::tmem_ld_x32(t_lane + c, v); ptx::tmem_ld_wait(); acc ^= fold32(v); }
      } else if (shape == 1) {
#pragma unroll 1
        for (int c = 0; c < ncol; c += 64) { uint32_t v[32], w[32]; ptx::tmem_ld_x32(t_lane + c, v); ptx::tmem_ld_x32(t_lane + c + 32, w); ptx::tmem_ld_wait(); acc ^= fold32(v) ^ fold32(w); }
      } else if (shape == 3) {
#pragma unroll 1
        for (int c = 0; c < ncol; c += 64) { uint32_t v[32], w[32]; ldtm_16x256b_x8(t_lane + c, v); ldtm_16x256b_x8(t_lane + (16u << 16) + c, w); ptx::tmem_ld_wait(); acc ^= fold32(v) ^ fold32(w); }
      } else if (shape == 4) {
#pragma unroll 1
        for (int c = 0; c < ncol; c += 64) { uint32_t v[32], w[32]; ldtm_16x128b_x16(t_lane + c, v); ldtm_16x128b_x16(t_lane + (16u << 16) + c, w); ptx::tmem_ld_wait(); acc ^= fold32(v) ^ fold32(w); }
      } else if (shape == 5) {
#pragma unroll 1
        for (int c = 0; c < ncol; c += 64) { uint32_t v[32], w[32]; ldtm_32x32b_x16(t_lane + c, v, 0); ldtm_32x32b_x16(t_lane + c + 16, v, 16); ldtm_32x32b_x16(t_lane + c + 32, w, 0); ldtm_32x32b_x16(t_lane + c + 48, w, 16); ptx::tmem_ld_wait(); acc ^= fold32(v) ^ fold32(w); }
      }
    }
    const long long t1 = clock64();
    ptx::named_bar_sync(1, nwarps * 32);
    if (e == 0 && lane == 0) { out[0] = t1 - t0; *s_stop = 1; }
    if (acc == 0x12345u) out[1] = acc;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 512);
}

}  // namespace nb
