// capi.cu -- extern "C" entry points of libnerf_b200.so (see include/nerf_b200.h).
#include <stdarg.h>
#include "common.cuh"
#include "small_kernels.cuh"
#include "bwd_simt.cuh"
#include "mlp_simt.cuh"
#include "fused_tc.cuh"
#include "fused_tc2.cuh"
#include "bwd_tc.cuh"
#include <stdlib.h>

namespace nb {
thread_local char g_err[512] = {0};
long long g_launches = 0;
int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// ---- optional device-time accounting of march_tc_kernel launches (bench.py roofline) ----
struct TimedLaunch { cudaEvent_t a, b; double flops; };
static bool g_timing = false;
static TimedLaunch g_timed[4096];
static int g_ntimed = 0;

static double net_macs_per_row(const NerfNetParams& n) {
  // MACs of the reference's Linear layers per sample row
  double m = (double)n.input_ch * n.W;
  for (int i = 1; i < n.D; ++i) m += (double)((n.skip >= 0 && i == n.skip + 1) ? n.W + n.input_ch : n.W) * n.W;
  if (n.use_viewdirs) m += (double)n.W * n.W + n.W + (double)(n.W + n.input_ch_views) * (n.W / 2) + 3.0 * (n.W / 2);
  else m += (double)n.W * n.output_ch;
  return m;
}

static long long* g_trace = nullptr;      // debug: device buffer of 4096 int64 clock stamps (or NULL)

static int check_tc_net(const NerfNetParams* net) {
  NB_CHECK_ARG(net != nullptr, "net is NULL");
  NB_CHECK_ARG(net->W == TC_W, "tensor-core path supports netwidth == 256 (got %d); use precision=FP32", net->W);
  NB_CHECK_ARG(net->D >= 2 && net->D <= TC_MAXD, "tensor-core path supports 2 <= netdepth <= %d (got %d)", TC_MAXD, net->D);
  NB_CHECK_ARG(net->input_ch >= 3 && net->input_ch <= 63, "input_ch must be in [3,63] (got %d)", net->input_ch);
  NB_CHECK_ARG(net->skip < net->D - 1, "skip layer must be < D-1");
  if (net->use_viewdirs) NB_CHECK_ARG(net->input_ch_views >= 3 && net->input_ch_views <= 63, "input_ch_views must be in [3,63]");
  return 0;
}

static int smem_optin(const void* fn, size_t bytes) {
  NB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}

static int num_sms() {
  static int n = 0;
  if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); if (n <= 0) n = 148; }
  return n;
}

// shared launcher of the fused tcgen05 pass
static int launch_march(const float* rays, int ray_stride, const float* z_vals, const float* pts,
                        const float* dirs, int dir_stride, const float* noise, long long N, int S,
                        const NerfNetParams* net, const void* packed, int L, int Lv, int white_bkgd, int do_composite,
                        const NerfPassOut* out, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  if (int rc = check_tc_net(net)) return rc;
  NB_CHECK_ARG(packed != nullptr, "packed weights are NULL (call nerf_b200_pack_weights)");
  NB_CHECK_ARG(N > 0 && S > 0, "empty ray batch must be handled by the caller (N=%lld S=%d)", N, S);
  NB_CHECK_ARG((long long)N * S < (1ll << 40), "batch too large");
  NB_CHECK_ARG(3 + 6 * L == net->input_ch || (L == 0 && net->input_ch == 3), "multires %d does not match input_ch %d", L, net->input_ch);
  const PackLayout PL = make_pack_layout(*net);
  const uint8_t* pk = static_cast<const uint8_t*>(packed);
  MarchParams p;
  memset(&p, 0, sizeof(p));
  p.rays = rays; p.ray_stride = ray_stride; p.z_vals = z_vals; p.pts = pts; p.noise = noise;
  p.N = N; p.S = S;
  p.chunks = pk + PL.off_chunks; p.bias = reinterpret_cast<const float*>(pk + PL.off_bias);
  p.heads = reinterpret_cast<const float*>(pk + PL.off_heads);
  p.biasb = pk + PL.off_biasb;
  p.D = net->D; p.skip = net->skip; p.use_viewdirs = net->use_viewdirs; p.L = L; p.IC = net->input_ch;
  p.white_bkgd = white_bkgd; p.do_composite = do_composite;
  if (out) p.out = *out;
  if (net->use_viewdirs) {
    NB_CHECK_ARG(dirs != nullptr, "viewdirs required by a use_viewdirs network");
    NB_CHECK_ARG(workspace != nullptr && workspace_bytes >= (size_t)N * 128 * 4, "workspace too small: need %zu bytes", (size_t)N * 128 * 4);
    NB_CHECK_ARG(3 + 6 * Lv == net->input_ch_views || (Lv == 0 && net->input_ch_views == 3), "multires_views mismatch");
    float* vb = static_cast<float*>(workspace);
    view_bias_kernel<<<cdiv(N, VB_RAYS), 128, 0, st>>>(dirs, dir_stride, N, Lv, net->input_ch_views,
                                                  reinterpret_cast<const float*>(pk + PL.off_vdir), vb);
    NB_LAUNCH_OK("view_bias_kernel");
    p.vb = vb;
  }
  // persistent grid: whole rays per CTA, balanced over the SMs
  const int sms = num_sms();
  long long rows = N * (long long)S;
  long long want = (rows + TC_ST - 1) / TC_ST;
  int grid = (int)(want < sms ? want : sms);
  if (grid > N) grid = (int)N;
  if (grid < 1) grid = 1;
  p.rays_per_cta = (int)((N + grid - 1) / grid);
  grid = (int)((N + p.rays_per_cta - 1) / p.rays_per_cta);
  NB_CHECK_ARG((long long)p.rays_per_cta * S < (1ll << 30), "rays_per_cta * S overflows");
  // default: the CTA-pair kernel (fused_tc2.cuh, cta_group::2); NERF_B200_PAIR=0 selects the single-CTA kernel
  // (fused_tc.cuh) that it superseded -- kept for A/B measurements, latched at the first launch.
  static int pair_mode = -1;
  if (pair_mode < 0) { const char* e = getenv("NERF_B200_PAIR"); pair_mode = (e && e[0] == '0') ? 0 : 1; }
  static bool optin = false;
  if (!optin) {
    if (int rc = smem_optin((const void*)march_tc_kernel, SM_ALLOC)) return rc;
    if (int rc = smem_optin((const void*)march_tc2_kernel, SM_ALLOC)) return rc;
    optin = true;
  }
  if (pair_mode) grid = (grid + 1) & ~1;                 // whole pairs; a padding CTA owns no rays
  p.trace = g_trace;
  { static int dbg = -1; if (dbg < 0) { const char* e = getenv("NERF_B200_DBG"); dbg = e ? atoi(e) : 0; } p.dbg = dbg; }
  TimedLaunch* tl = nullptr;
  if (g_timing && g_ntimed < 4096) {
    tl = &g_timed[g_ntimed++];
    cudaEventCreate(&tl->a); cudaEventCreate(&tl->b);
    tl->flops = 2.0 * net_macs_per_row(*net) * (double)rows;
    cudaEventRecord(tl->a, st);
  }
  if (pair_mode) {
    // tensor map over the rank-split chunk stream, viewed as [rows][256] uint16 (512-byte rows)
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
      void* fn = nullptr;
      cudaDriverEntryPointQueryResult qres;
      NB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
      NB_CHECK_ARG(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
      encode = reinterpret_cast<EncodeFn>(fn);
    }
    CUtensorMap wmap;                           // box = 16 stream rows = 8 KB = one ring stage
    const cuuint64_t gdim[2] = {256, (cuuint64_t)(PL.chunk_bytes / 512)};
    const cuuint64_t gstr[1] = {512};
    const cuuint32_t box[2] = {256, 16}, estr[2] = {1, 1};
    uint8_t* pair_stream = const_cast<uint8_t*>(pk + PL.off_pair);
    p.chunks = pair_stream;                     // the pair kernel walks the rank-split stream
    p.pair_half_bytes = PL.chunk_bytes / 2;
    CUresult cr = encode(&wmap, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, pair_stream, gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    NB_CHECK_ARG(cr == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", (int)cr);
    march_tc2_kernel<<<grid, TC_THREADS, SM_ALLOC, st>>>(p, wmap);
  }
  else march_tc_kernel<<<grid, TC_THREADS, SM_ALLOC, st>>>(p);
  if (tl) cudaEventRecord(tl->b, st);
  NB_LAUNCH_OK("march_tc_kernel");
  return 0;
}

}  // namespace nb

using namespace nb;

extern "C" {

int nerf_b200_abi_version(void) { return NERF_B200_ABI_VERSION; }
const char* nerf_b200_last_error(void) { return g_err; }
int64_t nerf_b200_launch_count(void) { return g_launches; }

int nerf_b200_timing_enable(int on) { g_timing = on != 0; return 0; }
int nerf_b200_timing_read(double* kernel_ms, int64_t* launches, double* algorithmic_flops) {
  double ms = 0, fl = 0;
  for (int i = 0; i < g_ntimed; ++i) {
    NB_CUDA(cudaEventSynchronize(g_timed[i].b));
    float t = 0;
    NB_CUDA(cudaEventElapsedTime(&t, g_timed[i].a, g_timed[i].b));
    ms += t; fl += g_timed[i].flops;
    cudaEventDestroy(g_timed[i].a); cudaEventDestroy(g_timed[i].b);
  }
  if (kernel_ms) *kernel_ms = ms;
  if (launches) *launches = g_ntimed;
  if (algorithmic_flops) *algorithmic_flops = fl;
  g_ntimed = 0;
  return 0;
}

int nerf_b200_embed(const float* x, int64_t M, int L, float* out, void* stream) {
  NB_CHECK_ARG(x && out, "NULL pointer");
  NB_CHECK_ARG(L >= 0 && L <= 16, "num_freqs out of range");
  if (M == 0) return 0;
  long long total = M * (3 + 6 * L);
  embed_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(x, M, L, out);
  NB_LAUNCH_OK("embed_kernel");
  return 0;
}

size_t nerf_b200_packed_bytes(const NerfNetParams* net) {
  if (!net || check_tc_net(net)) return 0;
  return make_pack_layout(*net).total;
}

int nerf_b200_pack_weights(const NerfNetParams* net, void* packed, size_t packed_bytes, void* stream) {
  if (int rc = check_tc_net(net)) return rc;
  const PackLayout PL = make_pack_layout(*net);
  NB_CHECK_ARG(packed && packed_bytes >= PL.total, "packed buffer too small (%zu < %zu)", packed_bytes, PL.total);
  cudaStream_t st = (cudaStream_t)stream;
  PackJob job;
  job.n = 0;
  unsigned off = (unsigned)PL.off_chunks;
  const int IC = net->input_ch, W = net->W;
  auto add = [&](const float* src, int ld, int k0, int kvalid, int nrows) {
    PackChunk& c = job.c[job.n++];
    c.src = src; c.ld = ld; c.k0 = k0; c.kvalid = kvalid < 0 ? 0 : (kvalid > 32 ? 32 : kvalid); c.nrows = nrows; c.dst_off = off;
    off += (unsigned)nrows * 64;
  };
  PackBiasJob bj;
  bj.n = 0;
  for (int l = 0; l < PL.NL; ++l) {
    if (tc_layer_has_bias(l, net->D)) bj.src[bj.n++] = (l < net->D) ? net->pts_b[l] : net->feature_b;
    if (l == 0) { add(net->pts_w[0], IC, 0, IC, 256); add(net->pts_w[0], IC, 32, IC - 32, 256); }
    else if (l < net->D) {
      const bool sk = (net->skip >= 0 && l == net->skip + 1);
      const int ld = sk ? W + IC : W, base = sk ? IC : 0;
      if (sk) { add(net->pts_w[l], ld, 0, IC, 256); add(net->pts_w[l], ld, 32, IC - 32, 256); }    // cat([input_pts, h])
      for (int c = 0; c < 8; ++c) add(net->pts_w[l], ld, base + 32 * c, 32, 256);
    } else if (l == net->D) { for (int c = 0; c < 8; ++c) add(net->feature_w, W, 32 * c, 32, 256); }
    else { for (int c = 0; c < 8; ++c) add(net->views_w, W + net->input_ch_views, 32 * c, 32, 128); }
  }
  NB_CHECK_ARG(job.n == PL.n_chunks && off == PL.off_chunks + PL.chunk_bytes, "internal: chunk table mismatch");
  bj.dst_off = (unsigned)PL.off_biasb;
  pack_bias_kernel<<<1, 256, 0, st>>>(bj, static_cast<uint8_t*>(packed));
  NB_LAUNCH_OK("pack_bias_kernel");
  dim3 grid(4, job.n);
  pack_chunks_kernel<<<grid, 256, 0, st>>>(job, static_cast<uint8_t*>(packed));
  NB_LAUNCH_OK("pack_chunks_kernel");
  {
    // rank-split copy for the CTA-pair kernel: [rank 0: every chunk's rows 0..N/2-1][rank 1: rows N/2..N-1], so that
    // two consecutive chunk halves of one CTA are contiguous (one 16 KB TMA box)
    uint8_t* base = static_cast<uint8_t*>(packed);
    const size_t views_bytes = net->use_viewdirs ? (size_t)8 * (TC_STAGE_BYTES / 2) : 0;
    const size_t reg_bytes = PL.chunk_bytes - views_bytes;
    for (int r = 0; r < 2; ++r) {
      uint8_t* dst = base + PL.off_pair + (size_t)r * (PL.chunk_bytes / 2);
      NB_CUDA(cudaMemcpy2DAsync(dst, 8192, base + PL.off_chunks + (size_t)r * 8192, 16384, 8192, reg_bytes / 16384, cudaMemcpyDeviceToDevice, st));
      if (views_bytes)
        NB_CUDA(cudaMemcpy2DAsync(dst + reg_bytes / 2, 4096, base + PL.off_chunks + reg_bytes + (size_t)r * 4096, 8192, 4096, 8, cudaMemcpyDeviceToDevice, st));
    }
  }
  PackTables t;
  t.net = *net; t.off_bias = PL.off_bias; t.off_heads = PL.off_heads; t.off_vdir = PL.off_vdir;
  pack_tables_kernel<<<8, 256, 0, st>>>(t, static_cast<uint8_t*>(packed));
  NB_LAUNCH_OK("pack_tables_kernel");
  return 0;
}

int nerf_b200_run_network(const float* pts, const float* viewdirs, int64_t N, int S, const NerfNetParams* net,
                          const void* packed, int multires, int multires_views, int precision, float* raw,
                          void* workspace, size_t workspace_bytes, void* stream) {
  NB_CHECK_ARG(pts && raw && net, "NULL pointer");
  if (N == 0 || S == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (precision == NERF_B200_PREC_FP32) {
    NB_CHECK_ARG(net->W <= SIMT_THREADS && net->W % 4 == 0, "exact path supports netwidth <= 256");
    NB_CHECK_ARG(!net->use_viewdirs || viewdirs, "viewdirs required");
    const long long M = N * (long long)S;
    const size_t sm = simt_smem_bytes(*net);
    if (int rc = smem_optin((const void*)mlp_simt_kernel, sm)) return rc;
    mlp_simt_kernel<<<cdiv(M, SIMT_ROWS), SIMT_THREADS, sm, st>>>(pts, viewdirs, 3, M, S, *net, multires, multires_views, raw, SimtSave{});
    NB_LAUNCH_OK("mlp_simt_kernel");
    return 0;
  }
  NB_CHECK_ARG(precision == NERF_B200_PREC_TC_FP16, "unknown precision %d", precision);
  NerfPassOut out;
  memset(&out, 0, sizeof(out));
  out.raw = raw;
  return launch_march(nullptr, 0, nullptr, pts, viewdirs, 3, nullptr, N, S, net, packed, multires, multires_views, 0, 0,
                      &out, workspace, workspace_bytes, st);
}

int nerf_b200_raw2outputs(const float* raw, const float* z_vals, const float* rays_d, int d_stride, const float* noise,
                          int64_t N, int S, int white_bkgd, const NerfPassOut* out, void* stream) {
  NB_CHECK_ARG(raw && z_vals && rays_d && out, "NULL pointer");
  NB_CHECK_ARG(S >= 1, "S must be >= 1");
  if (N == 0) return 0;
  raw2outputs_kernel<<<cdiv(N * 32, 256), 256, 0, (cudaStream_t)stream>>>(raw, z_vals, rays_d, d_stride, noise, N, S, white_bkgd, *out);
  NB_LAUNCH_OK("raw2outputs_kernel");
  return 0;
}

int nerf_b200_raw2outputs_bwd(const float* raw, const float* z_vals, const float* rays_d, int d_stride, const float* noise,
                              int64_t N, int S, int white_bkgd, const float* g_rgb, float* d_raw, void* stream) {
  NB_CHECK_ARG(raw && z_vals && rays_d && g_rgb && d_raw, "NULL pointer");
  NB_CHECK_ARG(S >= 1 && S <= 1024, "S must be in [1,1024]");
  if (N == 0) return 0;
  const int wpb = 4;
  const size_t sm = (size_t)wpb * 4 * S * sizeof(float);
  if (int rc = smem_optin((const void*)raw2outputs_bwd_kernel, sm)) return rc;
  raw2outputs_bwd_kernel<<<cdiv(N, wpb), wpb * 32, sm, (cudaStream_t)stream>>>(raw, z_vals, rays_d, d_stride, noise, N, S, white_bkgd, g_rgb, d_raw);
  NB_LAUNCH_OK("raw2outputs_bwd_kernel");
  return 0;
}

int nerf_b200_pack_rays(const float* rays_o, const float* rays_d, const float* view_src, const NerfCamera* cam, int64_t N,
                        int64_t pixel0, int ndc, float near, float far, int use_viewdirs, float* out, void* stream) {
  NB_CHECK_ARG(out != nullptr, "NULL output");
  NB_CHECK_ARG((rays_o && rays_d) || (cam && !rays_o && !rays_d), "give rays_o and rays_d, or a camera to generate them");
  NB_CHECK_ARG(!ndc || cam, "ndc needs the camera (H, W, focal)");
  if (N == 0) return 0;
  PackRaysArgs a;
  memset(&a, 0, sizeof(a));
  a.rays_o = rays_o; a.rays_d = rays_d; a.view_src = view_src; a.N = N; a.pixel0 = pixel0;
  a.ndc = ndc; a.use_viewdirs = use_viewdirs; a.stride = use_viewdirs ? 11 : 8; a.near = near; a.far = far;
  if (cam) {
    a.H = cam->H; a.W = cam->W; a.fx = cam->fx; a.fy = cam->fy; a.cx = cam->cx; a.cy = cam->cy;
    memcpy(a.c2w, cam->c2w, sizeof(a.c2w));
    a.ndc_cw = (float)(-1.0 / ((double)cam->W / (2.0 * (double)cam->fx)));      // run_nerf_helpers.py:181 (focal = K[0][0])
    a.ndc_ch = (float)(-1.0 / ((double)cam->H / (2.0 * (double)cam->fx)));
  }
  pack_rays_kernel<<<cdiv(N, 256), 256, 0, (cudaStream_t)stream>>>(a, out);
  NB_LAUNCH_OK("pack_rays_kernel");
  return 0;
}

int nerf_b200_sample_pdf(const float* bins, const float* weights, const float* u, int64_t u_row_stride, int64_t N, int B,
                         int n_samples, float* samples, void* stream) {
  NB_CHECK_ARG(bins && weights && u && samples, "NULL pointer");
  NB_CHECK_ARG(B >= 2 && B <= 2048 && n_samples >= 1, "bad B / n_samples");
  if (N == 0) return 0;
  const int wpb = 4;
  const size_t sm = (size_t)wpb * 3 * B * sizeof(float);
  if (int rc = smem_optin((const void*)sample_pdf_kernel, sm)) return rc;
  sample_pdf_kernel<<<cdiv(N, wpb), wpb * 32, sm, (cudaStream_t)stream>>>(bins, weights, u, u_row_stride, N, B, n_samples, samples);
  NB_LAUNCH_OK("sample_pdf_kernel");
  return 0;
}

int nerf_b200_coarse_z(const float* rays, int ray_stride, const float* t_vals, const float* t_rand, int64_t N, int S,
                       int lindisp, float* z_vals, void* stream) {
  NB_CHECK_ARG(rays && t_vals && z_vals, "NULL pointer");
  NB_CHECK_ARG(ray_stride >= 8, "ray_stride must be >= 8");
  if (N == 0) return 0;
  coarse_z_kernel<<<cdiv(N * S, 256), 256, 0, (cudaStream_t)stream>>>(rays, ray_stride, t_vals, t_rand, N, S, lindisp, z_vals);
  NB_LAUNCH_OK("coarse_z_kernel");
  return 0;
}

int nerf_b200_fine_z(const float* z_vals, const float* weights, const float* u, int64_t u_row_stride, int64_t N, int S,
                     int n_imp, float* z_fine, float* z_samples, float* z_std, void* stream) {
  NB_CHECK_ARG(z_vals && weights && u && z_fine, "NULL pointer");
  NB_CHECK_ARG(S >= 3 && n_imp >= 1 && S + n_imp <= 4096, "bad S / n_imp");
  if (N == 0) return 0;
  int P = 1;
  while (P < S + n_imp) P <<= 1;
  const int wpb = 4;
  const size_t sm = (size_t)wpb * (3 * S + P) * sizeof(float);
  if (int rc = smem_optin((const void*)fine_z_kernel, sm)) return rc;
  fine_z_kernel<<<cdiv(N, wpb), wpb * 32, sm, (cudaStream_t)stream>>>(z_vals, weights, u, u_row_stride, N, S, n_imp, P, z_fine, z_samples, z_std);
  NB_LAUNCH_OK("fine_z_kernel");
  return 0;
}

size_t nerf_b200_march_workspace_bytes(int64_t N, int S) {
  size_t tc = (size_t)N * 128 * 4, exact = (size_t)N * S * 28;   // view-bias table | pts + raw scratch
  return (tc > exact ? tc : exact) + 256;
}

int nerf_b200_march(const float* rays, const float* z_vals, const float* noise, int64_t N, int S, const NerfNetParams* net,
                    const void* packed, const NerfRenderCfg* cfg, const NerfPassOut* out, void* workspace,
                    size_t workspace_bytes, void* stream) {
  NB_CHECK_ARG(rays && z_vals && net && cfg && out, "NULL pointer");
  if (N == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  NB_CHECK_ARG(cfg->ray_stride >= (net->use_viewdirs ? 11 : 8), "ray_stride %d too small", cfg->ray_stride);
  if (cfg->precision == NERF_B200_PREC_FP32) {
    // exact mode (validation path): materialise pts -> fp32 CUDA-core MLP -> raw2outputs kernel
    const long long M = N * (long long)S;
    const size_t need = (size_t)M * 12 + (out->raw ? 0 : (size_t)M * 16);
    NB_CHECK_ARG(workspace && workspace_bytes >= need, "exact-mode workspace too small: need %zu bytes", need);
    NB_CHECK_ARG(net->W <= SIMT_THREADS && net->W % 4 == 0, "exact path supports netwidth <= 256");
    float* pts = static_cast<float*>(workspace);
    float* raw = out->raw ? out->raw : pts + (size_t)M * 3;
    pts_kernel<<<cdiv(M, 256), 256, 0, st>>>(rays, cfg->ray_stride, z_vals, M, S, pts);
    NB_LAUNCH_OK("pts_kernel");
    const size_t sm = simt_smem_bytes(*net);
    if (int rc = smem_optin((const void*)mlp_simt_kernel, sm)) return rc;
    mlp_simt_kernel<<<cdiv(M, SIMT_ROWS), SIMT_THREADS, sm, st>>>(pts, rays + 8, cfg->ray_stride, M, S, *net, cfg->multires, cfg->multires_views, raw, SimtSave{});
    NB_LAUNCH_OK("mlp_simt_kernel");
    raw2outputs_kernel<<<cdiv(N * 32, 256), 256, 0, st>>>(raw, z_vals, rays + 3, cfg->ray_stride, noise, N, S, cfg->white_bkgd, *out);
    NB_LAUNCH_OK("raw2outputs_kernel");
    return 0;
  }
  return launch_march(rays, cfg->ray_stride, z_vals, nullptr, rays + 8, cfg->ray_stride, noise, N, S, net, packed,
                      cfg->multires, cfg->multires_views, cfg->white_bkgd, 1, out, workspace, workspace_bytes, st);
}

int nerf_b200_render_rays_fwd(const float* rays, int64_t N, const NerfRenderCfg* cfg, const NerfNetParams* net_coarse,
                              const void* packed_coarse, const NerfNetParams* net_fine, const void* packed_fine,
                              const float* t_vals, const float* u_det, const float* t_rand, const float* u_rand,
                              const float* noise0, const float* noise1, float* z_coarse, const NerfPassOut* coarse,
                              float* z_fine, float* z_std, const NerfPassOut* fine, void* workspace,
                              size_t workspace_bytes, void* stream) {
  NB_CHECK_ARG(rays && cfg && net_coarse && t_vals && z_coarse && coarse, "NULL pointer");
  if (N == 0) return 0;
  const int Sc = cfg->N_samples, Ni = cfg->N_importance;
  NB_CHECK_ARG(Sc >= 1 && Ni >= 0, "bad sample counts");
  NB_CHECK_ARG(!cfg->perturb || t_rand, "perturb > 0 needs t_rand");
  // z sampling (run_nerf.py:357-379)
  if (int rc = nerf_b200_coarse_z(rays, cfg->ray_stride, t_vals, cfg->perturb ? t_rand : nullptr, N, Sc, cfg->lindisp, z_coarse, stream)) return rc;
  // coarse pass (:381-386)
  if (int rc = nerf_b200_march(rays, z_coarse, noise0, N, Sc, net_coarse, packed_coarse, cfg, coarse, workspace, workspace_bytes, stream)) return rc;
  if (Ni == 0) return 0;
  NB_CHECK_ARG(coarse->weights && z_fine && fine, "N_importance > 0 needs coarse->weights, z_fine and fine outputs");
  // hierarchical sampling (:392-396, :412); det <=> perturb == 0 (:393)
  const float* u = cfg->perturb ? u_rand : u_det;
  NB_CHECK_ARG(u != nullptr, "missing u (u_det for perturb==0, u_rand otherwise)");
  if (int rc = nerf_b200_fine_z(z_coarse, coarse->weights, u, cfg->perturb ? Ni : 0, N, Sc, Ni, z_fine, nullptr, z_std, stream)) return rc;
  // fine pass on all S_c + N_importance samples (:397-403); network_fine None -> coarse net (:399)
  const NerfNetParams* nf = net_fine ? net_fine : net_coarse;
  const void* pf = net_fine ? packed_fine : packed_coarse;
  return nerf_b200_march(rays, z_fine, noise1, N, Sc + Ni, nf, pf, cfg, fine, workspace, workspace_bytes, stream);
}

// ---- exact-mode backward of one pass (see bwd_simt.cuh) -------------------------------------------
static const int BWD_RAYS_PER_SLAB = 512;     // rays recomputed + back-propagated per slab (bounds the workspace)

size_t nerf_b200_march_bwd_workspace_bytes(int64_t N, int S) {
  const long long rows = (long long)(N < BWD_RAYS_PER_SLAB ? N : BWD_RAYS_PER_SLAB) * S;
  // upper bound over supported nets; the last term: three fp16 tile images of 256 columns (experimental tensor-core GEMMs)
  return (size_t)rows * (63 + 63 + 16 * 256 + 256 + 128 + 8 + 512 + 128 + 3 + 3 * 128) * 4 + 1024 + 3 * 128 * 512 + 3 * 1024;
}

// ---- experimental: the backward's large GEMMs on tensor cores (bwd_tc.cuh), NERF_B200_BWD_TC=1; never validated on a GPU ----
struct TcScratch { uint8_t* a; uint8_t* b; uint8_t* o; float scale; };
static bool bwd_tc_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("NERF_B200_BWD_TC"); on = (e && e[0] == '1') ? 1 : 0; }
  return on == 1;
}
static int pack_img(const float* src, int ld, long long M, int ncols, int C, float scale, uint8_t* img, cudaStream_t st) {
  const long long n = ((M + 127) / 128) * 128 * (C >> 3);
  tile_pack_kernel<<<cdiv(n, 256), 256, 0, st>>>(src, ld, M, ncols, C, scale, img);
  NB_LAUNCH_OK("tile_pack_kernel");
  return 0;
}

static int gemm_nn(const float* A, int lda, const float* B, int ldb, float* C, int ldc, long long M, int N, int K, int beta, cudaStream_t st) {
  dim3 grid(cdiv(M, GT), cdiv(N, GT));
  sgemm_nn_kernel<<<grid, 256, 0, st>>>(A, lda, B, ldb, C, ldc, M, N, K, beta);
  NB_LAUNCH_OK("sgemm_nn_kernel");
  return 0;
}
static int gemm_tn(const float* A, int lda, const float* B, int ldb, float* C, int ldc, long long M, int K1, int N, cudaStream_t st) {
  const int slab = 2048;
  dim3 grid(cdiv(K1, GT), cdiv(N, GT), cdiv(M, slab));
  sgemm_tn_kernel<<<grid, 256, 0, st>>>(A, lda, B, ldb, C, ldc, M, K1, N, slab);
  NB_LAUNCH_OK("sgemm_tn_kernel");
  return 0;
}
// C[K1,N] += A^T B like gemm_tn, through wgrad_tiles_kernel when the shape fits (K1 in {128,256}, N <= 256) and tc != NULL
static int gemm_tn_any(const TcScratch* tc, const float* A, int lda, const float* B, int ldb, float* C, int ldc, long long M, int K1, int N, cudaStream_t st) {
  if (!tc || !(K1 == 128 || K1 == 256) || N > 256 || N < 8) return gemm_tn(A, lda, B, ldb, C, ldc, M, K1, N, st);
  const int Nc = (N <= 64) ? 64 : (N <= 128 ? 128 : 256);
  if (int rc = pack_img(A, lda, M, K1, K1, tc->scale, tc->a, st)) return rc;
  if (int rc = pack_img(B, ldb, M, N, Nc, 1.0f, tc->b, st)) return rc;
  if (int rc = smem_optin((const void*)wgrad_tiles_kernel, WG_TOTAL)) return rc;
  const long long n_tiles = (M + 127) / 128;
  const int grid = (int)(n_tiles < num_sms() ? n_tiles : num_sms());
  wgrad_tiles_kernel<<<grid, WG_THREADS, WG_TOTAL, st>>>(tc->a, tc->b, n_tiles, K1, Nc, 1.0f / tc->scale, C, ldc, N);
  NB_LAUNCH_OK("wgrad_tiles_kernel");
  return 0;
}
// C[M,N] = A B like gemm_nn (beta = 0), through dgrad_tiles_kernel when N == 256, K in {128,256} and tc != NULL
static int gemm_nn_any(const TcScratch* tc, const float* A, int lda, const float* B, int ldb, float* C, int ldc, long long M, int N, int K, int beta, cudaStream_t st) {
  if (!tc || beta != 0 || N != 256 || !(K == 128 || K == 256)) return gemm_nn(A, lda, B, ldb, C, ldc, M, N, K, beta, st);
  if (int rc = pack_img(A, lda, M, K, K, tc->scale, tc->a, st)) return rc;
  if (int rc = pack_img(B, ldb, K, 256, 256, 1.0f, tc->b, st)) return rc;          // W rows = reduction index
  if (int rc = smem_optin((const void*)dgrad_tiles_kernel, DG_TOTAL)) return rc;
  const long long n_tiles = (M + 127) / 128;
  const int grid = (int)(n_tiles < num_sms() ? n_tiles : num_sms());
  dgrad_tiles_kernel<<<grid, DG_THREADS, DG_TOTAL, st>>>(tc->a, tc->b, nullptr, n_tiles, K, tc->o);
  NB_LAUNCH_OK("dgrad_tiles_kernel");
  tile_unpack_kernel<<<cdiv(M * 32, 256), 256, 0, st>>>(tc->o, M, 256, 1.0f / tc->scale, C, ldc);
  NB_LAUNCH_OK("tile_unpack_kernel");
  return 0;
}
static int mask_colsum(float* d, int ldd, const float* h, int ldh, long long M, int C, float* colsum, cudaStream_t st) {
  const int rpb = 256;
  dim3 grid(cdiv(M, rpb), cdiv(C, 64));
  relu_mask_colsum_kernel<<<grid, 64, 0, st>>>(d, ldd, h, ldh, M, C, colsum, rpb);
  NB_LAUNCH_OK("relu_mask_colsum_kernel");
  return 0;
}
#define NB_TRY(expr) do { if (int rc__ = (expr)) return rc__; } while (0)

int nerf_b200_march_bwd(const float* rays, const float* z_vals, const float* noise, int64_t N, int S, const NerfNetParams* net,
                        const void* packed, const NerfRenderCfg* cfg, const float* g_rgb, const NerfNetGrads* grads,
                        void* workspace, size_t workspace_bytes, void* stream) {
  (void)packed;
  NB_CHECK_ARG(rays && z_vals && net && cfg && g_rgb && grads, "NULL pointer");
  NB_CHECK_ARG(net->use_viewdirs, "backward is implemented for use_viewdirs networks (the reference's shipped configs)");
  NB_CHECK_ARG(net->W <= SIMT_THREADS && net->W % 4 == 0 && net->D <= NERF_B200_MAX_D, "unsupported network shape");
  if (N == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int W = net->W, W2 = W / 2, IC = net->input_ch, ICV = net->input_ch_views, D = net->D, rs = cfg->ray_stride;
  NB_CHECK_ARG(workspace && workspace_bytes >= nerf_b200_march_bwd_workspace_bytes(N, S), "march_bwd workspace too small");
  for (int64_t n0 = 0; n0 < N; n0 += BWD_RAYS_PER_SLAB) {
    const int64_t nn = (N - n0 < BWD_RAYS_PER_SLAB) ? N - n0 : BWD_RAYS_PER_SLAB;
    const long long M = nn * (long long)S;
    const float* ry = rays + n0 * rs;
    const float* zz = z_vals + n0 * S;
    const float* nz = noise ? noise + n0 * S : nullptr;
    float* p = static_cast<float*>(workspace);
    float* pts = p;        p += M * 3;
    SimtSave sv;
    sv.enc = p;            p += M * IC;
    sv.encv = p;           p += M * ICV;
    sv.h = p;              p += (size_t)D * M * W;
    sv.feat = p;           p += M * W;
    sv.hv = p;             p += M * W2;
    float* raw = p;        p += M * 4;
    float* d_raw = p;      p += M * 4;
    float* dh0 = p;        p += M * W;
    float* dh1 = p;        p += M * W;
    float* d_hv = p;       p += M * W2;
    TcScratch tcs, *tc = nullptr;
    if (bwd_tc_enabled() && W == 256) {
      // three 256-column fp16 tile images (1 KB-aligned) behind the fp32 buffers; one static loss scale for the slab's
      // activation gradients: |dL/drgb| <= 2 / (3 N)  (tools/bwd_precision_study.py)
      const size_t ib = (size_t)((M + 127) / 128) * 65536;
      uint8_t* q = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~(uintptr_t)1023);
      tcs.a = q; tcs.b = q + ib; tcs.o = q + 2 * ib;
      tcs.scale = exp2f(floorf(log2f(3.0f * (float)N * 512.0f)));
      tc = &tcs;
    }
    // 1. recompute the pass in fp32 with saved activations
    pts_kernel<<<cdiv(M, 256), 256, 0, st>>>(ry, rs, zz, M, S, pts);
    NB_LAUNCH_OK("pts_kernel");
    const size_t sm = simt_smem_bytes(*net);
    NB_TRY(smem_optin((const void*)mlp_simt_kernel, sm));
    mlp_simt_kernel<<<cdiv(M, SIMT_ROWS), SIMT_THREADS, sm, st>>>(pts, ry + 8, rs, M, S, *net, cfg->multires, cfg->multires_views, raw, sv);
    NB_LAUNCH_OK("mlp_simt_kernel");
    // 2. compositing adjoint -> dL/draw
    NB_TRY(nerf_b200_raw2outputs_bwd(raw, zz, ry + 3, rs, nz, nn, S, cfg->white_bkgd, g_rgb + n0 * 3, d_raw, stream));
    const float* h_last = sv.h + (size_t)(D - 1) * M * W;
    // 3. rgb_linear (run_nerf_helpers.py:114)
    NB_TRY(gemm_tn_any(tc, d_raw, 4, sv.hv, W2, grads->rgb_w, W2, M, 3, W2, st));
    NB_TRY(mask_colsum(d_raw, 4, nullptr, 0, M, 3, grads->rgb_b, st));
    NB_TRY(gemm_nn_any(tc, d_raw, 4, net->rgb_w, W2, d_hv, W2, M, W2, 3, 0, st));
    // 4. views_linears[0] on cat([feature, input_views]) (:108-112)
    NB_TRY(mask_colsum(d_hv, W2, sv.hv, W2, M, W2, grads->views_b, st));
    NB_TRY(gemm_tn_any(tc, d_hv, W2, sv.feat, W, grads->views_w, W + ICV, M, W2, W, st));
    NB_TRY(gemm_tn_any(tc, d_hv, W2, sv.encv, ICV, grads->views_w + W, W + ICV, M, W2, ICV, st));
    NB_TRY(gemm_nn_any(tc, d_hv, W2, net->views_w, W + ICV, dh0, W, M, W, W2, 0, st));                // d_feature
    // 5. feature_linear and alpha_linear both read the last hidden layer (:106-107)
    NB_TRY(gemm_tn_any(tc, dh0, W, h_last, W, grads->feature_w, W, M, W, W, st));
    NB_TRY(mask_colsum(dh0, W, nullptr, 0, M, W, grads->feature_b, st));
    NB_TRY(gemm_tn_any(tc, d_raw + 3, 4, h_last, W, grads->alpha_w, W, M, 1, W, st));
    NB_TRY(mask_colsum(d_raw + 3, 4, nullptr, 0, M, 1, grads->alpha_b, st));
    NB_TRY(gemm_nn_any(tc, dh0, W, net->feature_w, W, dh1, W, M, W, W, 0, st));
    NB_TRY(gemm_nn_any(tc, d_raw + 3, 4, net->alpha_w, W, dh1, W, M, W, 1, 1, st));
    // 6. pts_linears, last to first (:99-103); skip layer input = cat([input_pts, h])
    float* dcur = dh1;
    float* dnext = dh0;
    for (int l = D - 1; l >= 0; --l) {
      NB_TRY(mask_colsum(dcur, W, sv.h + (size_t)l * M * W, W, M, W, grads->pts_b[l], st));
      const bool after_skip = (l > 0) && (l - 1 == net->skip);
      const int Kl = (l == 0) ? IC : (after_skip ? W + IC : W);
      if (l == 0) {
        NB_TRY(gemm_tn_any(tc, dcur, W, sv.enc, IC, grads->pts_w[0], Kl, M, W, IC, st));
      } else {
        const float* hprev = sv.h + (size_t)(l - 1) * M * W;
        const int off = after_skip ? IC : 0;
        if (after_skip) NB_TRY(gemm_tn_any(tc, dcur, W, sv.enc, IC, grads->pts_w[l], Kl, M, W, IC, st));
        NB_TRY(gemm_tn_any(tc, dcur, W, hprev, W, grads->pts_w[l] + off, Kl, M, W, W, st));
        NB_TRY(gemm_nn_any(tc, dcur, W, net->pts_w[l] + off, Kl, dnext, W, M, W, W, 0, st));
        float* t = dcur; dcur = dnext; dnext = t;
      }
    }
  }
  return 0;
}

int nerf_b200_debug_set_trace(void* dev_buf_4096_i64) { g_trace = static_cast<long long*>(dev_buf_4096_i64); return 0; }

int nerf_b200_debug_mma_rate(int reps, int N, int b_sw64, void* out_2_i64, void* stream) {
  const size_t sm = 65536 + 32768 + 256 + 1024;
  if (int rc = smem_optin((const void*)mma_rate_kernel, sm)) return rc;
  mma_rate_kernel<<<1, 128, sm, (cudaStream_t)stream>>>(reps, N, b_sw64, static_cast<long long*>(out_2_i64));
  NB_LAUNCH_OK("mma_rate_kernel");
  return 0;
}

int nerf_b200_debug_epi_rate(int reps, int mode, int mma, void* out_2_i64, void* stream) {
  const size_t sm = 65536 + 49152 + 1024 + 256 + 1024;
  if (int rc = smem_optin((const void*)epi_rate_kernel, sm)) return rc;
  epi_rate_kernel<<<1, 384, sm, (cudaStream_t)stream>>>(reps, mode, mma, static_cast<long long*>(out_2_i64));
  NB_LAUNCH_OK("epi_rate_kernel");
  return 0;
}

int nerf_b200_debug_ldtm_rate(int reps, int shape, int nwarps, int mma, void* out_2_i64, void* stream) {
  const size_t sm = 49152 + 256 + 1024;
  if (int rc = smem_optin((const void*)ldtm_rate_kernel, sm)) return rc;
  ldtm_rate_kernel<<<1, 384, sm, (cudaStream_t)stream>>>(reps, shape, nwarps, mma, static_cast<long long*>(out_2_i64));
  NB_LAUNCH_OK("ldtm_rate_kernel");
  return 0;
}

int nerf_b200_debug_issue_probe(int reps, int nmma, int flags, void* out_2_i64, void* stream) {
  const size_t sm = 65536 + 32768 + 256 + 1024;
  if (int rc = smem_optin((const void*)issue_probe_kernel, sm)) return rc;
  issue_probe_kernel<<<1, 128, sm, (cudaStream_t)stream>>>(reps, nmma, flags, static_cast<long long*>(out_2_i64));
  NB_LAUNCH_OK("issue_probe_kernel");
  return 0;
}

int nerf_b200_debug_l2_stream(const void* buf, int buf_bytes, int chunk, int stages, int passes, int nblocks, void* out_i64, void* stream) {
  const size_t sm = (size_t)stages * chunk + 512 + 1024;
  if (int rc = smem_optin((const void*)l2_stream_probe_kernel, sm)) return rc;
  l2_stream_probe_kernel<<<nblocks, 64, sm, (cudaStream_t)stream>>>(static_cast<const uint8_t*>(buf), buf_bytes, chunk, stages, passes, static_cast<long long*>(out_i64));
  NB_LAUNCH_OK("l2_stream_probe_kernel");
  return 0;
}

// ---- experimental building blocks of the tensor-core backward (bwd_tc.cuh); not used by any default path ----
int nerf_b200_exp_tile_pack(const float* src, int64_t M, int C, float scale, void* img, void* stream) {
  NB_CHECK_ARG(src && img && M >= 0 && (C == 64 || C == 128 || C == 256), "bad arguments (C must be 64, 128 or 256)");
  if (M == 0) return 0;
  const long long n = ((M + 127) / 128) * 128 * (C >> 3);
  tile_pack_kernel<<<cdiv(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, C, M, C, C, scale, static_cast<uint8_t*>(img));
  NB_LAUNCH_OK("tile_pack_kernel");
  return 0;
}
int nerf_b200_exp_tile_unpack(const void* img, int64_t M, int C, float scale, float* dst, void* stream) {
  NB_CHECK_ARG(dst && img && M >= 0 && (C == 64 || C == 128 || C == 256), "bad arguments (C must be 64, 128 or 256)");
  if (M == 0) return 0;
  tile_unpack_kernel<<<cdiv(M * (C >> 3), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint8_t*>(img), M, C, scale, dst, C);
  NB_LAUNCH_OK("tile_unpack_kernel");
  return 0;
}
int nerf_b200_exp_tile_colsum(const void* img, int64_t n_tiles, int C, float scale, float* colsum, void* stream) {
  NB_CHECK_ARG(colsum && img && n_tiles >= 0 && (C == 64 || C == 128 || C == 256), "bad arguments (C must be 64, 128 or 256)");
  if (n_tiles == 0) return 0;
  const int grid = (int)(n_tiles < 4 * num_sms() ? n_tiles : 4 * num_sms());
  tile_colsum_kernel<<<grid, 256, C * sizeof(float), static_cast<cudaStream_t>(stream)>>>(static_cast<const uint8_t*>(img), n_tiles, C, scale, colsum);
  NB_LAUNCH_OK("tile_colsum_kernel");
  return 0;
}
int nerf_b200_exp_wgrad_tiles(const void* ximg, const void* yimg, int64_t n_tiles, int Mc, int Nc, float scale, float* dW, int ldw, void* stream) {
  NB_CHECK_ARG(ximg && yimg && dW && n_tiles >= 0, "null pointer");
  NB_CHECK_ARG((Mc == 128 || Mc == 256) && (Nc == 64 || Nc == 128 || Nc == 256) && ldw >= Nc, "unsupported shape Mc=%d Nc=%d ldw=%d", Mc, Nc, ldw);
  if (n_tiles == 0) return 0;
  if (int rc = smem_optin((const void*)wgrad_tiles_kernel, WG_TOTAL)) return rc;
  const int grid = (int)(n_tiles < num_sms() ? n_tiles : num_sms());
  wgrad_tiles_kernel<<<grid, WG_THREADS, WG_TOTAL, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint8_t*>(ximg), static_cast<const uint8_t*>(yimg),
                                                                                 n_tiles, Mc, Nc, scale, dW, ldw, Nc);
  NB_LAUNCH_OK("wgrad_tiles_kernel");
  return 0;
}
int nerf_b200_exp_dgrad_tiles(const void* ximg, const void* wimg, const void* himg, int64_t n_tiles, int Kc, void* oimg, void* stream) {
  NB_CHECK_ARG(ximg && wimg && oimg && n_tiles >= 0, "null pointer");
  NB_CHECK_ARG(Kc == 128 || Kc == 256, "unsupported reduction width Kc=%d", Kc);
  if (n_tiles == 0) return 0;
  if (int rc = smem_optin((const void*)dgrad_tiles_kernel, DG_TOTAL)) return rc;
  const int grid = (int)(n_tiles < num_sms() ? n_tiles : num_sms());
  dgrad_tiles_kernel<<<grid, DG_THREADS, DG_TOTAL, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint8_t*>(ximg), static_cast<const uint8_t*>(wimg),
                                                                                 static_cast<const uint8_t*>(himg), n_tiles, Kc, static_cast<uint8_t*>(oimg));
  NB_LAUNCH_OK("dgrad_tiles_kernel");
  return 0;
}

int nerf_b200_selftest_gemm_tn(const float* X, const float* Y, float* out, int lbo_bytes, int sbo_bytes, void* stream) {
  NB_CHECK_ARG(X && Y && out, "null pointer");
  NB_CHECK_ARG(lbo_bytes >= 0 && sbo_bytes >= 0 && lbo_bytes % 16 == 0 && sbo_bytes % 16 == 0, "lbo / sbo must be multiples of 16 bytes");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int sm = 131072 + 256 + 1024;
  if (int rc = smem_optin((const void*)selftest_gemm_tn_kernel, sm)) return rc;
  selftest_gemm_tn_kernel<<<1, 128, sm, st>>>(X, Y, out, (uint32_t)lbo_bytes, (uint32_t)sbo_bytes);
  NB_LAUNCH_OK("selftest_gemm_tn_kernel");
  return 0;
}

int nerf_b200_selftest_gemm(const float* A, const float* W, int K, int N, float* out, void* scratch, size_t scratch_bytes, void* stream) {
  NB_CHECK_ARG(A && W && out && scratch, "NULL pointer");
  NB_CHECK_ARG(K % 32 == 0 && K >= 32 && K <= 256 && (N == 128 || N == 256), "bad K/N");
  NB_CHECK_ARG(scratch_bytes >= (size_t)N * K * 2, "scratch too small");
  cudaStream_t st = (cudaStream_t)stream;
  PackJob job;
  job.n = 0;
  for (int c = 0; c < K / 32; ++c) { PackChunk& pc = job.c[job.n++]; pc.src = W; pc.ld = K; pc.k0 = 32 * c; pc.kvalid = 32; pc.nrows = N; pc.dst_off = (unsigned)c * N * 64; }
  dim3 grid(4, job.n);
  pack_chunks_kernel<<<grid, 256, 0, st>>>(job, static_cast<uint8_t*>(scratch));
  NB_LAUNCH_OK("pack_chunks_kernel");
  const size_t sm = 65536 + 16384 + 256 + 1024;
  if (int rc = smem_optin((const void*)selftest_gemm_kernel, sm)) return rc;
  selftest_gemm_kernel<<<1, 128, sm, st>>>(A, static_cast<const uint8_t*>(scratch), K, N, out);
  NB_LAUNCH_OK("selftest_gemm_kernel");
  return 0;
}

}  // extern "C"
