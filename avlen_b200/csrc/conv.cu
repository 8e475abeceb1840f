// Encoder building blocks: NHWC convolution / fully-connected as an
// im2col-gather GEMM with weights in the reference's native OIHW layout, GroupNorm(16) with fused
// residual + ReLU, the 128->64 area resize with the /255 RGB normalisation, max / average pooling.
// fp32 SIMT reference-accurate path (see gemm_tc.cu for the tensor-core path).
#include "nn_kernels.cuh"

namespace {

// ------------------------------------------------------------------ GroupNorm (NHWC, one CTA per sample)
// smt_resnet.py:22-33: nn.GroupNorm(16, C), eps 1e-5, affine.  y = gn(x) (+ residual) (ReLU).
constexpr int GN_THREADS = 512;
__global__ void __launch_bounds__(GN_THREADS) groupnorm_nhwc_kernel(const float* __restrict__ x,
                                                                    const float* __restrict__ gamma,
                                                                    const float* __restrict__ beta,
                                                                    const float* __restrict__ residual, float* y,
                                                                    int HW, int C, int groups, float eps, int relu) {
  __shared__ float ps[GN_THREADS], pq[GN_THREADS];
  __shared__ float gmean[64], grstd[64];
  const int n = blockIdx.x, tid = threadIdx.x;
  const int c = tid % C;            // C divides GN_THREADS
  const int rows_per_iter = GN_THREADS / C;
  const int cg = C / groups;
  const float* xs = x + (size_t)n * HW * C;
  float s = 0.f, q = 0.f;
  for (int p = tid / C; p < HW; p += rows_per_iter) {
    float v = xs[(size_t)p * C + c];
    s += v;
    q += v * v;
  }
  ps[tid] = s;
  pq[tid] = q;
  __syncthreads();
  if (tid < groups) {
    double S = 0.0, Q = 0.0;
    for (int t = 0; t < GN_THREADS; ++t) {
      if ((t % C) / cg == tid) { S += (double)ps[t]; Q += (double)pq[t]; }
    }
    double cnt = (double)HW * cg;
    double m = S / cnt;
    double var = Q / cnt - m * m;
    if (var < 0.0) var = 0.0;
    gmean[tid] = (float)m;
    grstd[tid] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  const int g = c / cg;
  const float a = grstd[g] * gamma[c];
  const float b = beta[c] - gmean[g] * a;
  float* ys = y + (size_t)n * HW * C;
  const float* rs = residual ? residual + (size_t)n * HW * C : nullptr;
  for (int p = tid / C; p < HW; p += rows_per_iter) {
    size_t i = (size_t)p * C + c;
    float v = fmaf(xs[i], a, b);
    if (rs) v += rs[i];
    if (relu) v = fmaxf(v, 0.f);
    ys[i] = v;
  }
}

// Split GroupNorm for large batches: (1) partial sums per (sample, group) from many CTAs per sample, accumulated in
// double with atomics; (2) a fully parallel vectorised apply pass (+ residual, ReLU).  stats: [N][groups][2] doubles.
__global__ void __launch_bounds__(256) gn_stats_kernel(const float* __restrict__ x, double* stats, int HW, int C,
                                                       int groups, int rows_per_cta) {
  __shared__ float ps[256], pq[256];
  const int n = blockIdx.x, tid = threadIdx.x;
  const int c = tid % C;  // C divides 256
  const int rpi = 256 / C;
  const int cg = C / groups;
  const int p0 = blockIdx.y * rows_per_cta;
  const int p1 = min(HW, p0 + rows_per_cta);
  const float* xs = x + (size_t)n * HW * C;
  float s = 0.f, q = 0.f;
  for (int p = p0 + tid / C; p < p1; p += rpi) {
    float v = xs[(size_t)p * C + c];
    s += v;
    q += v * v;
  }
  ps[tid] = s;
  pq[tid] = q;
  __syncthreads();
  if (tid < groups) {
    double S = 0.0, Q = 0.0;
    for (int t = 0; t < 256; ++t)
      if ((t % C) / cg == tid) { S += (double)ps[t]; Q += (double)pq[t]; }
    atomicAdd(&stats[((size_t)n * groups + tid) * 2], S);
    atomicAdd(&stats[((size_t)n * groups + tid) * 2 + 1], Q);
  }
}

// per-(sample, channel) affine: y = x * a + b  with a = rstd * gamma, b = beta - mean * a
__global__ void gn_finalize_kernel(const double* __restrict__ stats, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float2* ab, int N, int HW, int C, int groups,
                                   float eps) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * C) return;
  const int c = i % C, n = i / C;
  const int cg = C / groups, g = c / cg;
  const double cnt = (double)HW * cg;
  const double S = stats[((size_t)n * groups + g) * 2], Q = stats[((size_t)n * groups + g) * 2 + 1];
  const double m = S / cnt;
  double var = Q / cnt - m * m;
  if (var < 0.0) var = 0.0;
  const float a = (float)(1.0 / sqrt(var + (double)eps)) * gamma[c];
  ab[i] = make_float2(a, beta[c] - (float)m * a);
}

__global__ void __launch_bounds__(256) gn_apply_kernel(const float* __restrict__ x, const float2* __restrict__ ab,
                                                       const float* __restrict__ residual, float* y, long long total4,
                                                       int HW, int C, int relu) {
  // one float4 (4 consecutive channels of one pixel; C % 4 == 0) per iteration
  const int c4n = C >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4n) * 4;
    const long long pix = i / c4n;
    const int n = (int)(pix / HW);
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    const float4 r = residual ? reinterpret_cast<const float4*>(residual)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    const float2* p = ab + (size_t)n * C + c;
    const float2 p0 = __ldg(p), p1 = __ldg(p + 1), p2 = __ldg(p + 2), p3 = __ldg(p + 3);
    float4 o;
    o.x = fmaf(v.x, p0.x, p0.y) + r.x;
    o.y = fmaf(v.y, p1.x, p1.y) + r.y;
    o.z = fmaf(v.z, p2.x, p2.y) + r.z;
    o.w = fmaf(v.w, p3.x, p3.y) + r.w;
    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    reinterpret_cast<float4*>(y)[i] = o;
  }
}

// -------------------------------------------------- area resize (exact 2x2 mean) + optional scale (rgb / 255)
// smt_cnn.py:83-95 + common/utils.py:515-517 (interpolate(mode="area") 128 -> 64).  NHWC in, NHWC out.
__global__ void resize_half_kernel(const float* __restrict__ x, float* y, int N, int H, int W, int C, int Cp,
                                   float scale) {
  // Cp >= C output channels; channels [C, Cp) are zero (16-byte channel padding for the tensor-core conv loader)
  const int OH = H / 2, OW = W / 2;
  long long total = (long long)N * OH * OW * Cp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % Cp);
    long long t = i / Cp;
    if (c >= C) { y[i] = 0.f; continue; }
    int ow = (int)(t % OW);
    t /= OW;
    int oh = (int)(t % OH);
    int n = (int)(t / OH);
    const float* p = x + (((long long)n * H + 2 * oh) * W + 2 * ow) * C + c;
    // the reference divides by 255 first, then averages (smt_cnn.py:83-86)
    float a = p[0] * scale, b = p[C] * scale, d = p[(long long)W * C] * scale, e = p[(long long)W * C + C] * scale;
    y[i] = (a + b + d + e) * 0.25f;
  }
}

// cat([rgb / 255, depth], channel) for the av_nav VisualCNN (visual_cnn.py:143-150).  NHWC.
__global__ void concat_rgbd_kernel(const float* __restrict__ rgb, const float* __restrict__ depth, float* y,
                                   long long pixels, int c_rgb, int c_depth, float rgb_scale) {
  const int C = c_rgb + c_depth;
  long long total = pixels * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    long long p = i / C;
    y[i] = (c < c_rgb) ? rgb[p * c_rgb + c] * rgb_scale : depth[p * c_depth + (c - c_rgb)];
  }
}

// append `extra` (N, E) as E constant planes to an NHWC tensor (category label planes, audio_cnn.py:144-147)
__global__ void append_planes_kernel(const float* __restrict__ x, const float* __restrict__ extra, float* y,
                                     int N, int HW, int C, int E) {
  const int CO = C + E;
  long long total = (long long)N * HW * CO;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % CO);
    long long p = i / CO;
    int n = (int)(p / HW);
    y[i] = (c < C) ? x[p * C + c] : extra[(long long)n * E + (c - C)];
  }
}

// y (rows, Cp) = [x (rows, C), zeros]  (channel padding to a multiple of 4 for the tensor-core conv loader)
__global__ void pad_channels_kernel(const float* __restrict__ x, float* y, long long rows, int C, int Cp) {
  long long total = rows * Cp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % Cp);
    long long r = i / Cp;
    y[i] = (c < C) ? x[r * C + c] : 0.f;
  }
}

// max pool 3x3 stride 2 pad 1 (torchvision resnet18 stem), NHWC
__global__ void maxpool3x3s2_kernel(const float* __restrict__ x, float* y, int N, int H, int W, int C, int OH,
                                    int OW) {
  long long total = (long long)N * OH * OW * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    long long t = i / C;
    int ow = (int)(t % OW);
    t /= OW;
    int oh = (int)(t % OH);
    int n = (int)(t / OH);
    float m = -INFINITY;
    for (int r = 0; r < 3; ++r) {
      int ih = oh * 2 - 1 + r;
      if (ih < 0 || ih >= H) continue;
      for (int s = 0; s < 3; ++s) {
        int iw = ow * 2 - 1 + s;
        if (iw < 0 || iw >= W) continue;
        m = fmaxf(m, x[(((long long)n * H + ih) * W + iw) * C + c]);
      }
    }
    y[i] = m;
  }
}

// global average pool NHWC -> (N, C)
__global__ void avgpool_kernel(const float* __restrict__ x, float* y, int N, int HW, int C) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * C) return;
  int c = i % C, n = i / C;
  float s = 0.f;
  for (int p = 0; p < HW; ++p) s += x[((long long)n * HW + p) * C + c];
  y[i] = s / (float)HW;
}

// y[b, :] = W[:, a_b] + bias  (Linear on a one-hot action, policy.py:628-635 + action_encoder)
__global__ void onehot_linear_kernel(const long long* __restrict__ actions, const float* __restrict__ W,
                                     const float* __restrict__ bias, float* y, long long ldy, int B, int out_dim,
                                     int n_actions) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * out_dim) return;
  int o = i % out_dim, b = i / out_dim;
  int a = (int)actions[b];
  float v = bias[o];
  if (a >= 0 && a < n_actions) v += W[o * n_actions + a];
  y[(long long)b * ldy + o] = v;
}

// copy src (rows, cols) into a column slice of dst (rows, ldd)
__global__ void copy_cols_kernel(const float* __restrict__ src, long long lds, float* dst, long long ldd, int rows,
                                 int cols) {
  long long total = (long long)rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % cols);
    long long r = i / cols;
    dst[r * ldd + c] = src[r * lds + c];
  }
}

__global__ void zero_cols_kernel(float* y, long long ldy, int rows, int cols) {
  long long total = (long long)rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x)
    y[(i / cols) * ldy + (i % cols)] = 0.f;
}

__global__ void epilogue_cols_kernel(float* y, long long ldy, int rows, int cols, const float* __restrict__ scale,
                                     const float* __restrict__ bias, const float* __restrict__ residual, long long ldr,
                                     int relu) {
  long long total = (long long)rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % cols);
    long long r = i / cols;
    float v = y[r * ldy + c];
    if (scale) v *= scale[c];
    if (bias) v += bias[c];
    if (residual) v += residual[r * ldr + c];
    if (relu) v = fmaxf(v, 0.f);
    y[r * ldy + c] = v;
  }
}

static int ew_grid(long long total) {
  long long g = (total + 255) / 256;
  long long cap = (long long)avl_num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace

// NHWC convolution (also fully-connected layers after a flatten, as a kernel covering the whole map):
//   y[n, oh, ow, co] = act( scale[co] * sum_{r,s,ci} x[n, oh*stride - pad + r, ow*stride - pad + s, ci] * w[co, ci, r, s]
//                           + bias[co] + residual[n, oh, ow, co] )
// x: (N, H, W, C) ; w: (Cout, C, KH, KW) reference-native OIHW ; y: rows of length ldy (>= Cout), so the
// result can be written straight into a column slice of a wider feature matrix.
AVL_API int avl_conv2d_fwd(const float* x, int N, int H, int W, int C, const float* w, int Cout, int KH, int KW,
                           int stride, int pad, const float* scale, const float* bias, const float* residual,
                           long long ldr, int relu, float* y, long long ldy, void* stream) {
  if (N < 0 || H < 1 || W < 1 || C < 1 || Cout < 1 || KH < 1 || KW < 1 || stride < 1 || pad < 0) return AVL_ERR_ARG;
  if (N == 0) return AVL_OK;
  if (!x || !w || !y) return AVL_ERR_ARG;
  ConvGeom g;
  g.N = N; g.H = H; g.W = W; g.C = C; g.KH = KH; g.KW = KW; g.stride = stride; g.pad = pad;
  g.OH = (H + 2 * pad - KH) / stride + 1;
  g.OW = (W + 2 * pad - KW) / stride + 1;
  if (g.OH < 1 || g.OW < 1) return AVL_ERR_ARG;
  long long M = (long long)N * g.OH * g.OW;
  if (M > 2147483647LL) return AVL_ERR_UNSUPPORTED;
  const int K = KH * KW * C;
  GemmEpilogue ep;
  ep.bias = bias; ep.scale = scale; ep.residual = residual; ep.ldr = ldr; ep.relu = relu; ep.accumulate = 0;
  ep.m_dev = nullptr; ep.k_dev = nullptr;
  GemmOperand A = {x, 0, 1}, B = {w, (long long)K, 1};
  dim3 grid(avl_div_up(M, GBM), avl_div_up(Cout, GBN), 1);
  auto kern = gemm_kernel<true, true, true>;
  const int tiles = grid.x * grid.y;
  if (tiles * 4 <= avl_num_sms() && K >= 1024) {
    // few output tiles but a long reduction (FC layers at rollout batch sizes): split K over CTAs with atomic
    // accumulation into a zeroed output, then apply the epilogue in a second tiny kernel
    int splits = avl_num_sms() / tiles;
    int kps = ((K + splits - 1) / splits + GBK - 1) / GBK * GBK;
    if (kps < 128) kps = 128;
    splits = (K + kps - 1) / kps;
    grid.z = splits;
    AVL_LAUNCH(zero_cols_kernel, ew_grid(M * Cout), 256, 0, (cudaStream_t)stream, y, ldy, (int)M, Cout);
    AVL_LAUNCH_CHECK();
    GemmEpilogue raw = ep;
    raw.bias = raw.scale = raw.residual = nullptr; raw.relu = 0;
    AVL_LAUNCH(kern, grid, GTHREADS, 0, (cudaStream_t)stream, A, B, y, ldy, (int)M, Cout, K, g, raw, kps);
    AVL_LAUNCH_CHECK();
    AVL_LAUNCH(epilogue_cols_kernel, ew_grid(M * Cout), 256, 0, (cudaStream_t)stream, y, ldy, (int)M, Cout, scale,
               bias, residual, ldr, relu);
    AVL_LAUNCH_CHECK();
    return AVL_OK;
  }
  AVL_LAUNCH(kern, grid, GTHREADS, 0, (cudaStream_t)stream, A, B, y, ldy, (int)M, Cout, K, g, ep, K);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}

// stats_scratch: N*groups*2 doubles followed by N*C float2 (i.e. (2*N*groups + N*C) * 8 bytes).  Three kernels:
// partial sums from many CTAs per sample (double atomics), per-(n,c) affine, fully parallel float4 apply.
AVL_API int avl_groupnorm_fwd_split(const float* x, const float* gamma, const float* beta, const float* residual,
                                    float* y, int N, int HW, int C, int groups, float eps, int relu,
                                    double* stats_scratch, void* stream) {
  if (N < 0 || HW < 1 || C < 4 || (C & 3) || groups < 1 || groups > 64 || C % groups || 256 % C) return AVL_ERR_UNSUPPORTED;
  if (N == 0) return AVL_OK;
  if (!x || !gamma || !beta || !y || !stats_scratch) return AVL_ERR_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  AVL_CUDA_CHECK(cudaMemsetAsync(stats_scratch, 0, sizeof(double) * 2 * (size_t)N * groups, s));
  int rows_per_cta = 8192 / C;  // 32 KB of activations per CTA
  if (rows_per_cta > HW) rows_per_cta = HW;
  int splits = avl_div_up(HW, rows_per_cta);
  AVL_LAUNCH(gn_stats_kernel, dim3(N, splits), 256, 0, s, x, stats_scratch, HW, C, groups, rows_per_cta);
  AVL_LAUNCH_CHECK();
  float2* ab = reinterpret_cast<float2*>(stats_scratch + 2 * (size_t)N * groups);
  AVL_LAUNCH(gn_finalize_kernel, avl_div_up((long long)N * C, 256), 256, 0, s, stats_scratch, gamma, beta, ab, N, HW, C,
             groups, eps);
  AVL_LAUNCH_CHECK();
  long long total4 = (long long)N * HW * C / 4;
  AVL_LAUNCH(gn_apply_kernel, ew_grid(total4), 256, 0, s, x, ab, residual, y, total4, HW, C, relu);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}

AVL_API int avl_groupnorm_fwd(const float* x, const float* gamma, const float* beta, const float* residual, float* y,
                              int N, int HW, int C, int groups, float eps, int relu, void* stream) {
  if (N < 0 || HW < 1 || C < 1 || groups < 1 || groups > 64 || C % groups || GN_THREADS % C) return AVL_ERR_UNSUPPORTED;
  if (N == 0) return AVL_OK;
  if (!x || !gamma || !beta || !y) return AVL_ERR_ARG;
  AVL_LAUNCH(groupnorm_nhwc_kernel, N, GN_THREADS, 0, (cudaStream_t)stream, x, gamma, beta, residual, y, HW, C, groups,
             eps, relu);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}

AVL_API int avl_resize_half(const float* x, float* y, int N, int H, int W, int C, int C_out, float scale,
                            void* stream) {
  if (N < 0 || H < 2 || W < 2 || (H & 1) || (W & 1) || C < 1 || C_out < C) return AVL_ERR_ARG;
  if (N == 0) return AVL_OK;
  if (!x || !y) return AVL_ERR_ARG;
  AVL_LAUNCH(resize_half_kernel, ew_grid((long long)N * H * W * C_out / 4), 256, 0, (cudaStream_t)stream, x, y, N, H,
             W, C, C_out, scale);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}

AVL_API int avl_concat_rgbd(const float* rgb, const float* depth, float* y, long long pixels, int c_rgb, int c_depth,
                            float rgb_scale, void* stream) {
  if (pixels < 0 || c_rgb < 0 || c_depth < 0 || c_rgb + c_depth < 1) return AVL_ERR_ARG;
  if (pixels == 0) return AVL_OK;
  if (!y || (c_rgb && !rgb) || (c_depth && !depth)) return AVL_ERR_ARG;
  AVL_LAUNCH(concat_rgbd_kernel, ew_grid(pixels * (c_rgb + c_depth)), 256, 0, (cudaStream_t)stream, rgb, depth, y,
             pixels, c_rgb, c_depth, rgb_scale);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}

AVL_API int avl_append_planes(const float* x, const float* extra, float* y, int N, int HW, int C, int E,
                              void* stream) {
  if (N < 0 || HW < 1 || C < 1 || E < 1) return AVL_ERR_ARG;
  if (N == 0) return AVL_OK;
  if (!x || !extra || !y) return AVL_ERR_ARG;
  AVL_LAUNCH(append_planes_kernel, ew_grid((long long)N * HW * (C + E)), 256, 0, (cudaStream_t)stream, x, extra, y, N,
             HW, C, E);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}

AVL_API int avl_pad_channels(const float* x, float* y, long long rows, int C, int C_out, void* stream) {
  if (rows < 0 || C < 1 || C_out < C) return AVL_ERR_ARG;
  if (rows == 0) return AVL_OK;
  if (!x || !y) return AVL_ERR_ARG;
  AVL_LAUNCH(pad_channels_kernel, ew_grid(rows * C_out), 256, 0, (cudaStream_t)stream, x, y, rows, C, C_out);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}

AVL_API int avl_maxpool3x3s2(const float* x, float* y, int N, int H, int W, int C, void* stream) {
  if (N < 0 || H < 1 || W < 1 || C < 1) return AVL_ERR_ARG;
  if (N == 0) return AVL_OK;
  if (!x || !y) return AVL_ERR_ARG;
  int OH = (H + 2 - 3) / 2 + 1, OW = (W + 2 - 3) / 2 + 1;
  AVL_LAUNCH(maxpool3x3s2_kernel, ew_grid((long long)N * OH * OW * C), 256, 0, (cudaStream_t)stream, x, y, N, H, W, C,
             OH, OW);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}

AVL_API int avl_avgpool_global(const float* x, float* y, int N, int HW, int C, void* stream) {
  if (N < 0 || HW < 1 || C < 1) return AVL_ERR_ARG;
  if (N == 0) return AVL_OK;
  if (!x || !y) return AVL_ERR_ARG;
  AVL_LAUNCH(avgpool_kernel, avl_div_up((long long)N * C, 256), 256, 0, (cudaStream_t)stream, x, y, N, HW, C);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}

AVL_API int avl_onehot_linear(const long long* actions, const float* W, const float* bias, float* y, long long ldy,
                              int B, int out_dim, int n_actions, void* stream) {
  if (B < 0 || out_dim < 1 || n_actions < 1) return AVL_ERR_ARG;
  if (B == 0) return AVL_OK;
  if (!actions || !W || !bias || !y) return AVL_ERR_ARG;
  AVL_LAUNCH(onehot_linear_kernel, avl_div_up((long long)B * out_dim, 256), 256, 0, (cudaStream_t)stream, actions, W,
             bias, y, ldy, B, out_dim, n_actions);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}

AVL_API int avl_copy_cols(const float* src, long long lds, float* dst, long long ldd, int rows, int cols,
                          void* stream) {
  if (rows < 0 || cols < 0) return AVL_ERR_ARG;
  if (rows == 0 || cols == 0) return AVL_OK;
  if (!src || !dst) return AVL_ERR_ARG;
  AVL_LAUNCH(copy_cols_kernel, ew_grid((long long)rows * cols), 256, 0, (cudaStream_t)stream, src, lds, dst, ldd, rows,
             cols);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}
