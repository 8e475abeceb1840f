// Backward passes of the encoder building blocks (SURVEY.md §8a rows C, D, E, M when the encoders train:
// savi_pretraining.yaml / savi_interactive_*.yaml `freeze_encoders: False`, av_nav always) and the GRU state
// encoder (row H).  fp32 SIMT, reference-accurate; every kernel also runs in the host emulation.
//
//   conv dgrad : dx[n,ih,iw,ci] = sum_{r,s,co} dy[n,oh,ow,co] * w[co,ci,r,s],  oh*stride - pad + r == ih
//   conv wgrad : dw[co,ci,r,s] += sum_{n,oh,ow} dy[n,oh,ow,co] * x[n,oh*stride-pad+r, ow*stride-pad+s, ci]
//   GroupNorm backward (fused ReLU / residual split), GRU cell forward / backward (gate fusion).
// Weights and weight gradients stay in the reference's OIHW layout (nn.Conv2d.weight), activations NHWC.
#include "nn_kernels.cuh"

namespace {

// 128 x 64 output tile, 16-deep k tiles, 8 x 4 register tile per thread (same shape as gemm_kernel)
__device__ __forceinline__ void tile_fma(const float (*As)[GBM + 4], const float (*Bs)[GBN + 4], int ty, int tx,
                                         float (&acc)[8][4]) {
#pragma unroll
  for (int k = 0; k < GBK; ++k) {
    float a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = As[k][ty * 8 + i];
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
}

struct BwdGeom {
  int N, H, W, C, KH, KW, stride, pad, OH, OW, Cout;
};

// ------------------------------------------------------------------------------------ conv dgrad
// GEMM view: M = N*H*W input pixels, Ncols = C input channels, K = KH*KW*Cout with k = (r*KW + s)*Cout + co.
__global__ void __launch_bounds__(GTHREADS, 2) conv_dgrad_kernel(const float* __restrict__ dy,
                                                                 const float* __restrict__ w, float* dx, BwdGeom g,
                                                                 int accumulate) {
  __shared__ __align__(16) float As[GBK][GBM + 4];
  __shared__ __align__(16) float Bs[GBK][GBN + 4];
  const int tid = threadIdx.x;
  const int M = g.N * g.H * g.W, K = g.KH * g.KW * g.Cout;
  const int m0 = blockIdx.x * GBM, n0 = blockIdx.y * GBN;
  const int ty = tid >> 4, tx = tid & 15;
  const int kk = tid & 15;  // this thread's k within a tile (A and B)
  const int r16 = tid >> 4;
  long long a_base[8];
  int a_ih[8], a_iw[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    int m = m0 + r16 + 16 * j;
    if (m < M) {
      int iw = m % g.W;
      int t = m / g.W;
      int ih = t % g.H;
      int n = t / g.H;
      a_base[j] = (long long)n * g.OH * g.OW * g.Cout;
      a_ih[j] = ih + g.pad;
      a_iw[j] = iw + g.pad;
    } else {
      a_base[j] = -1;
      a_ih[j] = a_iw[j] = 0;
    }
  }
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float ra[8], rb[4];
  auto fetch = [&](int k0) {
    const int k = k0 + kk;
    const bool k_ok = k < K;
    int co = 0, r = 0, s = 0;
    if (k_ok) {
      int rs = k / g.Cout;
      co = k - rs * g.Cout;
      r = rs / g.KW;
      s = rs - r * g.KW;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = 0.f;
      if (k_ok && a_base[j] >= 0) {
        int th = a_ih[j] - r, tw = a_iw[j] - s;
        if (th >= 0 && tw >= 0) {
          int oh = th / g.stride, ow = tw / g.stride;
          if (oh * g.stride == th && ow * g.stride == tw && oh < g.OH && ow < g.OW)
            v = __ldg(dy + a_base[j] + ((long long)oh * g.OW + ow) * g.Cout + co);
        }
      }
      ra[j] = v;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int ci = n0 + r16 + 16 * j;
      rb[j] = (k_ok && ci < g.C) ? __ldg(w + (((long long)co * g.C + ci) * g.KH + r) * g.KW + s) : 0.f;
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < K; k0 += GBK) {
#pragma unroll
    for (int j = 0; j < 8; ++j) As[kk][r16 + 16 * j] = ra[j];
#pragma unroll
    for (int j = 0; j < 4; ++j) Bs[kk][r16 + 16 * j] = rb[j];
    __syncthreads();
    if (k0 + GBK < K) fetch(k0 + GBK);
    tile_fma(As, Bs, ty, tx, acc);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = m0 + ty * 8 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= g.C) continue;
      float* p = dx + (long long)m * g.C + n;
      *p = accumulate ? *p + acc[i][j] : acc[i][j];
    }
  }
}

// ------------------------------------------------------------------------------------ conv wgrad
// GEMM view: M = KH*KW*C with m = (r*KW + s)*C + ci, Ncols = Cout, reduction over the P = N*OH*OW output pixels,
// split over blockIdx.z with atomic accumulation into dw (OIHW).
__global__ void __launch_bounds__(GTHREADS, 2) conv_wgrad_kernel(const float* __restrict__ x,
                                                                 const float* __restrict__ dy, float* dw, BwdGeom g,
                                                                 int p_per_split) {
  __shared__ __align__(16) float As[GBK][GBM + 4];
  __shared__ __align__(16) float Bs[GBK][GBN + 4];
  __shared__ long long px_base[2][GBK];
  __shared__ int px_ih[2][GBK], px_iw[2][GBK];
  const int tid = threadIdx.x;
  const int KK = g.KH * g.KW * g.C;
  const long long P = (long long)g.N * g.OH * g.OW;
  const int m0 = blockIdx.x * GBM, n0 = blockIdx.y * GBN;
  const long long pbeg = (long long)blockIdx.z * p_per_split;
  if (pbeg >= P) return;
  const long long pend = (pbeg + p_per_split < P) ? pbeg + p_per_split : P;
  const int ty = tid >> 4, tx = tid & 15;
  // A: one m per thread (tid % 128), 8 pixels per tile (tid / 128 + 2 j)
  const int am = tid & 127, ak0 = tid >> 7;
  int a_r = 0, a_s = 0, a_ci = 0;
  const bool a_ok = (m0 + am) < KK;
  if (a_ok) {
    int m = m0 + am;
    int rs = m / g.C;
    a_ci = m - rs * g.C;
    a_r = rs / g.KW;
    a_s = rs - a_r * g.KW;
  }
  // B: one co per thread (tid % 64), 4 pixels per tile (tid / 64 + 4 j)
  const int bn = tid & 63, bk0 = tid >> 6;
  const bool b_ok = (n0 + bn) < g.Cout;

  auto decode = [&](long long p0, int buf) {
    if (tid < GBK) {
      long long p = p0 + tid;
      if (p < pend) {
        int ow = (int)(p % g.OW);
        long long t = p / g.OW;
        int oh = (int)(t % g.OH);
        long long n = t / g.OH;
        px_base[buf][tid] = n * g.H * g.W * g.C;
        px_ih[buf][tid] = oh * g.stride - g.pad;
        px_iw[buf][tid] = ow * g.stride - g.pad;
      } else {
        px_base[buf][tid] = -1;
        px_ih[buf][tid] = px_iw[buf][tid] = 0;
      }
    }
  };
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float ra[8], rb[4];
  auto fetch = [&](long long p0, int buf) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int k = ak0 + 2 * j;
      float v = 0.f;
      long long base = px_base[buf][k];
      if (a_ok && base >= 0) {
        int ih = px_ih[buf][k] + a_r, iw = px_iw[buf][k] + a_s;
        if (ih >= 0 && ih < g.H && iw >= 0 && iw < g.W) v = __ldg(x + base + ((long long)ih * g.W + iw) * g.C + a_ci);
      }
      ra[j] = v;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int k = bk0 + 4 * j;
      long long p = p0 + k;
      rb[j] = (b_ok && p < pend) ? __ldg(dy + p * g.Cout + n0 + bn) : 0.f;
    }
  };
  decode(pbeg, 0);
  __syncthreads();
  fetch(pbeg, 0);
  int buf = 0;
  for (long long p0 = pbeg; p0 < pend; p0 += GBK) {
#pragma unroll
    for (int j = 0; j < 8; ++j) As[ak0 + 2 * j][am] = ra[j];
#pragma unroll
    for (int j = 0; j < 4; ++j) Bs[bk0 + 4 * j][bn] = rb[j];
    const bool more = p0 + GBK < pend;
    if (more) decode(p0 + GBK, buf ^ 1);
    __syncthreads();
    if (more) fetch(p0 + GBK, buf ^ 1);
    tile_fma(As, Bs, ty, tx, acc);
    __syncthreads();
    buf ^= 1;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = m0 + ty * 8 + i;
    if (m >= KK) continue;
    int rs = m / g.C;
    int ci = m - rs * g.C;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int co = n0 + tx * 4 + j;
      if (co >= g.Cout) continue;
      atomicAdd(dw + ((long long)co * g.C + ci) * (g.KH * g.KW) + rs, acc[i][j]);
    }
  }
}

// --------------------------------------------------------------------------------- GroupNorm backward
// Forward was y = act(gn(x) * gamma + beta (+ residual)).  One CTA per sample.
//   g      = dy * (y > 0)            (if relu)          -> also the gradient of the residual branch (dres)
//   dx     = rstd * (g*gamma - mean_grp(g*gamma) - xhat * mean_grp(g*gamma*xhat))
//   dgamma += sum g * xhat ; dbeta += sum g            (atomics over samples)
constexpr int GNB_THREADS = 512;
__global__ void __launch_bounds__(GNB_THREADS) groupnorm_bwd_kernel(const float* __restrict__ x,
                                                                    const float* __restrict__ y,
                                                                    const float* __restrict__ dy,
                                                                    const float* __restrict__ gamma, float* dx,
                                                                    float* dres, float* dgamma, float* dbeta, int HW,
                                                                    int C, int groups, float eps, int relu) {
  __shared__ float p0[GNB_THREADS], p1[GNB_THREADS], p2[GNB_THREADS], p3[GNB_THREADS];
  __shared__ double cS[GNB_THREADS], cQ[GNB_THREADS], cG[GNB_THREADS], cGX[GNB_THREADS];  // per channel (C <= 512)
  __shared__ float gmean[64], grstd[64], gm1[64], gm2[64];
  const int n = blockIdx.x, tid = threadIdx.x;
  const int c = tid % C;
  const int rpi = GNB_THREADS / C;
  const int cg = C / groups;
  const size_t base = (size_t)n * HW * C;
  float sx = 0.f, sxx = 0.f, sg = 0.f, sgx = 0.f;
  for (int p = tid / C; p < HW; p += rpi) {
    size_t i = base + (size_t)p * C + c;
    float xv = x[i];
    float gv = dy[i];
    if (relu && !(y[i] > 0.f)) gv = 0.f;
    if (dres) dres[i] = gv;
    sx += xv;
    sxx += xv * xv;
    sg += gv;
    sgx += gv * xv;
  }
  p0[tid] = sx; p1[tid] = sxx; p2[tid] = sg; p3[tid] = sgx;
  __syncthreads();
  if (tid < C) {
    double S = 0, Q = 0, G = 0, GX = 0;
    for (int t = tid; t < GNB_THREADS; t += C) { S += p0[t]; Q += p1[t]; G += p2[t]; GX += p3[t]; }
    cS[tid] = S; cQ[tid] = Q; cG[tid] = G; cGX[tid] = GX;
  }
  __syncthreads();
  if (tid < groups) {
    double S = 0, Q = 0, P1 = 0, P2 = 0;
    for (int k = tid * cg; k < (tid + 1) * cg; ++k) {
      S += cS[k]; Q += cQ[k];
      P1 += (double)gamma[k] * cG[k];
      P2 += (double)gamma[k] * cGX[k];
    }
    double cnt = (double)HW * cg;
    double m = S / cnt;
    double var = Q / cnt - m * m;
    if (var < 0.0) var = 0.0;
    double rs = 1.0 / sqrt(var + (double)eps);
    gmean[tid] = (float)m;
    grstd[tid] = (float)rs;
    gm1[tid] = (float)(P1 / cnt);
    gm2[tid] = (float)(rs * (P2 - m * P1) / cnt);
  }
  __syncthreads();
  if (tid < C) {
    const int g = tid / cg;
    if (dgamma) atomicAdd(&dgamma[tid], (float)((double)grstd[g] * (cGX[tid] - (double)gmean[g] * cG[tid])));
    if (dbeta) atomicAdd(&dbeta[tid], (float)cG[tid]);
  }
  if (!dx) return;
  const int g = c / cg;
  const float mu = gmean[g], rs = grstd[g], m1 = gm1[g], m2 = gm2[g], ga = gamma[c];
  for (int p = tid / C; p < HW; p += rpi) {
    size_t i = base + (size_t)p * C + c;
    float gv = dy[i];
    if (relu && !(y[i] > 0.f)) gv = 0.f;
    float xh = (x[i] - mu) * rs;
    dx[i] = rs * (gv * ga - m1 - xh * m2);
  }
}

// dy *= (y > 0) over a strided (rows, cols) view, in place (ReLU fused into the conv epilogue)
__global__ void relu_mask_kernel(float* dy, long long ldd, const float* __restrict__ y, long long ldy, long long rows,
                                 int cols) {
  const long long n = rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    long long r = i / cols;
    int c = (int)(i - r * cols);
    if (!(y[r * ldy + c] > 0.f)) dy[r * ldd + c] = 0.f;
  }
}

// gradient of the 2x2 area mean (+ scale): dx[n, h, w, c] = 0.25 * scale * dy[n, h/2, w/2, c]  (not needed for
// observations, kept for completeness of the resize op) -- omitted: observations never need gradients.

// ------------------------------------------------------------------------------------------- GRU
// torch.nn.GRU cell, gate order (r, z, n):  r = s(gi_r + gh_r), z = s(gi_z + gh_z), n = tanh(gi_n + r * gh_n),
// h' = (1 - z) * n + z * h.   gi / gh already carry their biases.  Saves (r, z, n, gh_n) for the backward pass and
// writes hm_next = h' * mask_next (the masked state the next step consumes, rnn_state_encoder.py:84-87).
__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }

__global__ void gru_gate_fwd_kernel(const float* __restrict__ gi, const float* __restrict__ gh,
                                    const float* __restrict__ h_in, float* h_out, float* save /* [N,4H] or null */,
                                    const float* __restrict__ mask_next /* [N] or null */, float* hm_next, int N,
                                    int H) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * H) return;
  int b = i / H, j = i - b * H;
  const float* gib = gi + (size_t)b * 3 * H;
  const float* ghb = gh + (size_t)b * 3 * H;
  float r = sigmoidf_(gib[j] + ghb[j]);
  float z = sigmoidf_(gib[H + j] + ghb[H + j]);
  float hn = ghb[2 * H + j];
  float nn = tanhf(gib[2 * H + j] + r * hn);
  float h = h_in[i];
  float o = (1.f - z) * nn + z * h;
  h_out[i] = o;
  if (save) {
    float* s = save + (size_t)b * 4 * H;
    s[j] = r; s[H + j] = z; s[2 * H + j] = nn; s[3 * H + j] = hn;
  }
  if (hm_next) hm_next[i] = mask_next ? o * mask_next[b] : o;
}

// hm = h * mask[b]
__global__ void gru_mask_kernel(const float* __restrict__ h, const float* __restrict__ mask, float* hm, int N, int H) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * H) return;
  hm[i] = h[i] * mask[i / H];
}

// dh = dout (+ dhm_next * mask_next); writes dgi, dgh (N, 3H) and dhm = dh * z (the direct path; the GEMM
// dgh . W_hh is accumulated on top by the caller).
__global__ void gru_gate_bwd_kernel(const float* __restrict__ dout /* may be null */,
                                    const float* __restrict__ dhm_next /* may be null */,
                                    const float* __restrict__ mask_next /* [N] or null */,
                                    const float* __restrict__ save, const float* __restrict__ h_in, float* dgi,
                                    float* dgh, float* dhm, int N, int H) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * H) return;
  int b = i / H, j = i - b * H;
  float dh = dout ? dout[i] : 0.f;
  if (dhm_next) dh += dhm_next[i] * (mask_next ? mask_next[b] : 1.f);
  const float* s = save + (size_t)b * 4 * H;
  float r = s[j], z = s[H + j], nn = s[2 * H + j], hn = s[3 * H + j];
  float dn = dh * (1.f - z) * (1.f - nn * nn);
  float dz = dh * (h_in[i] - nn) * z * (1.f - z);
  float dr = dn * hn * r * (1.f - r);
  float* a = dgi + (size_t)b * 3 * H;
  float* c = dgh + (size_t)b * 3 * H;
  a[j] = dr; a[H + j] = dz; a[2 * H + j] = dn;
  c[j] = dr; c[H + j] = dz; c[2 * H + j] = dn * r;
  dhm[i] = dh * z;
}

static int ew_grid2(long long total) {
  long long gsz = (total + 255) / 256;
  long long cap = (long long)avl_num_sms() * 16;
  if (gsz > cap) gsz = cap;
  if (gsz < 1) gsz = 1;
  return (int)gsz;
}

static int fill_geom(BwdGeom& g, int N, int H, int W, int C, int Cout, int KH, int KW, int stride, int pad) {
  if (N < 0 || H < 1 || W < 1 || C < 1 || Cout < 1 || KH < 1 || KW < 1 || stride < 1 || pad < 0) return AVL_ERR_ARG;
  g.N = N; g.H = H; g.W = W; g.C = C; g.KH = KH; g.KW = KW; g.stride = stride; g.pad = pad; g.Cout = Cout;
  g.OH = (H + 2 * pad - KH) / stride + 1;
  g.OW = (W + 2 * pad - KW) / stride + 1;
  if (g.OH < 1 || g.OW < 1) return AVL_ERR_ARG;
  if ((long long)N * H * W > 2147483647LL || (long long)N * g.OH * g.OW > 2147483647LL) return AVL_ERR_UNSUPPORTED;
  return AVL_OK;
}

static void gemm_plain(cudaStream_t s, const float* A, long long sa_m, long long sa_k, const float* B, long long sb_n,
                       long long sb_k, float* C, long long ldc, int M, int N, int K, const float* bias, int accumulate,
                       int splits) {
  GemmEpilogue ep;
  ep.bias = bias; ep.scale = nullptr; ep.residual = nullptr; ep.ldr = 0; ep.relu = 0; ep.accumulate = accumulate;
  ep.m_dev = nullptr; ep.k_dev = nullptr;
  ConvGeom g = {};
  if (splits < 1) splits = 1;
  int kps = ((K + splits - 1) / splits + GBK - 1) / GBK * GBK;
  splits = (K + kps - 1) / kps;
  dim3 grid(avl_div_up(M, GBM), avl_div_up(N, GBN), splits);
  GemmOperand a = {A, sa_m, sa_k}, b = {B, sb_n, sb_k};
  const bool akc = sa_k == 1, bkc = sb_k == 1;
  auto k11 = gemm_kernel<false, true, true>;
  auto k10 = gemm_kernel<false, true, false>;
  auto k01 = gemm_kernel<false, false, true>;
  auto k00 = gemm_kernel<false, false, false>;
  if (akc && bkc) AVL_LAUNCH(k11, grid, GTHREADS, 0, s, a, b, C, ldc, M, N, K, g, ep, kps);
  else if (akc) AVL_LAUNCH(k10, grid, GTHREADS, 0, s, a, b, C, ldc, M, N, K, g, ep, kps);
  else if (bkc) AVL_LAUNCH(k01, grid, GTHREADS, 0, s, a, b, C, ldc, M, N, K, g, ep, kps);
  else AVL_LAUNCH(k00, grid, GTHREADS, 0, s, a, b, C, ldc, M, N, K, g, ep, kps);
  avl_count_launch();
}

struct GruBufs {
  float *GI, *GH, *HM, *SAVE, *DGI, *DGH, *DHM;
};
static size_t gru_layout(char* base, GruBufs& b, size_t T, size_t N, int H, bool bwd) {
  size_t off = 0;
  auto take = [&](size_t n) {
    off = (off + 255) & ~(size_t)255;
    float* p = base ? reinterpret_cast<float*>(base + off) : nullptr;
    off += n * sizeof(float);
    return p;
  };
  b.GI = take(T * N * 3 * H);
  b.GH = take(N * 3 * H);
  b.HM = take((bwd ? T : 1) * N * H);
  b.SAVE = bwd ? take(T * N * 4 * H) : nullptr;
  b.DGI = bwd ? take(T * N * 3 * H) : nullptr;
  b.DGH = bwd ? take(T * N * 3 * H) : nullptr;
  b.DHM = bwd ? take(2 * N * H) : nullptr;
  return off + 256;
}

}  // namespace

// dx (N,H,W,C) (+)= dgrad of y = conv(x, w) given dy (N,OH,OW,Cout); w OIHW (Cout,C,KH,KW).
AVL_API int avl_conv2d_dgrad(const float* dy, const float* w, float* dx, int N, int H, int W, int C, int Cout, int KH,
                             int KW, int stride, int pad, int accumulate, void* stream) {
  BwdGeom g;
  int rc = fill_geom(g, N, H, W, C, Cout, KH, KW, stride, pad);
  if (rc) return rc;
  if (N == 0) return AVL_OK;
  if (!dy || !w || !dx) return AVL_ERR_ARG;
  dim3 grid(avl_div_up((long long)N * H * W, GBM), avl_div_up(C, GBN));
  AVL_LAUNCH(conv_dgrad_kernel, grid, GTHREADS, 0, (cudaStream_t)stream, dy, w, dx, g, accumulate);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}

// dw (Cout,C,KH,KW) += wgrad ; dbias (Cout) += column sums of dy (either may be NULL).
AVL_API int avl_conv2d_wgrad(const float* x, const float* dy, float* dw, float* dbias, int N, int H, int W, int C,
                             int Cout, int KH, int KW, int stride, int pad, void* stream) {
  BwdGeom g;
  int rc = fill_geom(g, N, H, W, C, Cout, KH, KW, stride, pad);
  if (rc) return rc;
  if (N == 0) return AVL_OK;
  if (!dy || (dw && !x)) return AVL_ERR_ARG;
  const long long P = (long long)N * g.OH * g.OW;
  cudaStream_t s = (cudaStream_t)stream;
  if (dw) {
    const int KK = KH * KW * C;
    int tiles = avl_div_up(KK, GBM) * avl_div_up(Cout, GBN);
    int splits = avl_div_up(4LL * avl_num_sms(), tiles);
    long long max_splits = (P + 255) / 256;
    if (splits > max_splits) splits = (int)max_splits;
    if (splits < 1) splits = 1;
    long long pps = ((P + splits - 1) / splits + GBK - 1) / GBK * GBK;
    splits = (int)((P + pps - 1) / pps);
    dim3 grid(avl_div_up(KK, GBM), avl_div_up(Cout, GBN), splits);
    AVL_LAUNCH(conv_wgrad_kernel, grid, GTHREADS, 0, s, x, dy, dw, g, (int)pps);
    AVL_LAUNCH_CHECK();
  }
  if (dbias) {
    long long gy = P / 64;
    if (gy < 1) gy = 1;
    if (gy > 512) gy = 512;
    AVL_LAUNCH(colsum_kernel, dim3(avl_div_up(Cout, 128), (int)gy), 128, 0, s, dy, (long long)Cout, nullptr, (int)P, Cout,
               dbias);
    AVL_LAUNCH_CHECK();
  }
  return AVL_OK;
}

// dy[r, c] = 0 where y[r, c] <= 0 (ReLU that was fused into a conv / linear epilogue), strided rows.
AVL_API int avl_relu_mask(float* dy, long long ldd, const float* y, long long ldy, long long rows, int cols,
                          void* stream) {
  if (rows < 0 || cols < 0) return AVL_ERR_ARG;
  if (rows == 0 || cols == 0) return AVL_OK;
  if (!dy || !y) return AVL_ERR_ARG;
  AVL_LAUNCH(relu_mask_kernel, ew_grid2(rows * cols), 256, 0, (cudaStream_t)stream, dy, ldd, y, ldy, rows, cols);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}

// Backward of avl_groupnorm_fwd: x = the GroupNorm INPUT, y = its output (only read when relu != 0).
// dx may alias dy.  dres (masked upstream gradient = gradient of the residual input) / dgamma / dbeta optional;
// dgamma, dbeta are accumulated.
AVL_API int avl_groupnorm_bwd(const float* x, const float* y, const float* dy, const float* gamma, float* dx,
                              float* dres, float* dgamma, float* dbeta, int N, int HW, int C, int groups, float eps,
                              int relu, void* stream) {
  if (N < 0 || HW < 1 || C < 1 || groups < 1 || groups > 64 || C % groups || GNB_THREADS % C) return AVL_ERR_UNSUPPORTED;
  if (N == 0) return AVL_OK;
  if (!x || !dy || !gamma || (relu && !y)) return AVL_ERR_ARG;
  AVL_LAUNCH(groupnorm_bwd_kernel, N, GNB_THREADS, 0, (cudaStream_t)stream, x, y, dy, gamma, dx, dres, dgamma, dbeta, HW,
             C, groups, eps, relu);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}

// ------------------------------------------------------------------------------------------- GRU (row H)
// ss_baselines/av_nav/models/rnn_state_encoder.py:80-149 (single_forward / seq_forward) around nn.GRU(I -> H, 1 layer).
// x (T*N, I) time-major; h0 (N, H); masks (T*N) float (0 at episode starts: the state entering step t is
// h_{t-1} * mask_t); weights in nn.GRU layout (gate order r, z, n).  out (T*N, H); h_last (N, H) may alias nothing.
AVL_API long long avl_gru_workspace_bytes(int T, int N, int I, int H, int with_backward) {
  GruBufs b;
  (void)I;
  return (long long)gru_layout(nullptr, b, (size_t)T, (size_t)N, H, with_backward != 0);
}

AVL_API int avl_gru_forward(int T, int N, int I, int H, const float* x, const float* h0, const float* masks,
                            const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, float* out,
                            float* h_last, void* workspace, int with_backward, void* stream) {
  if (T < 1 || N < 0 || I < 1 || H < 1) return AVL_ERR_ARG;
  if (N == 0) return AVL_OK;
  if (!x || !h0 || !w_ih || !w_hh || !out || !workspace) return AVL_ERR_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  GruBufs b;
  gru_layout(static_cast<char*>(workspace), b, (size_t)T, (size_t)N, H, with_backward != 0);
  const int NH = N * H, grid = avl_div_up(NH, 256);
  // input projections of every step at once: GI (T*N, 3H) = x W_ih^T + b_ih
  gemm_plain(s, x, I, 1, w_ih, I, 1, b.GI, 3 * H, T * N, 3 * H, I, b_ih, 0, 1);
  AVL_CUDA_CHECK(cudaGetLastError());
  if (masks) AVL_LAUNCH(gru_mask_kernel, grid, 256, 0, s, h0, masks, b.HM, N, H);
  else AVL_CUDA_CHECK(cudaMemcpyAsync(b.HM, h0, sizeof(float) * NH, cudaMemcpyDeviceToDevice, s));
  AVL_LAUNCH_CHECK();
  for (int t = 0; t < T; ++t) {
    float* hm = with_backward ? b.HM + (size_t)t * NH : b.HM;
    float* hm_next = (t + 1 < T) ? (with_backward ? b.HM + (size_t)(t + 1) * NH : b.HM) : nullptr;
    gemm_plain(s, hm, H, 1, w_hh, H, 1, b.GH, 3 * H, N, 3 * H, H, b_hh, 0, 1);
    AVL_LAUNCH(gru_gate_fwd_kernel, grid, 256, 0, s, b.GI + (size_t)t * N * 3 * H, b.GH, hm, out + (size_t)t * NH,
               with_backward ? b.SAVE + (size_t)t * N * 4 * H : nullptr,
               (masks && t + 1 < T) ? masks + (size_t)(t + 1) * N : nullptr, hm_next, N, H);
    AVL_LAUNCH_CHECK();
  }
  if (h_last) AVL_CUDA_CHECK(cudaMemcpyAsync(h_last, out + (size_t)(T - 1) * NH, sizeof(float) * NH, cudaMemcpyDeviceToDevice, s));
  return AVL_OK;
}

// dout (T*N, H) upstream gradient of `out`; dh_last (N, H) optional gradient of the returned state.
// Produces dx (T*N, I) (NULL to skip), dh0 (N, H) (NULL to skip); parameter gradients are ACCUMULATED (NULL = skip).
AVL_API int avl_gru_backward(int T, int N, int I, int H, const float* x, const float* masks, const float* w_ih,
                             const float* w_hh, const float* dout, const float* dh_last, float* dx, float* dh0,
                             float* dw_ih, float* dw_hh, float* db_ih, float* db_hh, void* workspace, void* stream) {
  if (T < 1 || N < 0 || I < 1 || H < 1) return AVL_ERR_ARG;
  if (N == 0) return AVL_OK;
  if (!x || !w_ih || !w_hh || !workspace || (!dout && !dh_last)) return AVL_ERR_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  GruBufs b;
  gru_layout(static_cast<char*>(workspace), b, (size_t)T, (size_t)N, H, true);
  const int NH = N * H, grid = avl_div_up(NH, 256);
  float* dhm_cur = b.DHM;
  float* dhm_prev = b.DHM + NH;
  if (dh_last) AVL_CUDA_CHECK(cudaMemcpyAsync(dhm_prev, dh_last, sizeof(float) * NH, cudaMemcpyDeviceToDevice, s));
  for (int t = T - 1; t >= 0; --t) {
    // gradient arriving at out[t]: dout[t] + (t == T-1 ? dh_last : dhm[t+1] * mask[t+1])
    const float* carry = (t == T - 1) ? (dh_last ? dhm_prev : nullptr) : dhm_prev;
    const float* cmask = (t == T - 1 || !masks) ? nullptr : masks + (size_t)(t + 1) * N;
    AVL_LAUNCH(gru_gate_bwd_kernel, grid, 256, 0, s, dout ? dout + (size_t)t * NH : nullptr, carry, cmask,
               b.SAVE + (size_t)t * N * 4 * H, b.HM + (size_t)t * NH, b.DGI + (size_t)t * N * 3 * H,
               b.DGH + (size_t)t * N * 3 * H, dhm_cur, N, H);
    AVL_LAUNCH_CHECK();
    // dhm_cur += dgh_t (N, 3H) . W_hh (3H, H)
    gemm_plain(s, b.DGH + (size_t)t * N * 3 * H, 3 * H, 1, w_hh, 1, H, dhm_cur, H, N, H, 3 * H, nullptr, 1, 1);
    float* tmp = dhm_cur; dhm_cur = dhm_prev; dhm_prev = tmp;
  }
  // dhm_prev now holds dhm[0] (gradient of h0 * mask_0)
  if (dh0) {
    if (masks) AVL_LAUNCH(gru_mask_kernel, grid, 256, 0, s, dhm_prev, masks, dh0, N, H);
    else AVL_CUDA_CHECK(cudaMemcpyAsync(dh0, dhm_prev, sizeof(float) * NH, cudaMemcpyDeviceToDevice, s));
    AVL_LAUNCH_CHECK();
  }
  const int R = T * N;
  int splits = R / 256;
  if (splits < 1) splits = 1;
  if (splits > 32) splits = 32;
  if (dw_hh) gemm_plain(s, b.DGH, 1, 3 * H, b.HM, 1, H, dw_hh, H, 3 * H, H, R, nullptr, 1, splits);
  if (dw_ih) gemm_plain(s, b.DGI, 1, 3 * H, x, 1, I, dw_ih, I, 3 * H, I, R, nullptr, 1, splits);
  int gy = R / 64;
  if (gy < 1) gy = 1;
  if (gy > 256) gy = 256;
  if (db_hh) {
    AVL_LAUNCH(colsum_kernel, dim3(avl_div_up(3 * H, 128), gy), 128, 0, s, b.DGH, (long long)3 * H, nullptr, R, 3 * H, db_hh);
    AVL_LAUNCH_CHECK();
  }
  if (db_ih) {
    AVL_LAUNCH(colsum_kernel, dim3(avl_div_up(3 * H, 128), gy), 128, 0, s, b.DGI, (long long)3 * H, nullptr, R, 3 * H, db_ih);
    AVL_LAUNCH_CHECK();
  }
  if (dx) gemm_plain(s, b.DGI, 3 * H, 1, w_ih, 1, I, dx, I, R, I, 3 * H, nullptr, 0, 1);
  AVL_CUDA_CHECK(cudaGetLastError());
  return AVL_OK;
}
