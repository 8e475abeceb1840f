// Tensor-core backward of the encoder convolutions (SURVEY.md §8a rows C, D, E when the encoders train:
// savi_pretraining.yaml:53 `freeze_encoders: False`; ss_baselines/savi/models/smt_resnet.py:132-149, smt_cnn.py:78-115).
//
//   dgrad : dx = conv_stride1(dy [zero-upsampled when the forward stride was 2], W flipped and transposed)  — the forward
//           tensor-core kernels (halo-strip / im2col tcgen05) run it; this file only provides the weight re-packing
//           (avl_pack_conv_weight: ONE kernel per weight instead of ~6 ATen launches) and the zero-insertion kernel.
//   wgrad : dW[co, r, s, ci] = sum over (n, oh, ow) of dy[n, oh, ow, co] * x[n, oh*stride - pad + r, ow*stride - pad + s, ci]
//           A reduction over up to 2*10^7 pixels into a tiny (Cout x KH*KW*Cin) result: HBM-bound (x and dy are read
//           once, ~36 FLOP/B at 16 channels).  tc_conv_wgrad_kernel stages a strip of input rows (with its halo) and
//           the matching strip of dy rows in shared memory ONCE and lets every warp run TF32 tensor-core MMAs
//           (mma.sync m16n8k8, fp32 accumulate in registers) whose operands are read straight from the staged strip:
//           the A fragment is dy^T (16 output channels x 8 pixels), the B fragment of tap (r, s) is the SAME x strip
//           shifted by r rows and s pixels — no im2col expansion, neither in HBM nor in shared memory.  Accumulators
//           stay in registers across all strips of a CTA; per-CTA partial sums go to a workspace with plain stores and
//           a second kernel adds them in slice order (deterministic, no atomics).
//           Why warp-level MMAs and not tcgen05 here: the result has only 16..128 rows, so a tcgen05 M=64/128 tile
//           is mostly padding, and with shared-memory operand descriptors every tap re-fetches the dy tile
//           (operand-fetch bound); register fragments load dy once per 8 pixels and reuse it for all
//           KH*KW*Cin/8 column tiles.
#include "nn_kernels.cuh"
#include "tc_common.cuh"

#ifndef AVL_HOST_EMUL
namespace {

__device__ __forceinline__ float round_tf32_rn(float v) {  // round-to-nearest-even onto the TF32 grid
  uint32_t b = __float_as_uint(v);
  b += 0xFFFu + ((b >> 13) & 1u);
  return __uint_as_float(b & ~0x1FFFu);
}

// mode 0: out[(o, r, s, c)] = w[o][c][r][s]                              (Cout, KH, KW, c_pad)  forward packing
// mode 1: out[(c, r, s, o)] = w[o][c][KH-1-r][KW-1-s]                    (C, KH, KW, o_pad)     dgrad packing
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, float* out, int Cout, int C, int KH, int KW, int pad_to,
                                        int mode) {
  const int rows = mode ? C : Cout;
  const long long n = (long long)rows * KH * KW * pad_to;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int inner = (int)(i % pad_to);
    long long t = i / pad_to;
    const int s = (int)(t % KW);
    t /= KW;
    const int r = (int)(t % KH);
    const int outer = (int)(t / KH);
    float v = 0.f;
    if (mode == 0) {
      if (inner < C) v = w[(((long long)outer * C + inner) * KH + r) * KW + s];
    } else {
      if (inner < Cout) v = w[(((long long)inner * C + outer) * KH + (KH - 1 - r)) * KW + (KW - 1 - s)];
    }
    out[i] = round_tf32_rn(v);
  }
}

// up[n, 2*oh, 2*ow, :] = dy[n, oh, ow, :], zero elsewhere (H x W output, float4 granularity)
__global__ void zero_upsample2_kernel(const float4* __restrict__ dy, float4* up, int N, int OH, int OW, int H, int W,
                                      int c4) {
  const long long n = (long long)N * H * W * c4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4);
    long long t = i / c4;
    const int w = (int)(t % W);
    t /= W;
    const int h = (int)(t % H);
    const long long img = t / H;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!(h & 1) && !(w & 1) && (h >> 1) < OH && (w >> 1) < OW) v = __ldg(dy + ((img * OH + (h >> 1)) * OW + (w >> 1)) * c4 + c);
    up[i] = v;
  }
}

// ------------------------------------------------------------------------------------------------ wgrad
constexpr int WG_MAX_TILES = 18;  // m16 x n8 output tiles per warp: 72 accumulator registers

struct WgArgs {
  const float* x;
  const float* dy;
  float* part;       // [slices][Cout][taps][Cx]
  int N, H, W, Cx, Cout, KH, KW, stride, pad, OH, OW;
  int R;             // output rows per strip
  int strips_per_img;
  long long total_items;
  int in_rows, Wp;   // staged input rows / pixels per staged row (halo included)
  int xpitch, dpitch;  // floats per staged pixel (x / dy), chosen == 8 or 24 mod 32 (conflict-free fragment loads)
  int ncc;           // n8 tiles per tap: Cx / 8 (Cx == 4: tiles per kernel ROW = (KW + 1) / 2, two taps each)
  int ntiles_ct;     // tiles per 16 output channels
  int tpw;           // tiles per warp (divides ntiles_ct)
  int groups;        // warp groups per CTA (each owns tpw tiles of one cout tile)
  int ps;            // warps per group: they split the k-steps of a strip
  int group0_stride; // groups per blockIdx.y
  int total_groups;
  int ow_shift;      // log2(OW) or -1
};

// fp32 -> TF32 operand, rounded to nearest (ties away): add half an ulp of the 10-bit mantissa and let the tensor core
// drop the low 13 bits.  One integer add per value (cvt.rna.tf32.f32 expands to a 5-instruction sequence with
// Inf / NaN handling on sm_100a; activations and gradients are finite here).
__device__ __forceinline__ uint32_t f2tf32(float v) { return __float_as_uint(v) + 0x1000u; }
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// NT: compile-time tile slots per warp (>= p.tpw; surplus slots recompute tile 0 into accumulators nobody stores, so
// that the MMA loop has no per-tile branches)
// PAIR (Cx % 16 == 0): two adjacent column tiles of a tap share their operand loads — tile 2j takes the even and tile
// 2j + 1 the odd channels of a 16-channel block, so that one 8-byte shared-memory load feeds both MMAs (the K and N
// index spaces of an MMA may be permuted freely as long as the result is stored through the same permutation).  The
// rows of the dy^T fragment are permuted the same way for every shape: row g <-> output channel co0 + 2 g, row g + 8 <->
// co0 + 2 g + 1 (one 8-byte load per pixel instead of two 4-byte loads).
template <int NT, bool PAIR>
__global__ void __launch_bounds__(512, 1) tc_conv_wgrad_kernel(WgArgs p) {
  AVL_DYN_SMEM(smem_raw);
  float* Sx = reinterpret_cast<float*>(smem_raw);
  float* Sd = Sx + (size_t)p.in_rows * p.Wp * p.xpitch;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int grp = blockIdx.y * p.group0_stride + warp / p.ps;  // global warp-group index
  const int q = warp % p.ps;
  const bool active = grp < p.total_groups && (warp / p.ps) < p.groups;
  // this group's tiles: one cout tile, tpw consecutive column tiles
  const int groups_per_ct = p.ntiles_ct / p.tpw;
  const int ct = active ? grp / groups_per_ct : 0;
  const int tile0 = active ? (grp % groups_per_ct) * p.tpw : 0;
  const int co0 = ct * 16;
  const bool c4 = p.Cx == 4;
  int toff[NT];  // float offset of tile i inside the staged x strip (tap shift + channel chunk)
#pragma unroll
  for (int i = 0; i < NT; ++i) {
    const int id = tile0 + (i < p.tpw ? i : 0);
    int r, s, cc;
    if (c4) {
      r = id / p.ncc;
      s = 2 * (id - r * p.ncc);
      cc = 0;
    } else {
      const int tap = id / p.ncc;
      cc = id - tap * p.ncc;
      r = tap / p.KW;
      s = tap - r * p.KW;
    }
    toff[i] = (r * p.Wp + s) * p.xpitch + (PAIR ? (cc >> 1) * 16 : cc * 8);
  }
  float acc[NT][4];
#pragma unroll
  for (int i = 0; i < NT; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;

  const int xc4 = p.Cx >> 2, dc4 = p.Cout >> 2;
  const int strip_px = p.R * p.OW;
  const int nsteps = (strip_px + 7) >> 3;
  int xc4_shift = -1;  // log2(Cx / 4) when it is a power of two (every ResNet width), else divide
  for (int k = 0; k < 10; ++k)
    if ((1 << k) == xc4) xc4_shift = k;
  const int d_chunks = nsteps * 8 * dc4;
  int dc4_shift = -1;
  for (int k = 0; k < 10; ++k)
    if ((1 << k) == dc4) dc4_shift = k;
  const uint32_t sx_u = smem_u32(Sx), sd_u = smem_u32(Sd);

  for (long long item = blockIdx.x; item < p.total_items; item += gridDim.x) {
    const int n = (int)(item / p.strips_per_img);
    const int oh0 = (int)(item - (long long)n * p.strips_per_img) * p.R;
    // ---- stage the x strip (zero halo) and the dy strip (zero rows beyond the image / the last k-step)
    const int ih0 = oh0 * p.stride - p.pad;
    const int row_chunks = p.Wp * xc4;  // 16-byte chunks of one staged row (halo pixels included)
    for (int row = 0; row < p.in_rows; ++row) {
      const int ih = ih0 + row;
      const bool row_ok = ih >= 0 && ih < p.H;
      const float* grow = p.x + (((long long)n * p.H + (row_ok ? ih : 0)) * p.W - p.pad) * p.Cx;  // pixel col -> iw = col - pad
      const uint32_t srow = sx_u + (uint32_t)(row * p.Wp * p.xpitch) * 4u;
      for (int c = tid; c < row_chunks; c += blockDim.x) {
        const int col = xc4_shift >= 0 ? (c >> xc4_shift) : c / xc4, k4 = c - col * xc4;
        const int iw = col - p.pad;
        const bool ok = row_ok && iw >= 0 && iw < p.W;
        cp_async16(srow + (uint32_t)(col * p.xpitch + k4 * 4) * 4u, ok ? grow + (long long)col * p.Cx + k4 * 4 : p.x,
                   ok ? 16u : 0u);
      }
    }
    for (int c = tid; c < d_chunks; c += blockDim.x) {
      const int pix = dc4_shift >= 0 ? (c >> dc4_shift) : c / dc4, k4 = c - pix * dc4;
      const int ohl = p.ow_shift >= 0 ? (pix >> p.ow_shift) : pix / p.OW;
      const bool ok = pix < strip_px && oh0 + ohl < p.OH;
      const float* src = ok ? p.dy + (((long long)n * p.OH + oh0) * p.OW + pix) * p.Cout + k4 * 4 : p.dy;
      cp_async16(sd_u + (uint32_t)(pix * p.dpitch + k4 * 4) * 4u, src, ok ? 16u : 0u);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    if (active) {
      for (int step = q; step < nsteps; step += p.ps) {
        const int pa = step * 8 + t, pb = pa + 4;
        int oa, wa, ob, wb;
        if (p.ow_shift >= 0) {
          oa = pa >> p.ow_shift; wa = pa & (p.OW - 1);
          ob = pb >> p.ow_shift; wb = pb & (p.OW - 1);
        } else {
          oa = pa / p.OW; wa = pa - oa * p.OW;
          ob = pb / p.OW; wb = pb - ob * p.OW;
        }
        const int gx = PAIR ? 2 * g : g;
        const float* xa = Sx + ((oa * p.stride) * p.Wp + wa * p.stride) * p.xpitch + gx;
        const float* xb = Sx + ((ob * p.stride) * p.Wp + wb * p.stride) * p.xpitch + gx;
        const float2 a01 = *reinterpret_cast<const float2*>(Sd + pa * p.dpitch + co0 + 2 * g);
        const float2 a23 = *reinterpret_cast<const float2*>(Sd + pb * p.dpitch + co0 + 2 * g);
        uint32_t a[4];
        a[0] = f2tf32(a01.x);
        a[1] = f2tf32(a01.y);
        a[2] = f2tf32(a23.x);
        a[3] = f2tf32(a23.y);
        // operand loads are issued in batches ahead of their MMAs
        if (PAIR) {
          constexpr int NP = NT / 2;
          constexpr int BP = (NP % 3 == 0) ? 3 : ((NP % 2 == 0) ? 2 : 1);
#pragma unroll
          for (int j0 = 0; j0 < NP; j0 += BP) {
            float2 b0[BP], b1[BP];
#pragma unroll
            for (int j = 0; j < BP; ++j) {
              b0[j] = *reinterpret_cast<const float2*>(xa + toff[2 * (j0 + j)]);
              b1[j] = *reinterpret_cast<const float2*>(xb + toff[2 * (j0 + j)]);
            }
#pragma unroll
            for (int j = 0; j < BP; ++j) {
              mma_tf32(acc[2 * (j0 + j)], a, f2tf32(b0[j].x), f2tf32(b1[j].x));
              mma_tf32(acc[2 * (j0 + j) + 1], a, f2tf32(b0[j].y), f2tf32(b1[j].y));
            }
          }
        } else {
          constexpr int BT = (NT % 6 == 0) ? 6 : ((NT % 4 == 0) ? 4 : 2);
#pragma unroll
          for (int i0 = 0; i0 < NT; i0 += BT) {
            float b0[BT], b1[BT];
#pragma unroll
            for (int j = 0; j < BT; ++j) {
              b0[j] = xa[toff[i0 + j]];
              b1[j] = xb[toff[i0 + j]];
            }
#pragma unroll
            for (int j = 0; j < BT; ++j) mma_tf32(acc[i0 + j], a, f2tf32(b0[j]), f2tf32(b1[j]));
          }
        }
      }
    }
    __syncthreads();  // the strip buffers are rewritten by the next item
  }
  // ---- the ps warps of a group hold partial sums of the same tiles: add them through shared memory in warp order
  // (fixed order: deterministic), then ONE slice per CTA goes to the workspace with plain stores
  float* red = reinterpret_cast<float*>(smem_raw);  // [warp][tile][lane][4]; the strip buffers are free now
  if (p.ps > 1) {
    if (active) {
#pragma unroll
      for (int i = 0; i < NT; ++i)
        if (i < p.tpw)
          *reinterpret_cast<float4*>(red + (((size_t)warp * p.tpw + i) * 32 + lane) * 4) =
              make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
    __syncthreads();
    if (active && q == 0) {
#pragma unroll
      for (int i = 0; i < NT; ++i) {
        if (i >= p.tpw) continue;
        for (int k = 1; k < p.ps; ++k) {
          const float4 v = *reinterpret_cast<const float4*>(red + (((size_t)(warp + k) * p.tpw + i) * 32 + lane) * 4);
          acc[i][0] += v.x; acc[i][1] += v.y; acc[i][2] += v.z; acc[i][3] += v.w;
        }
      }
    }
  }
  if (!active || q != 0) return;
  const int taps = p.KH * p.KW;
  float* dst = p.part + (long long)blockIdx.x * p.Cout * taps * p.Cx;
#pragma unroll
  for (int i = 0; i < NT; ++i) {
    if (i >= p.tpw) continue;
    const int id = tile0 + i;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int co = co0 + 2 * g + (e >> 1);  // (row permutation of the dy^T fragment, see the kernel header)
      const int nn = 2 * t + (e & 1);
      int tap, ci;
      if (c4) {
        const int r = id / p.ncc, s = 2 * (id - r * p.ncc) + (nn >> 2);
        if (s >= p.KW) continue;
        tap = r * p.KW + s;
        ci = nn & 3;
      } else {
        tap = id / p.ncc;
        const int cc = id - tap * p.ncc;
        ci = PAIR ? (cc >> 1) * 16 + 2 * nn + (cc & 1) : cc * 8 + nn;
      }
      dst[((long long)co * taps + tap) * p.Cx + ci] = acc[i][e];
    }
  }
}

// dw[co][ci][tap] (OIHW) = sum over slices.  A block owns 64 consecutive elements of the partial layout
// [co][tap][Cx] (coalesced reads), its 8 thread rows take every 8th slice and the rows are added in row order:
// the summation order is fixed by the launch shape alone (deterministic).
__global__ void __launch_bounds__(512) wgrad_reduce_kernel(const float* __restrict__ part, int slices, int Cout, int taps,
                                                          int Cx, int Cw, float* dw) {
  __shared__ float sm[8][64];
  const long long per = (long long)Cout * taps * Cx;
  const long long e = (long long)blockIdx.x * 64 + threadIdx.x;
  float s = 0.f;
  if (e < per)
    for (int k = threadIdx.y; k < slices; k += 8) s += part[k * per + e];
  sm[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && e < per) {
    for (int k = 1; k < 8; ++k) s += sm[k][threadIdx.x];
    const int ci = (int)(e % Cx);
    const int tap = (int)((e / Cx) % taps);
    const int co = (int)(e / ((long long)Cx * taps));
    if (ci < Cw) dw[((long long)co * Cw + ci) * taps + tap] = s;
  }
}

struct WgPlan {
  WgArgs a;
  int threads, grid_x, grid_y;
  size_t smem, ws_floats;
};

int plan_wgrad(int N, int H, int W, int Cx, int Cout, int KH, int KW, int stride, int pad, WgPlan& pl) {
  WgArgs& a = pl.a;
  a = WgArgs{};
  if ((Cout % 16) || !(Cx == 4 || (Cx % 8) == 0) || stride < 1 || stride > 2 || Cx > 512 || Cout > 512) return AVL_ERR_UNSUPPORTED;
  a.N = N; a.H = H; a.W = W; a.Cx = Cx; a.Cout = Cout; a.KH = KH; a.KW = KW; a.stride = stride; a.pad = pad;
  a.OH = (H + 2 * pad - KH) / stride + 1;
  a.OW = (W + 2 * pad - KW) / stride + 1;
  if (a.OH < 1 || a.OW < 1) return AVL_ERR_ARG;
  const bool c4 = Cx == 4;
  a.ncc = c4 ? (KW + 1) / 2 : Cx / 8;
  a.ntiles_ct = c4 ? KH * a.ncc : KH * KW * a.ncc;
  a.tpw = 1;
  for (int d = 1; d <= WG_MAX_TILES; ++d)
    if (a.ntiles_ct % d == 0) a.tpw = d;
  a.total_groups = (Cout / 16) * (a.ntiles_ct / a.tpw);
  auto pitch = [](int c) {  // smallest pitch >= c (multiple of 4 floats) that is 8 or 24 mod 32; c == 4 stays 4
    if (c == 4) return 4;
    int pch = c;
    while ((pch & 31) != 8 && (pch & 31) != 24) pch += 4;
    return pch;
  };
  a.xpitch = pitch(Cx);
  a.dpitch = pitch(Cout);
  if (a.total_groups <= 8) {
    a.groups = a.total_groups;
    a.ps = 8 / a.groups;
    pl.threads = a.groups * a.ps * 32;
    pl.grid_y = 1;
  } else {
    a.groups = a.total_groups < 16 ? a.total_groups : 16;
    a.ps = 1;
    pl.threads = a.groups * 32;
    pl.grid_y = avl_div_up(a.total_groups, 16);
  }
  a.group0_stride = a.groups;
  a.Wp = (a.OW - 1) * stride + KW + (c4 ? 1 : 0);  // (Cx == 4: the last pair tile reads one pixel beyond the kernel row)
  const size_t budget = pl.threads <= 256 ? 100 * 1024 : 200 * 1024;
  // strip height: the largest strips that fit cost the fewest halo rows; among the heights that fit pick the one
  // that stages the fewest input rows per image (strips * in_rows), ties to the taller strip
  int best = 0;
  long long best_rows = 0;
  for (int R = 1; R <= a.OH; ++R) {
    const int in_rows = (R - 1) * stride + KH;
    const size_t bytes = ((size_t)in_rows * a.Wp * a.xpitch + (size_t)((R * a.OW + 7) / 8 * 8) * a.dpitch + 64) * 4;
    if (bytes > budget) break;
    const long long rows = (long long)avl_div_up(a.OH, R) * (in_rows + R);  // staged x rows + dy rows per image
    if (best == 0 || rows <= best_rows) { best = R; best_rows = rows; }
  }
  if (best == 0) return AVL_ERR_UNSUPPORTED;
  a.R = best;
  a.in_rows = (a.R - 1) * stride + KH;
  a.strips_per_img = avl_div_up(a.OH, a.R);
  a.total_items = (long long)N * a.strips_per_img;
  a.ow_shift = -1;
  for (int s = 0; s < 16; ++s)
    if ((1 << s) == a.OW) a.ow_shift = s;
  pl.smem = ((size_t)a.in_rows * a.Wp * a.xpitch + (size_t)((a.R * a.OW + 7) / 8 * 8) * a.dpitch + 64) * 4;
  const int per_sm = pl.threads <= 256 ? 2 : 1;
  long long gx = (long long)avl_num_sms() * per_sm / pl.grid_y;
  if (gx < 1) gx = 1;
  if (gx > a.total_items) gx = a.total_items;
  pl.grid_x = (int)gx;
  pl.ws_floats = (size_t)pl.grid_x * Cout * KH * KW * Cx;
  const size_t red_bytes = a.ps > 1 ? (size_t)(pl.threads / 32) * a.tpw * 32 * 4 * sizeof(float) : 0;
  if (pl.smem < red_bytes) pl.smem = red_bytes;
  return AVL_OK;
}

}  // namespace

AVL_API int avl_pack_conv_weight(const float* w_oihw, int Cout, int C, int KH, int KW, int pad_to, int mode, float* out,
                                 void* stream) {
  if (!w_oihw || !out || Cout < 1 || C < 1 || KH < 1 || KW < 1 || (mode != 0 && mode != 1)) return AVL_ERR_ARG;
  if (pad_to < (mode ? Cout : C)) return AVL_ERR_ARG;
  const long long n = (long long)(mode ? C : Cout) * KH * KW * pad_to;
  pack_conv_weight_kernel<<<avl_div_up(n, 256) > 1024 ? 1024 : avl_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(
      w_oihw, out, Cout, C, KH, KW, pad_to, mode);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}

AVL_API int avl_zero_upsample2(const float* dy, float* up, int N, int OH, int OW, int C, int H, int W, void* stream) {
  if (!dy || !up || N < 0 || OH < 1 || OW < 1 || C < 4 || (C & 3) || H < 1 || W < 1) return AVL_ERR_ARG;
  if (N == 0) return AVL_OK;
  if (((uintptr_t)dy & 15) || ((uintptr_t)up & 15)) return AVL_ERR_ARG;
  const long long n = (long long)N * H * W * (C / 4);
  const int blocks = avl_div_up(n, 256) > 148 * 16 ? 148 * 16 : avl_div_up(n, 256);
  zero_upsample2_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(dy),
                                                                 reinterpret_cast<float4*>(up), N, OH, OW, H, W, C / 4);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}

// Workspace floats needed by avl_tc_conv2d_wgrad for this shape (-2: shape not covered -> use avl_conv2d_wgrad).
AVL_API long long avl_tc_conv2d_wgrad_workspace(int N, int H, int W, int Cx, int Cout, int KH, int KW, int stride, int pad) {
  WgPlan pl;
  const int rc = plan_wgrad(N, H, W, Cx, Cout, KH, KW, stride, pad, pl);
  if (rc) return rc;
  return (long long)pl.ws_floats;
}

// dw (Cout, Cw, KH, KW) = sum_{n,oh,ow} dy (N,OH,OW,Cout) x x (N,H,W,Cx), Cx >= Cw (zero-padded input channels are
// dropped).  TF32 tensor-core products (operands rounded to nearest), fp32 accumulation, deterministic.
AVL_API int avl_tc_conv2d_wgrad(const float* x, const float* dy, float* dw, int N, int H, int W, int Cx, int Cw, int Cout,
                                int KH, int KW, int stride, int pad, float* workspace, long long ws_floats, void* stream) {
  if (!x || !dy || !dw || !workspace || N < 0 || Cw < 1 || Cw > Cx) return AVL_ERR_ARG;
  if (((uintptr_t)x & 15) || ((uintptr_t)dy & 15)) return AVL_ERR_UNSUPPORTED;
  WgPlan pl;
  const int rc = plan_wgrad(N, H, W, Cx, Cout, KH, KW, stride, pad, pl);
  if (rc) return rc;
  if ((long long)pl.ws_floats > ws_floats) return AVL_ERR_ARG;
  if (N == 0) {
    AVL_CUDA_CHECK(cudaMemsetAsync(dw, 0, sizeof(float) * Cout * Cw * KH * KW, (cudaStream_t)stream));
    return AVL_OK;
  }
  const dim3 grid(pl.grid_x, pl.grid_y);
  const cudaStream_t cs = (cudaStream_t)stream;
  const int tpw = pl.a.tpw;
  const bool pair = Cx != 4 && (pl.a.ncc % 2) == 0 && (tpw % 2) == 0;
  pl.a.x = x;
  pl.a.dy = dy;
  pl.a.part = workspace;
  auto launch = [&](auto kern) -> int {
    AVL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + 256));
    kern<<<grid, pl.threads, pl.smem, cs>>>(pl.a);
    return AVL_OK;
  };
  int lrc;
  if (pair) {
    if (tpw > 16) lrc = launch(tc_conv_wgrad_kernel<18, true>);
    else if (tpw > 8) lrc = launch(tc_conv_wgrad_kernel<16, true>);
    else if (tpw > 4) lrc = launch(tc_conv_wgrad_kernel<8, true>);
    else if (tpw > 2) lrc = launch(tc_conv_wgrad_kernel<4, true>);
    else lrc = launch(tc_conv_wgrad_kernel<2, true>);
  } else {
    if (tpw > 16) lrc = launch(tc_conv_wgrad_kernel<18, false>);
    else if (tpw > 8) lrc = launch(tc_conv_wgrad_kernel<16, false>);
    else if (tpw > 4) lrc = launch(tc_conv_wgrad_kernel<8, false>);
    else if (tpw > 2) lrc = launch(tc_conv_wgrad_kernel<4, false>);
    else lrc = launch(tc_conv_wgrad_kernel<2, false>);
  }
  if (lrc) return lrc;
  AVL_LAUNCH_CHECK();
  const long long per = (long long)Cout * KH * KW * Cx;
  wgrad_reduce_kernel<<<avl_div_up(per, 64), dim3(64, 8), 0, (cudaStream_t)stream>>>(workspace, pl.grid_x, Cout, KH * KW, Cx,
                                                                                    Cw, dw);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}
#endif  // AVL_HOST_EMUL
