// Whole-network inference of the two ResNet-18 variants on the hot path in ONE C-ABI call:
//   * custom_resnet18 (ss_baselines/savi/models/smt_resnet.py:56-164): 7x7 stride-1 stem, no max-pool, widths
//     16/32/64/128, GroupNorm(16), FC over the flattened map — the SMTCNN rgb / depth encoders (smt_cnn.py:78-115)
//     and the belief predictor's location head (belief_predictor.py:64-72);
//   * torchvision resnet18 (belief_predictor.py:74-82): 7x7 stride-2 stem, max-pool, widths 64..512, eval-mode
//     BatchNorm folded into the conv epilogue, global average pool + Linear.
// A rollout step at 64 envs is a chain of ~230 small kernels; issued one ctypes call at a time from Python the host
// (≈17 us per launch) was as slow as the device (tools/host_time.py: 4.0 ms issue vs 4.3 ms per step).  Here the
// ~45 launches of one network are enqueued from C++ behind a single call, and two independent networks (rgb and
// depth encoder; classifier and location predictor) can be enqueued on two streams that fork from / join into the
// caller's stream with events, so their small grids share the GPU.
#include "common.cuh"
#include "../../include/avlen_b200.h"

#ifndef AVL_HOST_EMUL
#include <cuda_fp16.h>
#include <string.h>

// typed entries of the halo-strip convolution and the cluster GroupNorm (fp16 activation storage)
int avl_tc_conv_halo_typed(const void* x, int in16, int N, int H, int W, int C, const void* w_packed, int Cout, int KH,
                           int KW, int stride, int pad, const float* scale, const float* bias, const float* residual,
                           long long ldr, int relu, void* y, int out16, long long ldy, cudaStream_t stream);
int avl_groupnorm_cluster_typed(const void* x, int in16, const float* gamma, const float* beta, const void* residual,
                                void* y, int out16, int N, int HW, int C, int groups, float eps, int relu,
                                void* stream);
long long avl_launch_count_internal();

namespace {

int g_f16_act = 1;

// 3x3 weights of stages 1 / 2 (already rounded to TF32, so the conversion is exact) as fp16, one launch
struct HalfPack {
  const float* src[10];
  __half* dst[10];
  int n[10];
};
__global__ void pack_half_kernel(HalfPack p) {
  const float* s = p.src[blockIdx.y];
  __half* d = p.dst[blockIdx.y];
  const int n = p.n[blockIdx.y];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) d[i] = __float2half_rn(s[i]);
}

struct Act {
  float* p;
  int H, W, C;
};

inline int out_dim(int size, int k, int stride, int pad) { return (size + 2 * pad - k) / stride + 1; }

// params table layout (device pointers; host array):
//   [0..2]   stem: weight, norm a, norm b            (GroupNorm: gamma, beta; folded BatchNorm: scale, bias)
//   [3 + 9*i ..] block i = 2*layer + b: w1, a1, b1, w2, a2, b2, wd, ad, bd   (wd == NULL: identity shortcut)
//   [75, 76] head: weight, bias
// Conv weights: (Cout, KH, KW, Cin) when use_tc != 0 (tensor-core path), the reference's OIHW otherwise.
constexpr int RN_STEM = 0, RN_BLOCK0 = 3, RN_PER_BLOCK = 9, RN_HEAD = RN_BLOCK0 + 8 * RN_PER_BLOCK, RN_COUNT = RN_HEAD + 2;

int conv(const float* x, int N, int H, int W, int C, const float* w, int Cout, int K, int stride, int pad,
         const float* scale, const float* bias, const float* residual, int relu, float* y, long long ldy, int use_tc,
         void* s) {
  const long long ldr = residual ? Cout : 0;
  if (use_tc)
    return avl_tc_conv2d_fwd(x, N, H, W, C, w, Cout, K, K, stride, pad, scale, bias, residual, ldr, relu, y, ldy, s);
  return avl_conv2d_fwd(x, N, H, W, C, w, Cout, K, K, stride, pad, scale, bias, residual, ldr, relu, y, ldy, s);
}

int gnorm(float* x, const float* gamma, const float* beta, const float* residual, int N, int HW, int C, int groups,
          float eps, int relu, void* s) {
  int rc = avl_groupnorm_fwd_cluster(x, gamma, beta, residual, x, N, HW, C, groups, eps, relu, s);
  if (rc == AVL_ERR_UNSUPPORTED) rc = avl_groupnorm_fwd(x, gamma, beta, residual, x, N, HW, C, groups, eps, relu, s);
  return rc;
}

struct Net {
  int N, H, W, Cin, norm_kind, stem_k, stem_stride, stem_pad, stem_maxpool, widths[4], groups, head_kind, out_dim;
  float eps;
};

size_t max_act_floats(const Net& n) {
  const int h1 = out_dim(n.H, n.stem_k, n.stem_stride, n.stem_pad), w1 = out_dim(n.W, n.stem_k, n.stem_stride, n.stem_pad);
  return (size_t)n.N * h1 * w1 * n.widths[0];  // the stem output is the largest activation (later stages shrink 2x)
}

#define RN_TRY(expr)        \
  do {                      \
    int _rc = (expr);       \
    if (_rc) return _rc;    \
  } while (0)

// One stride-2 stage (BasicBlocks blk0, blk0 + 1) with fp16 storage of everything between the two stride-2
// convolutions (fp32 in, fp32 out: they run on the im2col kernel) and the stage output (fp32).  `cur` (fp32) lives in
// buf[ci]; the three other buffers are scratch; on return cur / ci describe the stage's fp32 output.  wh: the fp16
// 3x3 weights (blk0.conv2, blk0+1.conv1, blk0+1.conv2).  Buffer `act` floats hold the stem output, so an fp32 tensor of
// this stage is at most half of one and an fp16 tensor at most a quarter.
int stage_f16(const Net& n, const float* const* P, int blk0, Act& cur, int& ci, float* const* buf, size_t act,
              __half* const* wh, int use_tc, void* s) {
  const int C = cur.C, C2 = 2 * C;
  const int oh = out_dim(cur.H, 3, 2, 1), ow = out_dim(cur.W, 3, 2, 1), HW2 = oh * ow;
  const float* const* B2 = P + RN_BLOCK0 + blk0 * RN_PER_BLOCK;
  const float* const* B3 = B2 + RN_PER_BLOCK;
  float* fb = buf[(ci + 1) & 3];
  float* qb = buf[(ci + 2) & 3];
  float* ob = buf[(ci + 3) & 3];
  float* f32a = fb;
  float* f32b = fb + act / 2;
  void* q1 = qb;
  void* q2 = qb + act / 4;
  void* q3 = qb + act / 2;
  void* q4 = qb + 3 * (act / 4);
  cudaStream_t cs = (cudaStream_t)s;
  RN_TRY(conv(cur.p, n.N, cur.H, cur.W, C, B2[0], C2, 3, 2, 1, nullptr, nullptr, nullptr, 0, f32a, C2, use_tc, s));
  RN_TRY(conv(cur.p, n.N, cur.H, cur.W, C, B2[6], C2, 1, 2, 0, nullptr, nullptr, nullptr, 0, f32b, C2, use_tc, s));
  RN_TRY(avl_groupnorm_cluster_typed(f32a, 0, B2[1], B2[2], nullptr, q1, 1, n.N, HW2, C2, n.groups, n.eps, 1, s));
  RN_TRY(avl_groupnorm_cluster_typed(f32b, 0, B2[7], B2[8], nullptr, q2, 1, n.N, HW2, C2, n.groups, n.eps, 0, s));
  RN_TRY(avl_tc_conv_halo_typed(q1, 1, n.N, oh, ow, C2, wh[0], C2, 3, 3, 1, 1, nullptr, nullptr, nullptr, 0, 0, q3, 1, C2, cs));
  RN_TRY(avl_groupnorm_cluster_typed(q3, 1, B2[4], B2[5], q2, q3, 1, n.N, HW2, C2, n.groups, n.eps, 1, s));
  RN_TRY(avl_tc_conv_halo_typed(q3, 1, n.N, oh, ow, C2, wh[1], C2, 3, 3, 1, 1, nullptr, nullptr, nullptr, 0, 0, q1, 1, C2, cs));
  RN_TRY(avl_groupnorm_cluster_typed(q1, 1, B3[1], B3[2], nullptr, q1, 1, n.N, HW2, C2, n.groups, n.eps, 1, s));
  RN_TRY(avl_tc_conv_halo_typed(q1, 1, n.N, oh, ow, C2, wh[2], C2, 3, 3, 1, 1, nullptr, nullptr, nullptr, 0, 0, q4, 1, C2, cs));
  RN_TRY(avl_groupnorm_cluster_typed(q4, 1, B3[4], B3[5], q3, ob, 0, n.N, HW2, C2, n.groups, n.eps, 1, s));
  cur = {ob, oh, ow, C2};
  ci = (ci + 3) & 3;
  return AVL_OK;
}

int run(const Net& n, const float* x, const float* const* P, float* out, long long ldo, int use_tc, float* ws,
        void* s) {
  const size_t act = (max_act_floats(n) + 63) & ~(size_t)63;
  float* buf[4] = {ws, ws + act, ws + 2 * act, ws + 3 * act};
  const bool gn = n.norm_kind == 0;
  // ---- stem
  Act cur = {buf[0], out_dim(n.H, n.stem_k, n.stem_stride, n.stem_pad), out_dim(n.W, n.stem_k, n.stem_stride, n.stem_pad),
             n.widths[0]};
  // fp16 activation storage for the stem output and stage 1 (the widest tensors): GroupNorm networks on the
  // tensor-core path whose stem and stage-1 convolutions fit the halo-strip kernel and the cluster GroupNorm
  bool f16_path = false;
  if (gn && use_tc && g_f16_act && !n.stem_maxpool && n.stem_stride == 1 && n.stem_pad == n.stem_k / 2 &&
      (n.widths[0] % 16) == 0 && n.widths[0] <= 64 && cur.W >= 16 && (long long)cur.H * cur.W * cur.C * 4 <= 8 * 32 * 1024 &&
      cur.H * cur.W >= 8) {
    int rc = avl_tc_conv_halo_typed(x, 0, n.N, n.H, n.W, n.Cin, P[RN_STEM], cur.C, n.stem_k, n.stem_k, 1, n.stem_pad, nullptr,
                                    nullptr, nullptr, 0, 0, buf[0], 1, cur.C, (cudaStream_t)s);
    if (rc == AVL_OK) f16_path = true;
    else if (rc != AVL_ERR_UNSUPPORTED) return rc;
  }
  if (f16_path) {
    // GroupNorm + stage 1 follow below (fp16 buffers)
  } else if (gn) {
    RN_TRY(conv(x, n.N, n.H, n.W, n.Cin, P[RN_STEM], cur.C, n.stem_k, n.stem_stride, n.stem_pad, nullptr, nullptr,
                nullptr, 0, cur.p, cur.C, use_tc, s));
    RN_TRY(gnorm(cur.p, P[RN_STEM + 1], P[RN_STEM + 2], nullptr, n.N, cur.H * cur.W, cur.C, n.groups, n.eps, 1, s));
  } else {
    RN_TRY(conv(x, n.N, n.H, n.W, n.Cin, P[RN_STEM], cur.C, n.stem_k, n.stem_stride, n.stem_pad, P[RN_STEM + 1],
                P[RN_STEM + 2], nullptr, 1, cur.p, cur.C, use_tc, s));
  }
  int ci = 0;  // index of the buffer holding `cur`
  int first_blk = 0;
  if (f16_path) {
    // ---- stage 1 on fp16 activations: t(fp16) in buf[0..2], the stage's output leaves as fp32 in buf[3]
    const int HW = cur.H * cur.W, C = cur.C;
    const int oh2 = out_dim(cur.H, 3, 2, 1), ow2 = out_dim(cur.W, 3, 2, 1), C2 = n.widths[1];
    const bool stage2 = ow2 >= 16 && (C2 % 16) == 0 && C2 <= 128 && C2 == 2 * C && (oh2 * ow2) >= 8 &&
                        (long long)oh2 * ow2 * C2 <= (long long)HW * C / 2 + 0 && (cur.H % 2) == 0 && (cur.W % 2) == 0;
    const int oh3 = out_dim(oh2, 3, 2, 1), ow3 = out_dim(ow2, 3, 2, 1), C3 = n.widths[2];
    const bool stage3 = stage2 && ow3 >= 16 && (C3 % 16) == 0 && C3 <= 64 && C3 == 2 * C2 && (oh3 * ow3) >= 8 &&
                        (oh2 % 2) == 0 && (ow2 % 2) == 0;
    HalfPack hp = {};
    __half* wh = reinterpret_cast<__half*>(reinterpret_cast<unsigned char*>(ws + 4 * act) + 256);
    const int n1 = C * 9 * C, n2 = C2 * 9 * C2, n3 = C3 * 9 * C3;
    for (int i = 0; i < 4; ++i) {
      hp.src[i] = P[RN_BLOCK0 + (i >> 1) * RN_PER_BLOCK + (i & 1) * 3];
      hp.dst[i] = wh + (size_t)i * n1;
      hp.n[i] = n1;
    }
    const int pack_idx2[3] = {RN_BLOCK0 + 2 * RN_PER_BLOCK + 3, RN_BLOCK0 + 3 * RN_PER_BLOCK, RN_BLOCK0 + 3 * RN_PER_BLOCK + 3};
    for (int i = 0; i < 3; ++i) {
      hp.src[4 + i] = P[pack_idx2[i]];
      hp.dst[4 + i] = wh + (size_t)4 * n1 + (size_t)i * n2;
      hp.n[4 + i] = n2;
    }
    for (int i = 0; i < 3; ++i) {
      hp.src[7 + i] = P[pack_idx2[i] + 2 * RN_PER_BLOCK];
      hp.dst[7 + i] = wh + (size_t)4 * n1 + (size_t)3 * n2 + (size_t)i * n3;
      hp.n[7 + i] = n3;
    }
    pack_half_kernel<<<dim3(avl_div_up(stage3 ? n3 : (stage2 ? n2 : n1), 256), stage3 ? 10 : (stage2 ? 7 : 4)), 256, 0,
                       (cudaStream_t)s>>>(hp);
    AVL_LAUNCH_CHECK();
    void* h0 = buf[0];
    void* h1 = buf[1];
    void* h2 = buf[2];
    RN_TRY(avl_groupnorm_cluster_typed(h0, 1, P[RN_STEM + 1], P[RN_STEM + 2], nullptr, h0, 1, n.N, HW, C, n.groups, n.eps,
                                       1, s));
    for (int blk = 0; blk < 2; ++blk) {
      const float* const* B = P + RN_BLOCK0 + blk * RN_PER_BLOCK;
      RN_TRY(avl_tc_conv_halo_typed(h0, 1, n.N, cur.H, cur.W, C, hp.dst[2 * blk], C, 3, 3, 1, 1, nullptr, nullptr, nullptr,
                                    0, 0, h1, 1, C, (cudaStream_t)s));
      RN_TRY(avl_groupnorm_cluster_typed(h1, 1, B[1], B[2], nullptr, h1, 1, n.N, HW, C, n.groups, n.eps, 1, s));
      RN_TRY(avl_tc_conv_halo_typed(h1, 1, n.N, cur.H, cur.W, C, hp.dst[2 * blk + 1], C, 3, 3, 1, 1, nullptr, nullptr,
                                    nullptr, 0, 0, h2, 1, C, (cudaStream_t)s));
      if (blk == 0) {
        RN_TRY(avl_groupnorm_cluster_typed(h2, 1, B[4], B[5], h0, h2, 1, n.N, HW, C, n.groups, n.eps, 1, s));
        void* t = h0; h0 = h2; h2 = t;  // block output becomes the next block's input / identity
      } else {
        RN_TRY(avl_groupnorm_cluster_typed(h2, 1, B[4], B[5], h0, buf[3], 0, n.N, HW, C, n.groups, n.eps, 1, s));
      }
    }
    cur.p = buf[3];
    ci = 3;
    first_blk = 2;
    // ---- stages 2 and 3: the two stride-2 convolutions read the previous stage's fp32 output through the im2col
    // kernel and write fp32; everything between them and the stage's output is fp16
    for (int st = 1; st <= 2 && first_blk == 2 * st; ++st) {
      if (!(st == 1 ? stage2 : stage3)) break;
      RN_TRY(stage_f16(n, P, 2 * st, cur, ci, buf, act, hp.dst + (st == 1 ? 4 : 7), use_tc, s));
      first_blk = 2 * st + 2;
    }
  }
  if (n.stem_maxpool) {
    Act nx = {buf[1], out_dim(cur.H, 3, 2, 1), out_dim(cur.W, 3, 2, 1), cur.C};
    RN_TRY(avl_maxpool3x3s2(cur.p, nx.p, n.N, cur.H, cur.W, cur.C, s));
    cur = nx;
    ci = 1;
  }
  // ---- 4 stages x 2 BasicBlocks
  for (int blk = first_blk; blk < 8; ++blk) {
    const float* const* B = P + RN_BLOCK0 + blk * RN_PER_BLOCK;
    const int planes = n.widths[blk >> 1];
    const int stride = ((blk & 1) == 0 && blk > 0) ? 2 : 1;
    float* t1 = buf[(ci + 1) & 3];
    float* t2 = buf[(ci + 2) & 3];
    float* idb = buf[(ci + 3) & 3];
    const int oh = out_dim(cur.H, 3, stride, 1), ow = out_dim(cur.W, 3, stride, 1);
    const float* identity = cur.p;
    if (gn) {
      RN_TRY(conv(cur.p, n.N, cur.H, cur.W, cur.C, B[0], planes, 3, stride, 1, nullptr, nullptr, nullptr, 0, t1, planes,
                  use_tc, s));
      RN_TRY(gnorm(t1, B[1], B[2], nullptr, n.N, oh * ow, planes, n.groups, n.eps, 1, s));
      RN_TRY(conv(t1, n.N, oh, ow, planes, B[3], planes, 3, 1, 1, nullptr, nullptr, nullptr, 0, t2, planes, use_tc, s));
      if (B[6]) {
        RN_TRY(conv(cur.p, n.N, cur.H, cur.W, cur.C, B[6], planes, 1, stride, 0, nullptr, nullptr, nullptr, 0, idb,
                    planes, use_tc, s));
        RN_TRY(gnorm(idb, B[7], B[8], nullptr, n.N, oh * ow, planes, n.groups, n.eps, 0, s));
        identity = idb;
      }
      RN_TRY(gnorm(t2, B[4], B[5], identity, n.N, oh * ow, planes, n.groups, n.eps, 1, s));
    } else {
      RN_TRY(conv(cur.p, n.N, cur.H, cur.W, cur.C, B[0], planes, 3, stride, 1, B[1], B[2], nullptr, 1, t1, planes,
                  use_tc, s));
      if (B[6]) {
        RN_TRY(conv(cur.p, n.N, cur.H, cur.W, cur.C, B[6], planes, 1, stride, 0, B[7], B[8], nullptr, 0, idb, planes,
                    use_tc, s));
        identity = idb;
      }
      RN_TRY(conv(t1, n.N, oh, ow, planes, B[3], planes, 3, 1, 1, B[4], B[5], identity, 1, t2, planes, use_tc, s));
    }
    cur = {t2, oh, ow, planes};
    ci = (ci + 2) & 3;
  }
  // ---- head
  if (n.head_kind == 0) {  // Linear over the NCHW-flattened map == a conv whose kernel covers the whole map
    if (cur.H != cur.W && use_tc == 0) {
      // SIMT path takes KH, KW separately through avl_conv2d_fwd
    }
    const long long ldr = 0;
    if (use_tc)
      return avl_tc_conv2d_fwd(cur.p, n.N, cur.H, cur.W, cur.C, P[RN_HEAD], n.out_dim, cur.H, cur.W, 1, 0, nullptr,
                               P[RN_HEAD + 1], nullptr, ldr, 0, out, ldo, s);
    return avl_conv2d_fwd(cur.p, n.N, cur.H, cur.W, cur.C, P[RN_HEAD], n.out_dim, cur.H, cur.W, 1, 0, nullptr,
                          P[RN_HEAD + 1], nullptr, ldr, 0, out, ldo, s);
  }
  float* pooled = buf[(ci + 1) & 3];
  RN_TRY(avl_avgpool_global(cur.p, pooled, n.N, cur.H * cur.W, cur.C, s));
  return avl_gemm(pooled, cur.C, 1, P[RN_HEAD], cur.C, 1, out, ldo, n.N, n.out_dim, cur.C, P[RN_HEAD + 1], 0, 0, 1, s);
}

bool fill(Net& n, int N, int H, int W, int Cin, const int* cfg) {
  // cfg (host, 12 ints): norm_kind, stem_k, stem_stride, stem_pad, stem_maxpool, w0, w1, w2, w3, groups, head_kind, out_dim
  n.N = N; n.H = H; n.W = W; n.Cin = Cin;
  n.norm_kind = cfg[0]; n.stem_k = cfg[1]; n.stem_stride = cfg[2]; n.stem_pad = cfg[3]; n.stem_maxpool = cfg[4];
  for (int i = 0; i < 4; ++i) n.widths[i] = cfg[5 + i];
  n.groups = cfg[9]; n.head_kind = cfg[10]; n.out_dim = cfg[11];
  if (N < 0 || H < 1 || W < 1 || Cin < 1 || n.stem_k < 1 || n.stem_stride < 1 || n.out_dim < 1) return false;
  for (int i = 0; i < 4; ++i)
    if (n.widths[i] < 1) return false;
  return true;
}

// side stream + fork / join events PER CALLER STREAM: two pairs enqueued from two different streams (the visual
// encoders of step s+1 on the main stream, the belief networks on the trainer's belief stream) must not serialise
// on one shared side stream.
struct Side {
  cudaStream_t caller;
  cudaStream_t side;
  cudaEvent_t fork, join;
};
constexpr int MAX_SIDES = 16;
Side g_sides[MAX_SIDES];
int g_num_sides = 0;

int side_for(cudaStream_t s, Side** out) {
  for (int i = 0; i < g_num_sides; ++i)
    if (g_sides[i].caller == s) { *out = &g_sides[i]; return AVL_OK; }
  Side* sd = &g_sides[g_num_sides < MAX_SIDES ? g_num_sides : 0];  // more caller streams than slots: share slot 0
  if (g_num_sides < MAX_SIDES) {
    AVL_CUDA_CHECK(cudaStreamCreateWithFlags(&sd->side, cudaStreamNonBlocking));
    AVL_CUDA_CHECK(cudaEventCreateWithFlags(&sd->fork, cudaEventDisableTiming));
    AVL_CUDA_CHECK(cudaEventCreateWithFlags(&sd->join, cudaEventDisableTiming));
    sd->caller = s;
    ++g_num_sides;
  }
  *out = sd;
  return AVL_OK;
}


// ---------------------------------------------------------------------------------------------- CUDA graphs
// At rollout batch a network is a chain of ~50 small kernels: ~0.2 ms of host launch time per network and a launch
// gap between every dependent pair.  A whole-network call whose arguments (pointers, shapes, parameter table,
// library toggles) repeat is captured once into a CUDA graph — fork / join of the second network included — and
// replayed with one cudaGraphLaunch.  A key is captured the SECOND time it is seen (the first, direct run is the
// warm-up: lazy attribute / stream / driver-entry initialisation must not happen under capture, and one-off
// argument sets are never captured).  Only small batches use this (large ones are not launch-bound).
int g_graphs_on = 1;
constexpr int GR_MAX_BATCH = 512;
constexpr int GR_SLOTS = 48;

struct GraphKey {
  const void* x[2];
  const void* out[2];
  const void* ws[2];
  long long ldo[2];
  long long epoch;
  int N, H, W, cin[2], use_tc, pair;
  float eps;
  int cfg[2][12];
  const float* params[2][RN_COUNT];
};
struct GraphSlot {
  GraphKey key;
  cudaGraphExec_t exec;  // null: key seen once, not captured yet
  long long launches;
  unsigned long long last_use;
  bool used;
};
GraphSlot g_slots[GR_SLOTS];
unsigned long long g_graph_clock = 0;
long long g_graph_hits = 0, g_graph_captures = 0;

GraphSlot* graph_find(const GraphKey& k) {
  for (int i = 0; i < GR_SLOTS; ++i)
    if (g_slots[i].used && memcmp(&g_slots[i].key, &k, sizeof(GraphKey)) == 0) return &g_slots[i];
  return nullptr;
}
GraphSlot* graph_new(const GraphKey& k) {
  GraphSlot* v = &g_slots[0];
  for (int i = 0; i < GR_SLOTS; ++i) {
    if (!g_slots[i].used) { v = &g_slots[i]; break; }
    if (g_slots[i].last_use < v->last_use) v = &g_slots[i];
  }
  if (v->used && v->exec) cudaGraphExecDestroy(v->exec);
  v->key = k;
  v->exec = nullptr;
  v->launches = 0;
  v->used = true;
  return v;
}
// Returns true when the call was served (rc holds its status): replayed, or captured + launched.  false: the caller
// launches directly (first sighting of the key, or capture unavailable).
template <class F>
bool graph_call(const GraphKey& k, cudaStream_t s, int& rc, F&& body) {
  ++g_graph_clock;
  GraphSlot* slot = graph_find(k);
  if (!slot) {
    slot = graph_new(k);
    slot->last_use = g_graph_clock;
    return false;  // first sighting: direct run (warm-up)
  }
  slot->last_use = g_graph_clock;
  if (!slot->exec) {
    // capture on a private stream (the caller's may be the legacy default stream, which cannot be captured); the
    // instantiated graph is launched into the caller's stream
    static cudaStream_t cap = nullptr;
    if (!cap) {
      if (cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); g_graphs_on = 0; return false; }
      Side* warm = nullptr;
      if (side_for(cap, &warm) != AVL_OK) { g_graphs_on = 0; return false; }  // fork / join resources exist before capture
    }
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) { cudaGetLastError(); return false; }
    if (cudaStreamBeginCapture(cap, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
      cudaGetLastError();
      g_graphs_on = 0;
      return false;
    }
    const long long l0 = avl_launch_count_internal();
    const int brc = body(cap);
    const long long l1 = avl_launch_count_internal();
    cudaGraph_t graph = nullptr;
    const cudaError_t e = cudaStreamEndCapture(cap, &graph);
    if (brc != AVL_OK || e != cudaSuccess || !graph) {
      cudaGetLastError();
      if (graph) cudaGraphDestroy(graph);
      g_graphs_on = 0;  // something on this path cannot be captured: stay on direct launches for good
      avl_add_launches(-(l1 - l0));
      if (brc != AVL_OK) { rc = brc; return true; }
      return false;
    }
    cudaGraphExec_t exec = nullptr;
    if (cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess || !exec) {
      cudaGetLastError();
      cudaGraphDestroy(graph);
      g_graphs_on = 0;
      avl_add_launches(-(l1 - l0));
      return false;
    }
    cudaGraphDestroy(graph);
    slot->exec = exec;
    slot->launches = l1 - l0;
    avl_add_launches(-(l1 - l0));  // counted at capture, but nothing ran yet: the replay below counts them
    ++g_graph_captures;
  } else {
    ++g_graph_hits;
  }
  if (cudaGraphLaunch(slot->exec, s) != cudaSuccess) {
    avl_set_cuda_error((int)cudaGetLastError());
    rc = AVL_ERR_CUDA;
    return true;
  }
  avl_add_launches(slot->launches);
  rc = AVL_OK;
  return true;
}

}  // namespace

AVL_API int avl_resnet18_param_count(void) { return RN_COUNT; }

// 1 (default): the fused GroupNorm ResNet-18 keeps its stem output and stage 1 in HBM as fp16 (tensor-core path only);
// 0: fp32 activations everywhere.  Returns the previous setting.
AVL_API int avl_set_f16_activations(int on) {
  avl_bump_config_epoch();
  int old = g_f16_act;
  g_f16_act = on ? 1 : 0;
  return old;
}

// cfg: 12 host ints (see fill()).  Workspace for ONE network.
AVL_API long long avl_resnet18_workspace_bytes(int N, int H, int W, const int* cfg) {
  Net n;
  if (!cfg || !fill(n, N, H, W, 4, cfg)) return -1;
  // 4 activation buffers + the fp16 copies of stage 1's four 3x3 weights (fp16 activation path)
  return (long long)(4 * ((max_act_floats(n) + 63) & ~(size_t)63) * sizeof(float) + 256 +
                     (4 * (size_t)n.widths[0] * 9 * n.widths[0] + 3 * (size_t)n.widths[1] * 9 * n.widths[1] +
                      3 * (size_t)n.widths[2] * 9 * n.widths[2]) * sizeof(__half) + 256);
}

// x (N, H, W, Cin) NHWC (Cin % 4 == 0 on the tensor-core path); out (N, out_dim) rows of stride ldo.
// eps: GroupNorm epsilon.  use_tc: weights are packed (Cout, KH, KW, Cin) and every conv runs on tcgen05.
AVL_API int avl_resnet18_forward(const float* x, int N, int H, int W, int Cin, const int* cfg, float eps,
                                 const float* const* params, float* out, long long ldo, int use_tc, void* workspace,
                                 void* stream) {
  Net n;
  if (!cfg || !fill(n, N, H, W, Cin, cfg)) return AVL_ERR_ARG;
  if (N == 0) return AVL_OK;
  if (!x || !params || !out || !workspace) return AVL_ERR_ARG;
  n.eps = eps;
  if (g_graphs_on && N <= GR_MAX_BATCH) {
    GraphKey k;
    memset(&k, 0, sizeof(k));
    k.x[0] = x; k.out[0] = out; k.ws[0] = workspace; k.ldo[0] = ldo; k.epoch = avl_config_epoch();
    k.N = N; k.H = H; k.W = W; k.cin[0] = Cin; k.use_tc = use_tc; k.pair = 0; k.eps = eps;
    memcpy(k.cfg[0], cfg, sizeof(k.cfg[0]));
    memcpy(k.params[0], params, sizeof(k.params[0]));
    int rc = AVL_OK;
    if (graph_call(k, (cudaStream_t)stream, rc, [&](cudaStream_t cs) {
          return run(n, x, params, out, ldo, use_tc, static_cast<float*>(workspace), cs);
        }))
      return rc;
  }
  return run(n, x, params, out, ldo, use_tc, static_cast<float*>(workspace), stream);
}

// Two independent networks (e.g. the rgb and the depth encoder) enqueued concurrently: network 0 on `stream`,
// network 1 on an internal side stream that forks from `stream` and joins it again before the call returns control
// to the stream order.  Arguments as avl_resnet18_forward, one per network.
AVL_API int avl_resnet18_forward_pair(const float* x0, const float* x1, int N, int H, int W, int Cin0, int Cin1,
                                      const int* cfg0, const int* cfg1, float eps, const float* const* params0,
                                      const float* const* params1, float* out0, float* out1, long long ldo0,
                                      long long ldo1, int use_tc, void* workspace0, void* workspace1, void* stream) {
  Net a, b;
  if (!cfg0 || !cfg1 || !fill(a, N, H, W, Cin0, cfg0) || !fill(b, N, H, W, Cin1, cfg1)) return AVL_ERR_ARG;
  if (N == 0) return AVL_OK;
  if (!x0 || !x1 || !params0 || !params1 || !out0 || !out1 || !workspace0 || !workspace1) return AVL_ERR_ARG;
  a.eps = b.eps = eps;
  cudaStream_t s = (cudaStream_t)stream;
  auto body = [&](cudaStream_t cs) -> int {
    Side* sd = nullptr;
    int rcs = side_for(cs, &sd);
    if (rcs) return rcs;
    AVL_CUDA_CHECK(cudaEventRecord(sd->fork, cs));
    AVL_CUDA_CHECK(cudaStreamWaitEvent(sd->side, sd->fork, 0));
    int rc1 = run(b, x1, params1, out1, ldo1, use_tc, static_cast<float*>(workspace1), sd->side);
    int rc0 = run(a, x0, params0, out0, ldo0, use_tc, static_cast<float*>(workspace0), cs);
    AVL_CUDA_CHECK(cudaEventRecord(sd->join, sd->side));
    AVL_CUDA_CHECK(cudaStreamWaitEvent(cs, sd->join, 0));
    return rc0 ? rc0 : rc1;
  };
  if (g_graphs_on && N <= GR_MAX_BATCH) {
    GraphKey k;
    memset(&k, 0, sizeof(k));
    k.x[0] = x0; k.x[1] = x1; k.out[0] = out0; k.out[1] = out1; k.ws[0] = workspace0; k.ws[1] = workspace1;
    k.ldo[0] = ldo0; k.ldo[1] = ldo1; k.epoch = avl_config_epoch();
    k.N = N; k.H = H; k.W = W; k.cin[0] = Cin0; k.cin[1] = Cin1; k.use_tc = use_tc; k.pair = 1; k.eps = eps;
    memcpy(k.cfg[0], cfg0, sizeof(k.cfg[0]));
    memcpy(k.cfg[1], cfg1, sizeof(k.cfg[1]));
    memcpy(k.params[0], params0, sizeof(k.params[0]));
    memcpy(k.params[1], params1, sizeof(k.params[1]));
    int rc = AVL_OK;
    if (graph_call(k, s, rc, body)) return rc;
  }
  return body(s);
}

// 1 (default): whole-network calls at small batch replay a cached CUDA graph; 0: always launch kernel by kernel.
AVL_API int avl_set_resnet_graphs(int on) {
  int old = g_graphs_on;
  g_graphs_on = on ? 1 : 0;
  return old;
}
// hits (replays) and captures so far
AVL_API long long avl_resnet_graph_stats(int what) { return what == 0 ? g_graph_hits : g_graph_captures; }
#endif  // AVL_HOST_EMUL
