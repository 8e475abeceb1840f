// Rollout-storage / PPO kernels: GAE scan, advantages, categorical heads,
// fused PPO loss (forward + gradient), external-memory ring insert, belief EMA,
// global-norm clip + Adam.  SURVEY.md §8a rows G, I, M (scalar part), N, O, Q.
// All of these are tiny HBM/latency-bound integer/float kernels: the point is
// one launch instead of hundreds and no host synchronisation.
#include "common.cuh"

namespace {

// ------------------------------------------------------------------ GAE (row N)
// ss_baselines/savi/models/rollout_storage.py:394-412 and
// ss_baselines/common/rollout_storage.py:114-132.  One thread per env, the T
// steps are a sequential scan.  Multiplication order follows the reference
// expression so that results are bit-identical to the PyTorch CPU loop
// (no FMA contraction: explicit round-to-nearest intrinsics).
__global__ void gae_kernel(const float* __restrict__ rewards, float* value_preds, const float* __restrict__ masks,
                           const float* __restrict__ next_value, float* returns, int steps, int n, int use_gae,
                           float gamma, float gamma_tau) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  if (use_gae) {
    float v_next = next_value[e];
    value_preds[(size_t)steps * n + e] = v_next;
    float gae = 0.f;
    for (int t = steps - 1; t >= 0; --t) {
      float m = masks[(size_t)(t + 1) * n + e];
      float v = value_preds[(size_t)t * n + e];
      // delta = r + gamma * v[t+1] * m - v[t]
      float delta = __fsub_rn(__fadd_rn(rewards[(size_t)t * n + e], __fmul_rn(__fmul_rn(gamma, v_next), m)), v);
      // gae = delta + gamma * tau * m * gae
      gae = __fadd_rn(delta, __fmul_rn(__fmul_rn(gamma_tau, m), gae));
      returns[(size_t)t * n + e] = __fadd_rn(gae, v);
      v_next = v;
    }
  } else {
    float ret = next_value[e];
    returns[(size_t)steps * n + e] = ret;
    for (int t = steps - 1; t >= 0; --t) {
      float m = masks[(size_t)(t + 1) * n + e];
      ret = __fadd_rn(__fmul_rn(__fmul_rn(ret, gamma), m), rewards[(size_t)t * n + e]);
      returns[(size_t)t * n + e] = ret;
    }
  }
}

// ------------------------------------------------------------ advantages (row O)
// ppo.py:90-95: adv = returns[:-1] - value_preds[:-1]; optional (adv-mean)/(std+1e-5), unbiased std.
__global__ void advantages_kernel(const float* __restrict__ returns, const float* __restrict__ value_preds,
                                  float* adv, int count, int normalize, float eps) {
  __shared__ float red[33];
  float s = 0.f;
  for (int i = threadIdx.x; i < count; i += blockDim.x) {
    float a = returns[i] - value_preds[i];
    adv[i] = a;
    s += a;
  }
  if (!normalize) return;
  float mean = block_sum(s, red) / (float)count;
  float q = 0.f;
  for (int i = threadIdx.x; i < count; i += blockDim.x) {
    float d = adv[i] - mean;
    q += d * d;
  }
  float var = block_sum(q, red) / (float)(count - 1);
  float inv = 1.f / (sqrtf(var) + eps);
  for (int i = threadIdx.x; i < count; i += blockDim.x) adv[i] = (adv[i] - mean) * inv;
}

// ------------------------------------------------- categorical heads (row I)
constexpr int kMaxA = 32;

__device__ __forceinline__ float row_lse(const float* z, int A, float& mx, float& sum) {
  mx = z[0];
  for (int j = 1; j < A; ++j) mx = fmaxf(mx, z[j]);
  sum = 0.f;
  for (int j = 0; j < A; ++j) sum += expf(z[j] - mx);
  return mx + logf(sum);
}

// common/utils.py:44-72.  uniforms == null -> mode() (first arg-max);
// otherwise inverse-CDF sampling on the supplied uniforms (the oracle does the same).
__global__ void categorical_act_kernel(const float* __restrict__ logits, const float* __restrict__ uniforms, int B,
                                       int A, long long* actions, float* log_probs, float* probs) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float z[kMaxA];
  for (int j = 0; j < A; ++j) z[j] = logits[(size_t)b * A + j];
  float mx, sum;
  float lse = row_lse(z, A, mx, sum);
  int a = 0;
  if (uniforms == nullptr) {
    for (int j = 1; j < A; ++j)
      if (z[j] > z[a]) a = j;
  } else {
    float u = uniforms[b], c = 0.f;
    a = A - 1;
    for (int j = 0; j < A; ++j) {
      c += expf(z[j] - mx) / sum;
      if (c > u) { a = j; break; }
    }
  }
  actions[b] = a;
  log_probs[b] = z[a] - lse;
  if (probs)
    for (int j = 0; j < A; ++j) probs[(size_t)b * A + j] = expf(z[j] - mx) / sum;
}

// log_probs(action) and per-row entropy (+ probs) for evaluate_actions*.
__global__ void categorical_eval_kernel(const float* __restrict__ logits, const long long* __restrict__ actions, int B,
                                        int A, float* log_probs, float* entropy, float* probs) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float z[kMaxA];
  for (int j = 0; j < A; ++j) z[j] = logits[(size_t)b * A + j];
  float mx, sum;
  float lse = row_lse(z, A, mx, sum);
  float h = 0.f;
  for (int j = 0; j < A; ++j) {
    float lp = z[j] - lse, p = expf(z[j] - mx) / sum;
    h -= p * lp;
    if (probs) probs[(size_t)b * A + j] = p;
  }
  int a = (int)actions[b];
  log_probs[b] = z[a] - lse;
  entropy[b] = h;
}

// dlogits = g_lp[b] * (onehot - p) + g_ent[b] * (-p (lp + H))
__global__ void categorical_eval_bwd_kernel(const float* __restrict__ logits, const long long* __restrict__ actions,
                                            const float* __restrict__ g_lp, const float* __restrict__ g_ent, int B,
                                            int A, float* dlogits) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float z[kMaxA];
  for (int j = 0; j < A; ++j) z[j] = logits[(size_t)b * A + j];
  float mx, sum;
  float lse = row_lse(z, A, mx, sum);
  float h = 0.f;
  for (int j = 0; j < A; ++j) h -= (expf(z[j] - mx) / sum) * (z[j] - lse);
  int a = (int)actions[b];
  float gl = g_lp ? g_lp[b] : 0.f, ge = g_ent ? g_ent[b] : 0.f;
  for (int j = 0; j < A; ++j) {
    float p = expf(z[j] - mx) / sum, lp = z[j] - lse;
    dlogits[(size_t)b * A + j] = gl * ((j == a ? 1.f : 0.f) - p) - ge * p * (lp + h);
  }
}

// ------------------------------------------------- fused PPO loss (rows O+Q)
struct PpoArgs {
  int B, A;
  const float* logits;
  const long long* actions;
  const float* old_lp;
  const float* adv;
  const float* values;
  const float* value_preds;
  const float* returns;
  const float* rl_mask;      // null -> plain mean (av_nav/ppo/ppo.py:99-109)
  const float* unct;         // (B, 2) or null
  const long long* unct_gt;  // (B) or null
  float clip, value_coef, ent_coef, unct_coef;
  int use_clipped_value;
  float* dlogits;
  float* dvalues;
  float* dunct;
  float* out;      // [value_loss, action_loss, entropy, unct_loss, total, values_mean, returns_mean, norm]
  float* partial;  // gridDim.x * 6 floats
  unsigned int* ticket;
};

__global__ void ppo_mask_sum_kernel(const float* __restrict__ rl_mask, int B, float* out_norm) {
  __shared__ float red[33];
  float s = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) s += rl_mask[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) *out_norm = s;
}

// savi/ppo/ppo.py:219-262 (and av_nav/ppo/ppo.py:93-131): every row's loss
// terms and the gradient of the total loss w.r.t. logits / value / uncertainty
// logits in one pass.  Tie handling of torch.min / torch.max backward (gradient
// split evenly) is reproduced.
__global__ void ppo_loss_kernel(PpoArgs p) {
  __shared__ float red[33];
  __shared__ bool is_last;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const float invB = 1.f / (float)p.B;
  const float norm = p.rl_mask ? p.out[7] : (float)p.B;
  float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // value, action, entropy, unct, v, ret
  if (b < p.B) {
    float z[kMaxA];
    const int A = p.A;
    for (int j = 0; j < A; ++j) z[j] = p.logits[(size_t)b * A + j];
    float mx, sum;
    float lse = row_lse(z, A, mx, sum);
    float h = 0.f;
    for (int j = 0; j < A; ++j) h -= (expf(z[j] - mx) / sum) * (z[j] - lse);
    const int a = (int)p.actions[b];
    const float alp = z[a] - lse;
    const float ratio = expf(alp - p.old_lp[b]);
    const float m = p.rl_mask ? p.rl_mask[b] : 1.f;
    const float am = p.adv[b] * m;
    const float rc = fminf(fmaxf(ratio, 1.f - p.clip), 1.f + p.clip);
    const float s1 = ratio * am, s2 = rc * am;
    acc[1] = -fminf(s1, s2);
    const float d1 = am * ratio;
    const float d2 = (ratio >= 1.f - p.clip && ratio <= 1.f + p.clip) ? am * ratio : 0.f;
    float dsur = (s1 < s2) ? d1 : ((s1 > s2) ? d2 : 0.5f * (d1 + d2));
    const float g_alp = -dsur / norm;
    const float g_ent = -p.ent_coef * invB;
    for (int j = 0; j < A; ++j) {
      float pj = expf(z[j] - mx) / sum, lp = z[j] - lse;
      p.dlogits[(size_t)b * A + j] = g_alp * ((j == a ? 1.f : 0.f) - pj) - g_ent * pj * (lp + h);
    }
    acc[2] = h;
    // value loss
    const float v = p.values[b], vo = p.value_preds[b], R = p.returns[b];
    float dv;
    if (p.use_clipped_value) {
      const float diff = v - vo;
      const float vc = vo + fminf(fmaxf(diff, -p.clip), p.clip);
      const float l1 = (v - R) * (v - R), l2 = (vc - R) * (vc - R);
      acc[0] = fmaxf(l1, l2);
      const float g1 = 2.f * (v - R);
      const float g2 = (diff >= -p.clip && diff <= p.clip) ? 2.f * (vc - R) : 0.f;
      dv = (l1 > l2) ? g1 : ((l1 < l2) ? g2 : 0.5f * (g1 + g2));
    } else {
      acc[0] = (R - v) * (R - v);
      dv = 2.f * (v - R);
    }
    p.dvalues[b] = p.value_coef * 0.5f * invB * dv;
    acc[4] = v;
    acc[5] = R;
    if (p.unct) {
      const float u0 = p.unct[2 * b], u1 = p.unct[2 * b + 1];
      const float um = fmaxf(u0, u1);
      const float e0 = expf(u0 - um), e1 = expf(u1 - um);
      const float ulse = um + logf(e0 + e1);
      const int g = (int)p.unct_gt[b];
      acc[3] = ulse - (g ? u1 : u0);
      const float q0 = e0 / (e0 + e1), q1 = e1 / (e0 + e1);
      p.dunct[2 * b] = p.unct_coef * invB * (q0 - (g == 0 ? 1.f : 0.f));
      p.dunct[2 * b + 1] = p.unct_coef * invB * (q1 - (g == 1 ? 1.f : 0.f));
    }
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    float s = block_sum(acc[k], red);
    if (threadIdx.x == 0) p.partial[(size_t)blockIdx.x * 6 + k] = s;
  }
  __threadfence();
  if (threadIdx.x == 0) {
    unsigned int t = atomicAdd(p.ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    __threadfence();
    float tot[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (unsigned int g = 0; g < gridDim.x; ++g)
      for (int k = 0; k < 6; ++k) tot[k] += __ldcg(&p.partial[(size_t)g * 6 + k]);
    float value_loss = 0.5f * tot[0] * invB;
    float action_loss = tot[1] / norm;
    float entropy = tot[2] * invB;
    float unct_loss = p.unct ? tot[3] * invB : 0.f;
    p.out[0] = value_loss;
    p.out[1] = action_loss;
    p.out[2] = entropy;
    p.out[3] = unct_loss;
    p.out[4] = value_loss * p.value_coef + action_loss - entropy * p.ent_coef + p.unct_coef * unct_loss;
    p.out[5] = tot[4] * invB;
    p.out[6] = tot[5] * invB;
    if (!p.rl_mask) p.out[7] = norm;
    *p.ticket = 0u;
  }
}

// --------------------------------------------- external memory insert (row G)
// rollout_storage.py:930-941 + :284-295 on a SINGLE-copy memory (total, N, dim):
// the reference's T+1 identical copies are redundant given the per-step mask
// snapshots.  One block per env.
__global__ void extmem_insert_kernel(float* memory, float* masks, const float* __restrict__ feats,
                                     const float* __restrict__ not_done, float* snapshot, int n, int total,
                                     int capacity, int dim, int idx) {
  __shared__ float red[33];
  const int e = blockIdx.x;
  for (int i = threadIdx.x; i < dim; i += blockDim.x)
    memory[((size_t)idx * n + e) * dim + i] = feats[(size_t)e * dim + i];
  float* mrow = masks + (size_t)e * total;
  float s = 0.f;
  for (int i = threadIdx.x; i < total; i += blockDim.x) s += mrow[i];
  s = block_sum(s, red);  // exact: entries are 0/1 and total < 2^24
  const bool overflow = (s == (float)capacity);
  int evict = idx - capacity;
  if (evict < 0) evict += total;  // python negative index
  const float nd = not_done[e];
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    float m = mrow[i];
    if (overflow && i == evict) m = 0.f;
    if (i == idx) m = 1.f;
    m *= nd;
    mrow[i] = m;
    if (snapshot) snapshot[(size_t)e * total + i] = m;
  }
}

// The same with the ring position read from device memory (a rollout step captured into a CUDA graph must not bake the
// position into its kernel arguments); counter_advance_kernel moves it afterwards.
__global__ void extmem_insert_dev_kernel(float* memory, float* masks, const float* __restrict__ feats,
                                         const float* __restrict__ not_done, float* snapshot, int n, int total,
                                         int capacity, int dim, const int* __restrict__ idx_dev) {
  __shared__ float red[33];
  const int idx = *idx_dev;
  const int e = blockIdx.x;
  for (int i = threadIdx.x; i < dim; i += blockDim.x)
    memory[((size_t)idx * n + e) * dim + i] = feats[(size_t)e * dim + i];
  float* mrow = masks + (size_t)e * total;
  float s = 0.f;
  for (int i = threadIdx.x; i < total; i += blockDim.x) s += mrow[i];
  s = block_sum(s, red);
  const bool overflow = (s == (float)capacity);
  int evict = idx - capacity;
  if (evict < 0) evict += total;
  const float nd = not_done[e];
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    float m = mrow[i];
    if (overflow && i == evict) m = 0.f;
    if (i == idx) m = 1.f;
    m *= nd;
    mrow[i] = m;
    if (snapshot) snapshot[(size_t)e * total + i] = m;
  }
}
__global__ void counter_advance_kernel(int* c, int mod) { *c = (*c + 1) % mod; }

// ------------------------------------------------ belief EMA update (row M)
// belief_predictor.py:139-230 batched: odom<->base transforms and EMA per env.
__device__ __forceinline__ void odom_to_base(float gx, float gy, const float* pose, float& bx, float& by) {
  float angle = -pose[2];
  float dx = gx - pose[0], dy = gy - pose[1];
  float dth = atan2f(dy, dx) - angle;
  float d = sqrtf(dx * dx + dy * dy);
  bx = d * cosf(dth);
  by = d * sinf(dth);
}
__device__ __forceinline__ void base_to_odom(float bx, float by, const float* pose, float& gx, float& gy) {
  float angle = -pose[2];
  float d = sqrtf(bx * bx + by * by);
  float th = atan2f(by, bx);
  gx = pose[0] + d * cosf(th + angle);
  gy = pose[1] + d * sinf(th + angle);
}

__global__ void spec_nonzero_kernel(const float* __restrict__ spec, int per_env, int* flag) {
  __shared__ float red[33];
  const float* s = spec + (size_t)blockIdx.x * per_env;
  float acc = 0.f;
  for (int i = threadIdx.x; i < per_env; i += blockDim.x) acc += s[i];
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) flag[blockIdx.x] = (acc != 0.f) ? 1 : 0;
}

__global__ void belief_update_kernel(int n, const int* __restrict__ nonzero, const float* __restrict__ pose,
                                     const unsigned char* __restrict__ dones, const float* __restrict__ pointgoal_pred,
                                     const float* __restrict__ label_pred, int label_stride, float w,
                                     int current_pred_only, float* last_pointgoal, int* has_pointgoal,
                                     float* last_label, int* has_label, float* location_belief,
                                     float* category_belief) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const bool done = dones && dones[e];
  const bool nz = nonzero[e] != 0;
  const float* ps = pose + (size_t)e * 4;
  if (pointgoal_pred) {
    bool has = has_pointgoal[e] != 0 && !done;
    float ax, ay;
    if (nz) {
      float bx = -pointgoal_pred[2 * e + 1], by = pointgoal_pred[2 * e];
      if (!has || current_pred_only) {
        ax = bx; ay = by;
      } else {
        float ox, oy;
        odom_to_base(last_pointgoal[2 * e], last_pointgoal[2 * e + 1], ps, ox, oy);
        ax = (1.f - w) * bx + w * ox;
        ay = (1.f - w) * by + w * oy;
      }
      float gx, gy;
      base_to_odom(ax, ay, ps, gx, gy);
      last_pointgoal[2 * e] = gx;
      last_pointgoal[2 * e + 1] = gy;
      has = true;
    } else {
      if (!has) { ax = 10.f; ay = 10.f; }
      else odom_to_base(last_pointgoal[2 * e], last_pointgoal[2 * e + 1], ps, ax, ay);
    }
    has_pointgoal[e] = has ? 1 : 0;
    location_belief[2 * e] = ax;
    location_belief[2 * e + 1] = ay;
  }
  if (label_pred) {
    bool has = has_label[e] != 0 && !done;
    for (int j = 0; j < 21; ++j) {
      float cur = label_pred[(size_t)e * label_stride + j];
      float out;
      if (nz) {
        out = (!has || current_pred_only) ? cur : (1.f - w) * cur + w * last_label[(size_t)e * 21 + j];
        last_label[(size_t)e * 21 + j] = out;
      } else {
        out = has ? last_label[(size_t)e * 21 + j] : (1.f / 21.f);
      }
      category_belief[(size_t)e * 21 + j] = out;
    }
    has_label[e] = (nz || has) ? 1 : 0;
  }
}

// ---------------------------------------- global-norm clip + Adam (ppo.py:62,297-300)
__global__ void sumsq_kernel(const float* __restrict__ g, long long n, float* partial, unsigned int* ticket,
                             float* out) {
  __shared__ float red[33];
  __shared__ bool is_last;
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = g[i];
    s += v * v;
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = s;
    __threadfence();
    unsigned int t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    __threadfence();
    double tot = 0.0;
    for (unsigned int b = 0; b < gridDim.x; ++b) tot += (double)__ldcg(&partial[b]);
    out[0] = (float)tot;
    *ticket = 0u;
  }
}

__global__ void adam_kernel(float* p, const float* __restrict__ g, float* m, float* v, long long n, float lr,
                            float beta1, float beta2, float eps, float bc1, float bc2_sqrt, float max_norm,
                            const float* normsq, float grad_scale) {
  float coef = 1.f;
  if (max_norm > 0.f) {
    float nrm = sqrtf(*normsq) * grad_scale;
    coef = fminf(max_norm / (nrm + 1e-6f), 1.f);
  }
  coef *= grad_scale;
  const float step_size = lr / bc1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float gi = g[i] * coef;
    float mi = m[i];
    mi = mi + (gi - mi) * (1.f - beta1);
    float vi = v[i] * beta2 + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - step_size * (mi / denom);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Row R: the imitation loss of PPO.update_dialog (ss_baselines/savi/ppo/ppo.py:134-142):
//   rows = nonzero(o_masks);  loss = CrossEntropyLoss(weight = w)(logits[rows], o_actions[rows])
//        = sum_i w[y_i] * (logsumexp(l_i) - l_i[y_i]) / sum_i w[y_i]
// One CTA, two passes over B = NUM_DIALOG_STEPS * N rows: the row selection never leaves the device (the reference's
// `nonzero` is a host synchronisation) and the gradient dlogits = mask * w[y] * (softmax - onehot) / sum w is written
// in the same launch.  out[0] = loss, out[1] = sum of weights, out[2] = number of selected rows.  No selected row or
// zero weight sum gives NaN like the reference (0/0).
__global__ void masked_weighted_ce_kernel(const float* __restrict__ logits, const float* __restrict__ targets,
                                          const long long* __restrict__ mask, const float* __restrict__ weight, int B,
                                          int A, float* dlogits, float* out) {
  __shared__ float red[33];
  float ls = 0.f, ws = 0.f, cnt = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    if (mask[b] == 0) continue;
    const float* l = logits + (size_t)b * A;
    int y = (int)targets[b];
    y = y < 0 ? 0 : (y >= A ? A - 1 : y);
    float mx = l[0];
    for (int a = 1; a < A; ++a) mx = fmaxf(mx, l[a]);
    float se = 0.f;
    for (int a = 0; a < A; ++a) se += expf(l[a] - mx);
    const float w = weight ? weight[y] : 1.f;
    ls += w * (logf(se) + mx - l[y]);
    ws += w;
    cnt += 1.f;
  }
  ls = block_sum(ls, red);
  ws = block_sum(ws, red);
  cnt = block_sum(cnt, red);
  if (threadIdx.x == 0) {
    out[0] = ls / ws;
    out[1] = ws;
    out[2] = cnt;
  }
  if (!dlogits) return;
  const float inv = 1.f / ws;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    float* d = dlogits + (size_t)b * A;
    if (mask[b] == 0) {
      for (int a = 0; a < A; ++a) d[a] = 0.f;
      continue;
    }
    const float* l = logits + (size_t)b * A;
    int y = (int)targets[b];
    y = y < 0 ? 0 : (y >= A ? A - 1 : y);
    float mx = l[0];
    for (int a = 1; a < A; ++a) mx = fmaxf(mx, l[a]);
    float se = 0.f;
    for (int a = 0; a < A; ++a) se += expf(l[a] - mx);
    const float w = (weight ? weight[y] : 1.f) * inv;
    for (int a = 0; a < A; ++a) d[a] = w * (expf(l[a] - mx) / se - (a == y ? 1.f : 0.f));
  }
}

}  // namespace

#ifndef AVL_HOST_EMUL
AVL_API int avl_gae(const float* rewards, float* value_preds, const float* masks, const float* next_value,
                    float* returns, int steps, int n_envs, int use_gae, float gamma, float tau, void* stream) {
  if (steps < 0 || n_envs < 0) return AVL_ERR_ARG;
  if (n_envs == 0) return AVL_OK;
  if (!rewards || !value_preds || !masks || !next_value || !returns) return AVL_ERR_ARG;
  // python evaluates gamma * tau in double before the tensor multiply (rollout_storage.py:404)
  float gt = (float)((double)gamma * (double)tau);
  gae_kernel<<<avl_div_up(n_envs, 128), 128, 0, (cudaStream_t)stream>>>(rewards, value_preds, masks, next_value,
                                                                       returns, steps, n_envs, use_gae, gamma, gt);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}

// double-precision scalars variant: gamma and tau arrive exactly as the python floats
AVL_API int avl_gae_f64(const float* rewards, float* value_preds, const float* masks, const float* next_value,
                        float* returns, int steps, int n_envs, int use_gae, double gamma, double tau, void* stream) {
  if (steps < 0 || n_envs < 0) return AVL_ERR_ARG;
  if (n_envs == 0) return AVL_OK;
  if (!rewards || !value_preds || !masks || !next_value || !returns) return AVL_ERR_ARG;
  gae_kernel<<<avl_div_up(n_envs, 128), 128, 0, (cudaStream_t)stream>>>(
      rewards, value_preds, masks, next_value, returns, steps, n_envs, use_gae, (float)gamma, (float)(gamma * tau));
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}

AVL_API int avl_advantages(const float* returns, const float* value_preds, float* adv, int count, int normalize,
                           float eps, void* stream) {
  if (count < 0) return AVL_ERR_ARG;
  if (count == 0) return AVL_OK;
  if (!returns || !value_preds || !adv) return AVL_ERR_ARG;
  if (normalize && count < 2) return AVL_ERR_ARG;
  advantages_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(returns, value_preds, adv, count, normalize, eps);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}

AVL_API int avl_categorical_act(const float* logits, const float* uniforms, int B, int A, long long* actions,
                                float* log_probs, float* probs, void* stream) {
  if (B < 0 || A < 1) return AVL_ERR_ARG;
  if (A > kMaxA) return AVL_ERR_UNSUPPORTED;
  if (B == 0) return AVL_OK;
  if (!logits || !actions || !log_probs) return AVL_ERR_ARG;
  categorical_act_kernel<<<avl_div_up(B, 128), 128, 0, (cudaStream_t)stream>>>(logits, uniforms, B, A, actions,
                                                                              log_probs, probs);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}

AVL_API int avl_categorical_eval(const float* logits, const long long* actions, int B, int A, float* log_probs,
                                 float* entropy, float* probs, void* stream) {
  if (B < 0 || A < 1) return AVL_ERR_ARG;
  if (A > kMaxA) return AVL_ERR_UNSUPPORTED;
  if (B == 0) return AVL_OK;
  if (!logits || !actions || !log_probs || !entropy) return AVL_ERR_ARG;
  categorical_eval_kernel<<<avl_div_up(B, 128), 128, 0, (cudaStream_t)stream>>>(logits, actions, B, A, log_probs,
                                                                               entropy, probs);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}

AVL_API int avl_categorical_eval_bwd(const float* logits, const long long* actions, const float* g_log_probs,
                                     const float* g_entropy, int B, int A, float* dlogits, void* stream) {
  if (B < 0 || A < 1) return AVL_ERR_ARG;
  if (A > kMaxA) return AVL_ERR_UNSUPPORTED;
  if (B == 0) return AVL_OK;
  if (!logits || !actions || !dlogits) return AVL_ERR_ARG;
  categorical_eval_bwd_kernel<<<avl_div_up(B, 128), 128, 0, (cudaStream_t)stream>>>(logits, actions, g_log_probs,
                                                                                   g_entropy, B, A, dlogits);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}

// workspace: at least avl_ppo_loss_workspace(B) bytes, zero-initialised once by the caller.
AVL_API long long avl_ppo_loss_workspace(int B) { return (long long)(avl_div_up(B, 256) * 6 + 4) * 4; }

AVL_API int avl_ppo_loss_fwd_bwd(int B, int A, const float* logits, const long long* actions, const float* old_lp,
                                 const float* adv, const float* values, const float* value_preds,
                                 const float* returns, const float* rl_mask, const float* unct,
                                 const long long* unct_gt, float clip, float value_coef, float ent_coef,
                                 float unct_coef, int use_clipped_value, float* dlogits, float* dvalues,
                                 float* dunct, float* out8, void* workspace, void* stream) {
  if (B < 1 || A < 1) return AVL_ERR_ARG;
  if (A > kMaxA) return AVL_ERR_UNSUPPORTED;
  if (!logits || !actions || !old_lp || !adv || !values || !value_preds || !returns || !dlogits || !dvalues ||
      !out8 || !workspace)
    return AVL_ERR_ARG;
  if ((unct != nullptr) != (unct_gt != nullptr) || (unct != nullptr) != (dunct != nullptr)) return AVL_ERR_ARG;
  PpoArgs p;
  p.B = B; p.A = A; p.logits = logits; p.actions = actions; p.old_lp = old_lp; p.adv = adv; p.values = values;
  p.value_preds = value_preds; p.returns = returns; p.rl_mask = rl_mask; p.unct = unct; p.unct_gt = unct_gt;
  p.clip = clip; p.value_coef = value_coef; p.ent_coef = ent_coef; p.unct_coef = unct_coef;
  p.use_clipped_value = use_clipped_value; p.dlogits = dlogits; p.dvalues = dvalues; p.dunct = dunct; p.out = out8;
  p.ticket = reinterpret_cast<unsigned int*>(workspace);
  p.partial = reinterpret_cast<float*>(workspace) + 4;
  cudaStream_t s = (cudaStream_t)stream;
  if (rl_mask) {
    ppo_mask_sum_kernel<<<1, 1024, 0, s>>>(rl_mask, B, out8 + 7);
    AVL_LAUNCH_CHECK();
  }
  ppo_loss_kernel<<<avl_div_up(B, 256), 256, 0, s>>>(p);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}

// logits (B, A); targets (B) float (o_actions storage dtype); mask (B) int64 (o_masks); weight (A) or NULL;
// dlogits (B, A) or NULL; out3 (3 floats, device).
AVL_API int avl_masked_weighted_ce(const float* logits, const float* targets, const long long* mask,
                                   const float* weight, int B, int A, float* dlogits, float* out3, void* stream) {
  if (B < 1 || A < 1) return AVL_ERR_ARG;
  if (!logits || !targets || !mask || !out3) return AVL_ERR_ARG;
  AVL_LAUNCH(masked_weighted_ce_kernel, 1, 1024, 0, (cudaStream_t)stream, logits, targets, mask, weight, B, A, dlogits,
             out3);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}

AVL_API int avl_extmem_insert(float* memory, float* masks, const float* feats, const float* not_done,
                              float* mask_snapshot, int n_envs, int total_size, int capacity, int dim, int idx,
                              void* stream) {
  if (n_envs < 0 || total_size < 1 || capacity < 0 || dim < 1 || idx < 0 || idx >= total_size) return AVL_ERR_ARG;
  if (n_envs == 0) return AVL_OK;
  if (!memory || !masks || !feats || !not_done) return AVL_ERR_ARG;
  extmem_insert_kernel<<<n_envs, 128, 0, (cudaStream_t)stream>>>(memory, masks, feats, not_done, mask_snapshot,
                                                                n_envs, total_size, capacity, dim, idx);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}

AVL_API int avl_extmem_insert_dev(float* memory, float* masks, const float* feats, const float* not_done,
                                  float* mask_snapshot, int n_envs, int total_size, int capacity, int dim, int* idx_dev,
                                  void* stream) {
  if (n_envs < 0 || total_size < 1 || capacity < 0 || dim < 1) return AVL_ERR_ARG;
  if (n_envs == 0) return AVL_OK;
  if (!memory || !masks || !feats || !not_done || !idx_dev) return AVL_ERR_ARG;
  extmem_insert_dev_kernel<<<n_envs, 128, 0, (cudaStream_t)stream>>>(memory, masks, feats, not_done, mask_snapshot,
                                                                    n_envs, total_size, capacity, dim, idx_dev);
  AVL_LAUNCH_CHECK();
  counter_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(idx_dev, total_size);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}

AVL_API int avl_belief_update(int n_envs, const float* spectrogram, int spec_elems_per_env, const float* pose,
                              const unsigned char* dones, const float* pointgoal_pred, const float* label_pred,
                              int label_stride, float weighting_factor, int current_pred_only,
                              float* last_pointgoal, int* has_pointgoal, float* last_label, int* has_label,
                              float* location_belief, float* category_belief, int* nonzero_scratch, void* stream) {
  if (n_envs < 0) return AVL_ERR_ARG;
  if (n_envs == 0) return AVL_OK;
  if (!spectrogram || !pose || !nonzero_scratch) return AVL_ERR_ARG;
  if (pointgoal_pred && (!last_pointgoal || !has_pointgoal || !location_belief)) return AVL_ERR_ARG;
  if (label_pred && (!last_label || !has_label || !category_belief || label_stride < 21)) return AVL_ERR_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  spec_nonzero_kernel<<<n_envs, 256, 0, s>>>(spectrogram, spec_elems_per_env, nonzero_scratch);
  AVL_LAUNCH_CHECK();
  belief_update_kernel<<<avl_div_up(n_envs, 128), 128, 0, s>>>(n_envs, nonzero_scratch, pose, dones, pointgoal_pred,
                                                              label_pred, label_stride, weighting_factor,
                                                              current_pred_only, last_pointgoal, has_pointgoal,
                                                              last_label, has_label, location_belief, category_belief);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}

// workspace: (1024 + 8) floats, zero-initialised once by the caller.  normsq_out[0] receives sum(g^2).
AVL_API int avl_grad_sumsq(const float* grad, long long n, float* normsq_out, void* workspace, void* stream) {
  if (n < 0 || !grad || !normsq_out || !workspace) return AVL_ERR_ARG;
  unsigned int* ticket = reinterpret_cast<unsigned int*>(workspace);
  float* partial = reinterpret_cast<float*>(workspace) + 8;
  int grid = avl_div_up(n > 0 ? n : 1, 256 * 8);
  if (grid > 1024) grid = 1024;
  sumsq_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(grad, n, partial, ticket, normsq_out);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}

// Fused clip_grad_norm_(max_norm) + torch.optim.Adam step over flat buffers.
// grad_scale multiplies the gradient first (1/world_size after a SUM all-reduce).
// normsq is the device scalar written by avl_grad_sumsq on the UNSCALED gradient.
AVL_API int avl_clip_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n,
                               float lr, float beta1, float beta2, float eps, int step, float max_norm,
                               const float* normsq, float grad_scale, void* stream) {
  if (n < 0 || step < 1) return AVL_ERR_ARG;
  if (n == 0) return AVL_OK;
  if (!param || !grad || !exp_avg || !exp_avg_sq) return AVL_ERR_ARG;
  if (max_norm > 0.f && !normsq) return AVL_ERR_ARG;
  double bc1 = 1.0 - pow((double)beta1, (double)step);
  double bc2 = 1.0 - pow((double)beta2, (double)step);
  int grid = avl_div_up(n, 256 * 4);
  int cap = avl_num_sms() * 8;
  if (grid > cap) grid = cap;
  adam_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                     (float)bc1, (float)sqrt(bc2), max_norm, normsq, grad_scale);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}
// ------------------------------------------------------------------------------------------------------------------
// Synthetic VectorEnv step (avlen_b200/synth_env.py): episode bookkeeping + toy kinematics
// of all environments in ONE launch.  Stands in for the reference's env workers (graph walk simulator.py:496-517,
// episode reset, _audio_index advance simulator.py:668, silent-source test simulator.py:646) behind the VectorEnv
// `step` call; as ~30 separate elementwise launches it sat on the critical path between two policy steps.
// r: (n, 3) uniforms; dones = r[:,0] < done_prob; FORWARD=1 moves 0.5 m along the heading, LEFT=2 / RIGHT=3 turn by
// 0.5236 rad; a finished episode resets pose / heading / step count; rewards = r[:,1] - 0.5.  Products and sums are
// rounded separately (no FMA contraction) so that the result equals the elementwise formulation bit for bit.
namespace {
__global__ void synth_env_step_kernel(int n, const long long* __restrict__ actions, const float* __restrict__ r,
                                      float done_prob, float* heading, float* pose_xy, float* episode_step,
                                      int* audio_index, const int* __restrict__ clip_secs,
                                      const float* __restrict__ silent_after, float* rewards, unsigned char* dones,
                                      float* masks, float* pose_obs, int* silent, float* category_belief,
                                      float* location_belief) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool done = r[3 * i] < done_prob;
  const long long a = actions[i];
  float h = heading[i];
  if (a == 2) h = __fadd_rn(h, 0.5236f);
  else if (a == 3) h = __fsub_rn(h, 0.5236f);
  const float half_fwd = (a == 1) ? 0.5f : 0.f;
  float px = __fadd_rn(pose_xy[2 * i], __fmul_rn(half_fwd, cosf(h)));
  float py = __fadd_rn(pose_xy[2 * i + 1], __fmul_rn(half_fwd, -sinf(h)));
  float es = __fadd_rn(episode_step[i], 1.f);
  const float nd = done ? 0.f : 1.f;
  es = __fmul_rn(es, nd);
  px = __fmul_rn(px, nd);
  py = __fmul_rn(py, nd);
  h = __fmul_rn(h, nd);
  heading[i] = h;
  pose_xy[2 * i] = px;
  pose_xy[2 * i + 1] = py;
  episode_step[i] = es;
  const int secs = clip_secs[i];
  audio_index[i] = secs > 0 ? (audio_index[i] + 1) % secs : 0;
  rewards[i] = __fsub_rn(r[3 * i + 1], 0.5f);
  dones[i] = done ? 1 : 0;
  masks[i] = nd;
  pose_obs[4 * i] = px;
  pose_obs[4 * i + 1] = py;
  pose_obs[4 * i + 2] = h;
  pose_obs[4 * i + 3] = es;
  silent[i] = es > silent_after[i] ? 1 : 0;
  if (category_belief)
    for (int j = 0; j < 21; ++j) category_belief[21 * i + j] = 0.f;
  if (location_belief) location_belief[2 * i] = location_belief[2 * i + 1] = 0.f;
}
}  // namespace

AVL_API int avl_synth_env_step(int n_envs, const long long* actions, const float* uniforms, float done_prob,
                               float* heading, float* pose_xy, float* episode_step, int* audio_index,
                               const int* clip_secs, const float* silent_after, float* rewards, unsigned char* dones,
                               float* masks, float* pose_obs, int* silent, float* category_belief,
                               float* location_belief, void* stream) {
  if (n_envs < 0) return AVL_ERR_ARG;
  if (n_envs == 0) return AVL_OK;
  if (!actions || !uniforms || !heading || !pose_xy || !episode_step || !audio_index || !clip_secs || !silent_after ||
      !rewards || !dones || !masks || !pose_obs || !silent)
    return AVL_ERR_ARG;
  synth_env_step_kernel<<<avl_div_up(n_envs, 128), 128, 0, (cudaStream_t)stream>>>(
      n_envs, actions, uniforms, done_prob, heading, pose_xy, episode_step, audio_index, clip_secs, silent_after,
      rewards, dones, masks, pose_obs, silent, category_belief, location_belief);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}
#endif  // AVL_HOST_EMUL
