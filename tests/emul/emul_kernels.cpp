// Host-emulated execution of the SIMT kernels (see cuda_emul.h).  Built by
// tests/emul/build.py into tests/emul/_emul.so and called through ctypes from
// the CPU test-suite.  TEST INFRASTRUCTURE ONLY.
#define AVL_HOST_EMUL 1
#include "cuda_emul.h"
AVL_EMUL_DEFINE_GLOBALS

extern "C" int avl_set_cuda_error(int e) { return e; }
int avl_num_sms() { return 8; }
extern "C" void avl_count_launch() {}

#include "../../avlen_b200/csrc/audio.cu"
#include "../../avlen_b200/csrc/rl.cu"
#include "../../avlen_b200/csrc/smt.cu"
#include "../../avlen_b200/csrc/conv.cu"
#include "../../avlen_b200/csrc/nn_bwd.cu"
#include "../../avlen_b200/csrc/clip_text.cu"

#define EMUL_API extern "C" __attribute__((visibility("default")))

EMUL_API int emul_audio_render(int n_envs, int sr, const float* sounds, const long long* clip_off, const int* index,
                               const float* rirs, const long long* rir_off, const int* rir_len, const int* silent,
                               const long long* d_clip_off, const long long* d_rir_off, const int* d_rir_len,
                               float* audiogoal, float* spectrogram, int grid) {
  std::vector<cf> tw(kM + 1);
  emul::launch(dim3((kM + 1 + 255) / 256), dim3(256), [&] { twiddle_init_kernel(tw.data()); });
  const int split = grid < 0 ? 1 : 0;  // negative grid: one CTA per (env, ear) (RenderArgs::split)
  if (grid < 0) grid = -grid;
  std::vector<cf> scratch((size_t)3 * (kM + 1) * grid);
  int status = 0;
  RenderArgs a;
  a.split = split;
  a.n_envs = n_envs; a.sr = sr; a.sounds = sounds; a.clip_off = clip_off; a.index = index; a.rirs = rirs;
  a.rir_off = rir_off; a.rir_len = rir_len; a.silent = silent; a.d_clip_off = d_clip_off; a.d_rir_off = d_rir_off;
  a.d_rir_len = d_rir_len; a.audiogoal = audiogoal; a.spectrogram = spectrogram; a.tw = tw.data();
  a.scratch = scratch.data(); a.status = &status;
  emul::launch(dim3(grid), dim3(kThreads), [&] { audio_render_kernel(a); });
  return status;
}

// Spectral asset banks: rows made by audio_spectra_kernel from the SAME time-domain descriptors (RIR i -> row i, with the
// distractor RIRs behind them; source row i = (clip_off[i], index[i]), distractor source rows behind them at second 0),
// then audio_render_spectral_kernel.
EMUL_API int emul_audio_render_spectral(int n_envs, int sr, const float* sounds, const long long* clip_off, const int* index,
                                        const float* rirs, const long long* rir_off, const int* rir_len, const int* silent,
                                        const long long* d_clip_off, const long long* d_rir_off, const int* d_rir_len,
                                        float* audiogoal, float* spectrogram, int grid) {
  std::vector<cf> tw(kM + 1);
  emul::launch(dim3((kM + 1 + 255) / 256), dim3(256), [&] { twiddle_init_kernel(tw.data()); });
  const bool dis = d_clip_off != nullptr;
  const int nr = dis ? 2 * n_envs : n_envs;
  std::vector<long long> roff(nr), soff(nr), rir_row(n_envs), d_rir_row(n_envs), src_row0(n_envs), d_src_row0(n_envs);
  std::vector<int> rlen(nr), sidx(nr);
  for (int i = 0; i < n_envs; ++i) {
    roff[i] = rir_off[i]; rlen[i] = rir_len[i]; soff[i] = clip_off[i]; sidx[i] = index[i];
    rir_row[i] = rir_len[i] > 0 ? i : -1;
    src_row0[i] = i;
    if (dis) {
      roff[n_envs + i] = d_rir_off[i]; rlen[n_envs + i] = d_rir_len[i]; soff[n_envs + i] = d_clip_off[i]; sidx[n_envs + i] = 0;
      d_rir_row[i] = d_rir_len[i] > 0 ? n_envs + i : -1;
      d_src_row0[i] = n_envs + i;
    }
  }
  std::vector<cf> rspec((size_t)nr * 2 * (kM + 1)), sspec((size_t)nr * (kM + 1));
  int status = 0;
  SpectraArgs sa;
  sa.n = nr; sa.sr = sr; sa.kind = 0; sa.bank = rirs; sa.off = roff.data(); sa.len_or_index = rlen.data(); sa.out = rspec.data();
  sa.tw = tw.data(); sa.status = &status;
  emul::launch(dim3(grid), dim3(kThreads), [&] { audio_spectra_kernel(sa); });
  sa.kind = 1; sa.bank = sounds; sa.off = soff.data(); sa.len_or_index = sidx.data(); sa.out = sspec.data();
  emul::launch(dim3(grid), dim3(kThreads), [&] { audio_spectra_kernel(sa); });
  SpectralArgs a;
  a.n_envs = n_envs; a.sr = sr; a.src_spec = sspec.data(); a.src_row0 = src_row0.data(); a.index = nullptr;
  a.rir_spec = rspec.data(); a.rir_row = rir_row.data(); a.silent = silent;
  a.d_src_row0 = dis ? d_src_row0.data() : nullptr; a.d_rir_row = dis ? d_rir_row.data() : nullptr;
  a.audiogoal = audiogoal; a.spectrogram = spectrogram; a.tw = tw.data();
  emul::launch(dim3(grid), dim3(kThreads), [&] { audio_render_spectral_kernel(a); });
  return status;
}

EMUL_API int emul_spectrogram(int n, int sr, const float* audio, float* spectrogram, int grid) {
  std::vector<cf> tw(kM + 1);
  emul::launch(dim3((kM + 1 + 255) / 256), dim3(256), [&] { twiddle_init_kernel(tw.data()); });
  emul::launch(dim3(grid), dim3(kThreads), [&] { spectrogram_kernel(audio, n, sr, spectrogram, tw.data()); });
  return 0;
}

EMUL_API int emul_gae(const float* rewards, float* value_preds, const float* masks, const float* next_value,
                      float* returns, int steps, int n, int use_gae, double gamma, double tau) {
  emul::launch(dim3((n + 127) / 128), dim3(128), [&] {
    gae_kernel(rewards, value_preds, masks, next_value, returns, steps, n, use_gae, (float)gamma, (float)(gamma * tau));
  });
  return 0;
}

EMUL_API int emul_advantages(const float* returns, const float* value_preds, float* adv, int count, int normalize,
                             float eps) {
  emul::launch(dim3(1), dim3(1024), [&] { advantages_kernel(returns, value_preds, adv, count, normalize, eps); });
  return 0;
}

EMUL_API int emul_categorical_act(const float* logits, const float* uniforms, int B, int A, long long* actions,
                                  float* log_probs, float* probs) {
  emul::launch(dim3((B + 127) / 128), dim3(128),
               [&] { categorical_act_kernel(logits, uniforms, B, A, actions, log_probs, probs); });
  return 0;
}

EMUL_API int emul_categorical_eval(const float* logits, const long long* actions, int B, int A, float* log_probs,
                                   float* entropy, float* probs, const float* g_lp, const float* g_ent,
                                   float* dlogits) {
  emul::launch(dim3((B + 127) / 128), dim3(128),
               [&] { categorical_eval_kernel(logits, actions, B, A, log_probs, entropy, probs); });
  if (dlogits)
    emul::launch(dim3((B + 127) / 128), dim3(128),
                 [&] { categorical_eval_bwd_kernel(logits, actions, g_lp, g_ent, B, A, dlogits); });
  return 0;
}

EMUL_API int emul_ppo_loss(int B, int A, const float* logits, const long long* actions, const float* old_lp,
                           const float* adv, const float* values, const float* value_preds, const float* returns,
                           const float* rl_mask, const float* unct, const long long* unct_gt, float clip,
                           float value_coef, float ent_coef, float unct_coef, int use_clipped_value, float* dlogits,
                           float* dvalues, float* dunct, float* out8) {
  int grid = (B + 255) / 256;
  std::vector<float> ws(grid * 6 + 4, 0.f);
  PpoArgs p;
  p.B = B; p.A = A; p.logits = logits; p.actions = actions; p.old_lp = old_lp; p.adv = adv; p.values = values;
  p.value_preds = value_preds; p.returns = returns; p.rl_mask = rl_mask; p.unct = unct; p.unct_gt = unct_gt;
  p.clip = clip; p.value_coef = value_coef; p.ent_coef = ent_coef; p.unct_coef = unct_coef;
  p.use_clipped_value = use_clipped_value; p.dlogits = dlogits; p.dvalues = dvalues; p.dunct = dunct; p.out = out8;
  p.ticket = reinterpret_cast<unsigned int*>(ws.data());
  p.partial = ws.data() + 4;
  if (rl_mask) emul::launch(dim3(1), dim3(1024), [&] { ppo_mask_sum_kernel(rl_mask, B, out8 + 7); });
  emul::launch(dim3(grid), dim3(256), [&] { ppo_loss_kernel(p); });
  return 0;
}

EMUL_API int emul_extmem_insert(float* memory, float* masks, const float* feats, const float* not_done,
                                float* snapshot, int n, int total, int capacity, int dim, int idx) {
  emul::launch(dim3(n), dim3(128),
               [&] { extmem_insert_kernel(memory, masks, feats, not_done, snapshot, n, total, capacity, dim, idx); });
  return 0;
}

EMUL_API int emul_belief_update(int n, const float* spectrogram, int per_env, const float* pose,
                                const unsigned char* dones, const float* pointgoal_pred, const float* label_pred,
                                int label_stride, float w, int current_pred_only, float* last_pointgoal,
                                int* has_pointgoal, float* last_label, int* has_label, float* location_belief,
                                float* category_belief) {
  std::vector<int> nz(n);
  emul::launch(dim3(n), dim3(256), [&] { spec_nonzero_kernel(spectrogram, per_env, nz.data()); });
  emul::launch(dim3((n + 127) / 128), dim3(128), [&] {
    belief_update_kernel(n, nz.data(), pose, dones, pointgoal_pred, label_pred, label_stride, w, current_pred_only,
                         last_pointgoal, has_pointgoal, last_label, has_label, location_belief, category_belief);
  });
  return 0;
}

EMUL_API int emul_clip_adam(float* param, const float* grad, float* m, float* v, long long n, float lr, float beta1,
                            float beta2, float eps, int step, float max_norm, float grad_scale, float* normsq_out) {
  std::vector<float> ws(1024 + 8, 0.f);
  int grid = (int)((n + 2047) / 2048);
  if (grid > 1024) grid = 1024;
  if (grid < 1) grid = 1;
  emul::launch(dim3(grid), dim3(256), [&] {
    sumsq_kernel(grad, n, ws.data() + 8, reinterpret_cast<unsigned int*>(ws.data()), normsq_out);
  });
  double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  emul::launch(dim3(4), dim3(256), [&] {
    adam_kernel(param, grad, m, v, n, lr, beta1, beta2, eps, (float)bc1, (float)sqrt(bc2), max_norm, normsq_out,
                grad_scale);
  });
  return 0;
}

EMUL_API int emul_masked_weighted_ce(const float* logits, const float* targets, const long long* mask,
                                     const float* weight, int B, int A, float* dlogits, float* out3) {
  emul::launch(dim3(1), dim3(1024),
               [&] { masked_weighted_ce_kernel(logits, targets, mask, weight, B, A, dlogits, out3); });
  return 0;
}
