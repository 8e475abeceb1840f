/* avlen_b200 — C ABI of the B200 (sm_100a) hot path of AVLEN / SAVi (libavlen_b200.so).
 *
 * Conventions: plain C, every pointer is a DEVICE pointer unless stated otherwise, row-major, float = fp32,
 * `stream` is a cudaStream_t passed as void*, every call only enqueues work (no hidden synchronisation unless the
 * comment says so), the library never allocates except inside avl_audio_create.  Return value: 0 = ok,
 * -1 = invalid argument, -2 = unsupported size, -3 = CUDA runtime error (see avl_last_cuda_error*).
 *
 * The reference (merlresearch/avlen) has no FFI on this path: its boundary is the Python class API in
 * ss_baselines/ and soundspaces/.  Each entry point below cites the reference code it replaces (file:line under
 * the reference root); avlen_b200/ mirrors the reference's Python classes and calls these through ctypes
 * (INTEGRATION.md shows the binding a reference maintainer would add).
 */
#ifndef AVLEN_B200_H
#define AVLEN_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------------------------------------- library */
int avl_version(void);
int avl_last_cuda_error(void);                 /* cudaError_t of the last failed runtime call (0 = none)  */
const char* avl_last_cuda_error_string(void);
int avl_device_sm_count(void);
long long avl_launch_count(void);
long long avl_launch_count_add(long long n); /* kernels of this library replayed through a caller-owned CUDA graph */              /* kernels launched by this library so far in this process */
int avl_set_tensor_cores(int level);           /* tcgen05 TF32: 0 off, 1 (default) encoder convs/FCs, 2 also SMT dense; returns old */
int avl_get_tensor_cores(void);

/* ------------------------------------------------------------------------------- rows A + B: audio sensors
 * soundspaces/simulator.py:644-699  SoundSpacesSim._compute_audiogoal   (causal FIR src * rir, 3 branches,
 *                                   silent frame, empty RIR file, distractor :682-697)
 * soundspaces/tasks/nav.py:87-101   SpectrogramSensor.compute_spectrogram (|STFT 512/160/400| -> 4x4 zero-padded
 *                                   block mean -> log1p, channels last)                                        */
int avl_audio_create(int sampling_rate, void** handle_out /* host */);
int avl_audio_destroy(void* handle);
int avl_audio_status(void* handle, int* status_out /* host; synchronises; 1 = an RIR was truncated */);
/* sounds: flat bank of mono clips; clip_off[i]: start of env i's clip; index[i]: _audio_index (second rendered);
 * rirs: bank of interleaved (L,2) RIRs; rir_off[i] in frames; rir_len[i] (0 = empty file -> zeros);
 * silent[i] != 0 -> exact zeros; d_*: optional distractor source/RIR (all three or none);
 * audiogoal_out (N,2,sr) may be NULL; spectrogram_out (N,65,ceil((1+sr/160)/4),2): (65,26,2) at 16 kHz, (65,69,2) at
 * 44.1 kHz (nav.py:78).  sr <= 16769: one 32768-point circular convolution per ear, RIRs up to 32768 - sr + 1 samples;
 * above (Replica): partitioned overlap-save over 16384-sample blocks, RIRs up to 49152 samples.                */
int avl_audio_render_spectrogram(void* handle, int n_envs, const float* sounds, const long long* clip_off,
                                 const int* index, const float* rirs, const long long* rir_off, const int* rir_len,
                                 const int* silent, const long long* d_clip_off, const long long* d_rir_off,
                                 const int* d_rir_len, float* audiogoal_out, float* spectrogram_out, void* stream);
int avl_audio_spectrogram(void* handle, int n, const float* audio /* (N,2,sr) */, float* spectrogram_out,
                          void* stream);
/* 1 (default): batches of at most half the SM count render each ear in its own CTA (rollout latency); 0: one CTA per
 * env always.  Returns the old setting. */
int avl_set_audio_channel_split(int on);
/* Spectral asset banks (the resident form of simulator.py:650-659's RIR files and :661-681's source seconds): the
 * forward transforms of a rendering depend on the assets only, so they are made once and kept in HBM; a rendering is then
 * a spectral product and ONE inverse transform per (env, ear).  sr <= 16769 only (AVL_ERR_UNSUPPORTED above).
 * avl_audio_spectrum_bins(): complex bins per spectrum row (16385).
 * rir_spectra: row i of spectra_out (n, 2, bins, 2) = both ears of RIR i (rir_off in frames, rir_len 0 -> zeros).
 * source_spectra: row i of spectra_out (n, bins, 2) = second index[i] of the clip at clip_off[i] with its history.
 * render_spectral: same outputs as avl_audio_render_spectrogram; env i's source row is src_row0[i] + index[i] (index
 * may be NULL), rir_row[i] < 0 = empty RIR file; the distractor pair (both or neither) renders second 0 of its clip. */
int avl_audio_spectrum_bins(void);
int avl_audio_rir_spectra(void* handle, int n, const float* rirs, const long long* rir_off, const int* rir_len,
                          float* spectra_out, void* stream);
int avl_audio_source_spectra(void* handle, int n, const float* sounds, const long long* clip_off, const int* index,
                             float* spectra_out, void* stream);
int avl_audio_render_spectral(void* handle, int n_envs, const float* src_spectra, const long long* src_row0,
                              const int* index, const float* rir_spectra, const long long* rir_row, const int* silent,
                              const long long* d_src_row0, const long long* d_rir_row, float* audiogoal_out,
                              float* spectrogram_out, void* stream);

/* ------------------------------------------------------------------------- rows N, O: returns / advantages
 * ss_baselines/savi/models/rollout_storage.py:394-412, ss_baselines/common/rollout_storage.py:114-132 (bit-exact)
 * ss_baselines/savi/ppo/ppo.py:90-95                                                                          */
int avl_gae(const float* rewards, float* value_preds, const float* masks, const float* next_value, float* returns,
            int steps, int n_envs, int use_gae, float gamma, float tau, void* stream);
int avl_gae_f64(const float* rewards, float* value_preds, const float* masks, const float* next_value,
                float* returns, int steps, int n_envs, int use_gae, double gamma, double tau, void* stream);
int avl_advantages(const float* returns, const float* value_preds, float* adv, int count, int normalize, float eps,
                   void* stream);

/* --------------------------------------------------------------------------------- row I: categorical heads
 * ss_baselines/common/utils.py:44-72 (CustomFixedCategorical sample/mode/log_probs/entropy, probs)
 * uniforms == NULL -> mode() (first arg-max); else inverse-CDF sampling on the supplied uniforms.              */
int avl_categorical_act(const float* logits, const float* uniforms, int B, int A, long long* actions,
                        float* log_probs, float* probs, void* stream);
int avl_categorical_eval(const float* logits, const long long* actions, int B, int A, float* log_probs,
                         float* entropy, float* probs, void* stream);
int avl_categorical_eval_bwd(const float* logits, const long long* actions, const float* g_log_probs,
                             const float* g_entropy, int B, int A, float* dlogits, void* stream);

/* ------------------------------------------------------------------------------- rows O + Q: fused PPO loss
 * ss_baselines/savi/ppo/ppo.py:219-262 (rl_mask / uncertainty CE variant), ss_baselines/av_nav/ppo/ppo.py:93-131.
 * out8 = [value_loss, action_loss, entropy, unct_loss, total, mean(values), mean(returns), normaliser];
 * dlogits/dvalues/dunct = gradient of the total loss.  workspace: avl_ppo_loss_workspace(B) bytes, zeroed once. */
long long avl_ppo_loss_workspace(int B);
int avl_ppo_loss_fwd_bwd(int B, int A, const float* logits, const long long* actions, const float* old_lp,
                         const float* adv, const float* values, const float* value_preds, const float* returns,
                         const float* rl_mask, const float* unct, const long long* unct_gt, float clip,
                         float value_coef, float ent_coef, float unct_coef, int use_clipped_value, float* dlogits,
                         float* dvalues, float* dunct, float* out8, void* workspace, void* stream);

/* ----------------------------------------------------------------------------- row G: external memory insert
 * ss_baselines/savi/models/rollout_storage.py:930-941 (+ mask snapshot :284-286) on the single-copy ring buffer
 * memory (total, N, dim), masks (N, total); bit-exact.                                                         */
int avl_extmem_insert(float* memory, float* masks, const float* feats, const float* not_done, float* mask_snapshot,
                      int n_envs, int total_size, int capacity, int dim, int idx, void* stream);
/* The same with the ring position in DEVICE memory (read by the kernel, advanced modulo total_size afterwards): a rollout
 * step captured into a CUDA graph replays with the right slot.                                                          */
int avl_extmem_insert_dev(float* memory, float* masks, const float* feats, const float* not_done, float* mask_snapshot,
                          int n_envs, int total_size, int capacity, int dim, int* idx_dev, void* stream);

/* --------------------------------------------------------------------- row M (scalar part): belief update
 * ss_baselines/savi/models/belief_predictor.py:139-230 (EMA + odom<->base transforms), batched.                */
int avl_belief_update(int n_envs, const float* spectrogram, int spec_elems_per_env, const float* pose,
                      const unsigned char* dones, const float* pointgoal_pred, const float* label_pred,
                      int label_stride, float weighting_factor, int current_pred_only, float* last_pointgoal,
                      int* has_pointgoal, float* last_label, int* has_label, float* location_belief,
                      float* category_belief, int* nonzero_scratch, void* stream);

/* ------------------------------------------------------------------- compact observation storage
 * x / 255 + exact 2x2 area mean (smt_cnn.py:83-95) reading the storage dtype directly (0 fp32, 1 fp16 depth, 2 uint8 rgb);
 * sample_index (optional, int64 [N]): output row n reads source sample sample_index[n] of the time-major storage — the
 * minibatch copies of rollout_storage.py:716-760 are never made.  y: (N, H/2, W/2, C_out) fp32, channels >= C zero.     */
int avl_resize_half_typed(const void* x, int dtype, const long long* sample_index, float* y, int N, int H, int W, int C,
                          int C_out, float scale, void* stream);
/* RolloutStorage.insert (savi/models/rollout_storage.py:214-295) as ONE launch: up to 48 (dst, src, n, kind) segments
 * (host arrays of device pointers); kind 0 byte copy (n bytes), 1 fp32 -> uint8, 2 fp32 -> fp16, 3 int64 -> fp32,
 * 4 fp32 -> int64 (n elements).                                                                                        */
int avl_multi_copy(int count, void* const* dst, const void* const* src, const long long* n, const unsigned char* kind,
                   void* stream);
/* ------------------------------------------------------------------- row T: batch_obs, host side
 * ss_baselines/common/utils.py:129-156 stacks the per-env observation arrays with np.stack / torch.tensor on the host.
 * avl_host_gather copies n pieces of HOST memory (dst[i][0..bytes[i]) = src[i][0..bytes[i])) with a small persistent
 * pool of native threads — the per-env frames land in one pinned staging buffer per sensor; no CUDA call is made.
 * avl_set_host_gather_threads: copy threads including the caller (0 = automatic); returns the old setting.             */
int avl_host_gather(const void* const* src, void* const* dst, const long long* bytes, int n);
int avl_set_host_gather_threads(int threads);
int avl_set_host_gather_streaming(int on); /* 1 (default): non-temporal stores into the staging buffer (the copy engine reads it next; dirty cache lines made the H2D copy 5x slower); 0: memcpy; returns old */
/* ------------------------------------------------------------------- RIR bank + spectrogram cache
 * Replaces the per-miss wav read (soundspaces/simulator.py:650-659) and the per-simulator dict caches keyed by
 * (source, receiver, azimuth) (:711-734; cleared on scene / sound change :393-395; `_audio_index` advances on a miss only,
 * :668).  tab_off / tab_len: dense (4, V, V) [azimuth][receiver][source] table into the packed RIR bank; valid: (n, V*V*4)
 * bytes; cache: (n, V*V*4, E) fp32.  lookup: RIR descriptors + hit flags, hits are marked silent for the render;
 * commit: hits read their cached spectrogram, misses store theirs and advance the clip position.                        */
int avl_spec_cache_lookup(int n, int V, const int* src, const int* recv, const int* az, const long long* tab_off,
                          const int* tab_len, const unsigned char* clear, unsigned char* valid, const int* silent_in,
                          long long* rir_off, int* rir_len, int* silent_out, unsigned char* hit, void* stream);
int avl_spec_cache_commit(int n, int V, int E, const int* src, const int* recv, const int* az, const unsigned char* hit,
                          const int* silent_in, float* cache, unsigned char* valid, float* spec, int* index,
                          const int* clip_secs, void* stream);
/* ------------------------------------------------------------------- graph-walk environment step
 * One launch for all envs: graph walk (soundspaces/simulator.py:496-517), first oracle action of the shortest path
 * (:758-787), reward (ss_baselines/common/environments.py:98-135), PoseSensor (soundspaces/tasks/nav.py:745-775), episode
 * end + VectorEnv auto-reset from a per-env episode table.  Scene tables: nbr (V,4) int32, hops (V,V) int16, next_dir
 * (V,V) int8 [target][node], points (V,2).  iargs: V, E, with_time_penalty, with_distance_reward, with_query_constraint,
 * consecutive_constraint, soft_query_reward, num_total_query, max_steps; fargs: grid_size, slack_reward, distance_scale,
 * success_reward, query_reward.  state: int32 [7][n] node, rot, source, ep_step, ep_cursor, start_node, start_rot.        */
int avl_graph_env_step(int n, const int* iargs, const float* fargs, const int* nbr, const short* hops,
                       const signed char* next_dir, const float* points, const int* ep_start, const int* ep_rot,
                       const int* ep_source, int* state, float* prev_dist, const long long* actions,
                       const unsigned char* is_queried, const long long* query_num, const float* cons_reward, float* rewards,
                       unsigned char* dones, float* masks, float* pose, long long* oracle, float* target_distance,
                       unsigned char* new_episode, int* azimuth, void* stream);
/* ------------------------------------------------------------------- AVLEN interactive step: query bookkeeping
 * Replaces the per-env Python loops of PPOTrainer._collect_rollout_step (ss_baselines/savi/ppo/ppo_trainer.py:394-416,
 * :449-460, :487-588, :639-694, :769-787).  state: int32 [5][n] = queried, dialog step, episode step, last query step,
 * query count; dialog_store: (n, L) int64 tokens of each env's current dialog.  Bit-exact vs the reference trace.      */
int avl_query_pre(int n, const unsigned char* new_episode, int* state, const float* pe, int pe_rows, int emb,
                  float* query_state, float* last_query_info, void* stream);
int avl_query_after_option(int n, const long long* actions_option, const float* target_distance,
                           const long long* pending_dialog, int L, int num_dialog_steps, float consecutive_reward,
                           int query_within_radius, int* state, long long* dialog_store, unsigned char* is_queried,
                           long long* query_num, float* cons_reward, long long* rl_mask, long long* cur_dialog,
                           float* agent_step, void* stream);
int avl_option_arbitrate(int n, const long long* actions_goal, const long long* actions_vln, const float* probs_goal, int A,
                         const long long* oracle, int oracle_when_queried, int allow_stop, int num_dialog_steps, int* state,
                         long long* dialog_store, int L, long long* actions, long long* o_mask, long long* ucnt_gt,
                         float* masks_vln, void* stream);
/* ------------------------------------------------- synthetic VectorEnv step (bench / tests)
 * Stands in for the reference's env workers behind VectorEnv.step (graph walk soundspaces/simulator.py:496-517,
 * _audio_index advance :668, silent-source test :646): episode bookkeeping + toy kinematics of all envs, one launch.
 * uniforms (n, 3); state arrays are updated in place; category_belief (n, 21) / location_belief (n, 2), when given,
 * are zero-filled (the belief predictor fills them afterwards).                                                    */
int avl_synth_env_step(int n_envs, const long long* actions, const float* uniforms, float done_prob, float* heading,
                       float* pose_xy, float* episode_step, int* audio_index, const int* clip_secs,
                       const float* silent_after, float* rewards, unsigned char* dones, float* masks, float* pose_obs,
                       int* silent, float* category_belief, float* location_belief, void* stream);

/* ----------------------------------------------------------------------- clip_grad_norm_ + Adam (flat buffers)
 * ss_baselines/savi/ppo/ppo.py:62, :297-300 (torch.optim.Adam + nn.utils.clip_grad_norm_)                      */
int avl_grad_sumsq(const float* grad, long long n, float* normsq_out, void* workspace /* 1032 floats, zeroed */,
                   void* stream);
int avl_clip_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                       float beta1, float beta2, float eps, int step, float max_norm, const float* normsq,
                       float grad_scale, void* stream);

/* --------------------------------------------------------------------- rows C, D, E, M: encoder building blocks
 * NHWC activations, weights in the reference's nn.Conv2d / nn.Linear layout.
 * ss_baselines/savi/models/audio_cnn.py:62-94,136-151; av_nav/models/visual_cnn.py:104-154;
 * ss_baselines/savi/models/smt_resnet.py:14-164 (GroupNorm(16) + residual + ReLU); smt_cnn.py:78-95 (/255, area resize);
 * belief_predictor.py:58-82 (torchvision resnet18: folded BatchNorm scale/bias, max-pool, global average pool).   */
int avl_conv2d_fwd(const float* x, int N, int H, int W, int C, const float* w_oihw, int Cout, int KH, int KW,
                   int stride, int pad, const float* scale, const float* bias, const float* residual, long long ldr,
                   int relu, float* y, long long ldy, void* stream);
int avl_tc_conv2d_fwd(const float* x, int N, int H, int W, int C, const float* w_ohwi, int Cout, int KH, int KW,
                      int stride, int pad, const float* scale, const float* bias, const float* residual,
                      long long ldr, int relu, float* y, long long ldy, void* stream);
int avl_groupnorm_fwd(const float* x, const float* gamma, const float* beta, const float* residual, float* y, int N,
                      int HW, int C, int groups, float eps, int relu, void* stream);
int avl_groupnorm_fwd_split(const float* x, const float* gamma, const float* beta, const float* residual, float* y,
                            int N, int HW, int C, int groups, float eps, int relu, double* stats_scratch,
                            void* stream);
/* one pass over HBM: one thread-block cluster per sample, per-group statistics exchanged through distributed shared
 * memory; AVL_ERR_UNSUPPORTED (-2, nothing launched) for shapes it does not cover                                   */
int avl_groupnorm_fwd_cluster(const float* x, const float* gamma, const float* beta, const float* residual, float* y,
                              int N, int HW, int C, int groups, float eps, int relu, void* stream);
int avl_resize_half(const float* x, float* y, int N, int H, int W, int C, int C_out, float scale, void* stream);
int avl_pad_channels(const float* x, float* y, long long rows, int C, int C_out, void* stream);
int avl_concat_rgbd(const float* rgb, const float* depth, float* y, long long pixels, int c_rgb, int c_depth,
                    float rgb_scale, void* stream);
int avl_append_planes(const float* x, const float* extra, float* y, int N, int HW, int C, int E, void* stream);
int avl_maxpool3x3s2(const float* x, float* y, int N, int H, int W, int C, void* stream);
int avl_avgpool_global(const float* x, float* y, int N, int HW, int C, void* stream);
int avl_onehot_linear(const long long* actions, const float* W, const float* bias, float* y, long long ldy, int B,
                      int out_dim, int n_actions, void* stream);
int avl_copy_cols(const float* src, long long lds, float* dst, long long ldd, int rows, int cols, void* stream);

/* ------------------------------------------------- rows E, M: whole ResNet-18 inference behind one call
 * custom_resnet18 (ss_baselines/savi/models/smt_resnet.py:56-164; SMTCNN encoders smt_cnn.py:78-115, belief
 * location head belief_predictor.py:64-72) and torchvision resnet18 with folded eval BatchNorm
 * (belief_predictor.py:74-82).  cfg: 12 host ints {norm_kind (0 GroupNorm, 1 folded BN), stem_k, stem_stride,
 * stem_pad, stem_maxpool, width0..3, groups, head_kind (0 Linear over the flattened map, 1 avg-pool + Linear),
 * out_dim}.  params: avl_resnet18_param_count() device pointers {stem w, a, b; 8 x (w1, a1, b1, w2, a2, b2, wd,
 * ad, bd); head w, b}; conv weights (Cout, KH, KW, Cin) when use_tc, OIHW otherwise.  _pair enqueues two
 * independent networks on two streams that fork from / join into `stream`.                                       */
int avl_resnet18_param_count(void);
long long avl_resnet18_workspace_bytes(int N, int H, int W, const int* cfg /* host */);
int avl_resnet18_forward(const float* x, int N, int H, int W, int Cin, const int* cfg /* host */, float eps,
                         const float* const* params /* host array */, float* out, long long ldo, int use_tc,
                         void* workspace, void* stream);
int avl_resnet18_forward_pair(const float* x0, const float* x1, int N, int H, int W, int Cin0, int Cin1,
                              const int* cfg0, const int* cfg1, float eps, const float* const* params0,
                              const float* const* params1, float* out0, float* out1, long long ldo0, long long ldo1,
                              int use_tc, void* workspace0, void* workspace1, void* stream);

/* ----------------------------------------------------------------------------------- dense building blocks
 * C[M,N] (+)= sum_k A(m,k) B(n,k) with element strides; bias / ReLU / accumulate / split-K (atomic).           */
int avl_gemm(const float* A, long long sa_m, long long sa_k, const float* B, long long sb_n, long long sb_k, float* C,
             long long ldc, int M, int N, int K, const float* bias, int relu, int accumulate, int splits,
             void* stream);
/* tcgen05 TF32: C = act(scale * A[M,K] B[N,K]^T + bias + residual); m_dev = optional device row count.        */
int avl_tc_gemm(const float* A, long long lda, const float* B, float* C, long long ldc, int M, int N, int K,
                const float* scale, const float* bias, const float* residual, long long ldr, int relu,
                const int* m_dev, void* stream);
/* fp32-accurate (3xTF32: operands split into two TF32 numbers, three MMAs per K slice, fp32 accumulate) dense GEMM
 * on tcgen05 fed by TMA — the scene-memory transformer's linears (reference: fp32 torch.matmul).  b_transposed: B is
 * stored [K][N] with row stride ldb.  -2 when the shape / alignment is not covered (nothing launched).               */
int avl_tc_gemm_3x(const float* A, long long lda, const float* B, long long ldb, int b_transposed, float* C,
                   long long ldc, int M, int N, int K, const float* bias, const float* residual, long long ldr,
                   int relu, const int* m_dev, void* stream);
int avl_set_tc_3xtf32(int on);   /* returns old */
/* Weight gradient of a Linear, dW[N,K] += dY[rows,N]^T X[rows,K] (reference: autograd of F.linear, fp32), 3xTF32 on
 * tcgen05 with both operands MN-major straight from TMA; the reduction over rows is split over the grid and summed in a
 * fixed order (deterministic).  rows_dev = optional device-side row count.  -2: shape / alignment not covered.    */
int avl_tc_wgrad_3x(const float* dY, long long ldy, const float* X, long long ldx, float* dW, long long lddw, int rows,
                    int N, int K, const int* rows_dev, void* stream);
int avl_set_wgrad_desc(int lbo_bytes, int sbo_bytes);   /* diagnostic */
/* fp16 ACTIVATION STORAGE for the widest encoder tensors (stem output + stage 1 of custom_resnet18, smt_resnet.py:56-164):
 * avl_resnet18_forward keeps them in HBM as fp16 (10-bit mantissa = what the TF32 tensor core keeps of an fp32 operand),
 * fp32 accumulation and fp32 GroupNorm arithmetic; avl_set_f16_activations(0) restores fp32 storage.  The two typed
 * entries below expose the kernels for tests / benches: x (and w, packed (Cout,KH,KW,C)) fp16 when in16, y fp16 when
 * out16; stride 1, pad = K/2.  -2: shape not covered.                                                             */
int avl_set_f16_activations(int on);   /* returns old */
/* fp16 has a 5-bit exponent: the convolutions that write fp16 saturate to +-65504 (never inf) and the GroupNorm that
 * reads the tensor raises a sticky flag when it meets a saturated value.  Returns 1 if that happened since the last
 * reset (synchronises the device; the host then falls back to fp32 storage).                                       */
int avl_f16_overflow(int reset);
/* Whole-network calls (avl_resnet18_forward / _pair) at batch <= 512 whose arguments repeat are captured into a CUDA
 * graph the second time they are seen and replayed afterwards (1 graph launch instead of ~50-100 kernel launches).  */
int avl_set_resnet_graphs(int on);     /* returns old */
long long avl_resnet_graph_stats(int what);   /* 0: replays, 1: captures */
int avl_set_tc_conv_halo_stride2(int on); /* stride-2 same-padded convs on the halo-strip kernel (stride-1 strip, even positions stored); returns old */
int avl_set_tc_conv_halo_group(int on); /* halo-strip conv: 2 / 4 adjacent pixels per MMA row when Cout <= 32; returns old */
int avl_tc_conv_halo_f16(const void* x, int in16, int N, int H, int W, int C, const void* w_packed, int Cout, int KH,
                         int KW, int pad, int relu, void* y, int out16, void* stream);
int avl_groupnorm_fwd_cluster_f16(const void* x, const float* gamma, const float* beta, const void* residual, void* y,
                                  int out16, int N, int HW, int C, int groups, float eps, int relu, void* stream);
int avl_layernorm_fwd(const float* x, const float* res, const float* gamma, const float* beta, float* y,
                      float* stats, int rows, int cols, void* stream);
int avl_layernorm_bwd(const float* x, const float* res, const float* gamma, const float* stats, const float* dy,
                      float* dx, float* dgamma, float* dbeta, int rows, int cols, void* stream);
int avl_attn_self_fwd(const float* qkv, const int* off, int B, int D, float* out, float* lse, void* stream);
int avl_attn_self_bwd(const float* qkv, const int* off, int B, int D, const float* out, const float* lse,
                      const float* dout, float* dqkv, void* stream);
int avl_attn_cross_fwd(const float* q, const float* kv, const int* off, int B, int D, float* out, float* probs,
                       void* stream);
int avl_attn_cross_bwd(const float* q, const float* kv, const int* off, const float* probs, const float* dout, int B,
                       int D, float* dq, float* dkv, void* stream);

/* ------------------------------------------------------------------- row F: scene-memory transformer encoder
 * ss_baselines/savi/models/smt_state_encoder.py:109-276 around torch.nn.Transformer (1+1 layers, post-norm).
 * params / grads: host arrays of avl_smt_param_count() device pointers, order documented in
 * avlen_b200/savi/models/smt_state_encoder.py::SMT_PARAM_KEYS; gradients are accumulated (NULL = skip).         */
int avl_smt_param_count(void);
long long avl_smt_workspace_bytes(int B, int rows_cap, int F, int D, int with_backward, int need_dx);
int avl_smt_forward(int B, int M, int F, int D, int pose_index, int pretraining, int rows_cap, const float* x,
                    const float* memory, int n_mem_envs, const int* env_index, const float* masks, const float* goal,
                    const float* const* params /* host array */, float* out, void* workspace, int with_backward,
                    int need_dx, void* stream);
int avl_smt_backward(int B, int M, int F, int D, int pose_index, int rows_cap, const float* goal,
                     const float* const* params /* host array */, float* const* grads /* host array */,
                     const float* gout, float* dx, float* dgoal, void* workspace, void* stream);
int avl_smt_status(int B, int rows_cap, int F, int D, void* workspace, int* total_rows /* host */,
                   int* overflow /* host */);   /* synchronises */

/* ------------------------------------------------------------ rows C, D, E backward (encoders are trained in
 * savi_pretraining.yaml / av_nav; the reference gets these from autograd over nn.Conv2d / nn.GroupNorm).
 * dgrad: dx (N,H,W,C) (+)= ; wgrad: dw (Cout,C,KH,KW) += , dbias (Cout) += (either may be NULL).                 */
int avl_conv2d_dgrad(const float* dy, const float* w_oihw, float* dx, int N, int H, int W, int C, int Cout, int KH,
                     int KW, int stride, int pad, int accumulate, void* stream);
int avl_conv2d_wgrad(const float* x, const float* dy, float* dw, float* dbias, int N, int H, int W, int C, int Cout,
                     int KH, int KW, int stride, int pad, void* stream);
/* Tensor-core backward of the convolutions (csrc/conv_bwd_tc.cu; replaces autograd of nn.Conv2d in the trainable-encoder
 * regime: smt_resnet.py:132-149 under savi_pretraining.yaml:53).
 *   avl_pack_conv_weight : OIHW weight -> (Cout, KH, KW, pad_to) [mode 0, forward] or the flipped / channel-transposed
 *                          (C, KH, KW, pad_to) weight of the data-gradient convolution [mode 1], TF32-rounded, one launch;
 *   avl_zero_upsample2   : (N, OH, OW, C) -> (N, H, W, C), values at the even positions (stride-2 data gradients: dx is
 *                          then a stride-1 avl_tc_conv2d_fwd with the mode-1 weight);
 *   avl_tc_conv2d_wgrad  : dw (Cout, Cw, KH, KW) = sum_{n,oh,ow} dy x x, x (N, H, W, Cx) with Cx == 4 or Cx % 8 == 0 and
 *                          Cx >= Cw, Cout % 16 == 0, stride 1 or 2; TF32 mma, fp32 accumulate, deterministic;
 *                          workspace floats from avl_tc_conv2d_wgrad_workspace (-2: shape not covered).                 */
int avl_pack_conv_weight(const float* w_oihw, int Cout, int C, int KH, int KW, int pad_to, int mode, float* out,
                         void* stream);
int avl_zero_upsample2(const float* dy, float* up, int N, int OH, int OW, int C, int H, int W, void* stream);
long long avl_tc_conv2d_wgrad_workspace(int N, int H, int W, int Cx, int Cout, int KH, int KW, int stride, int pad);
int avl_tc_conv2d_wgrad(const float* x, const float* dy, float* dw, int N, int H, int W, int Cx, int Cw, int Cout, int KH,
                        int KW, int stride, int pad, float* workspace, long long ws_floats, void* stream);
int avl_relu_mask(float* dy, long long ldd, const float* y, long long ldy, long long rows, int cols, void* stream);
int avl_groupnorm_bwd(const float* x, const float* y, const float* dy, const float* gamma, float* dx, float* dres,
                      float* dgamma, float* dbeta, int N, int HW, int C, int groups, float eps, int relu,
                      void* stream);
/* The same in ONE pass over HBM (thread-block cluster per sample, x and the masked dy staged in shared memory, group sums
 * through distributed shared memory; per-sample parameter-gradient rows summed in order: deterministic).  dgamma / dbeta
 * are accumulated into; scratch: 2 * N * C floats.  -2: shape not covered (use avl_groupnorm_bwd).               */
int avl_groupnorm_bwd_cluster(const float* x, const float* y, const float* dy, const float* gamma, float* dx, float* dres,
                              float* dgamma, float* dbeta, int N, int HW, int C, int groups, float eps, int relu,
                              float* scratch, void* stream);
int avl_set_tc_conv_halo(int on, int rows_per_strip); /* halo-strip kernel for stride-1 same convs; returns old */
int avl_set_tc_tma(int on);      /* dense GEMMs: 1 TMA-fed kernel where it applies (default), 0 cp.async kernel; returns old */
int avl_set_tc_swizzle(int on);  /* generic kernel operand tiles: 1 SWIZZLE_128B (default), 0 SWIZZLE_NONE; returns old */
int avl_set_tc_splitk(int on);   /* split-K for small-M / long-K tensor-core problems; returns old */
int avl_set_tc_stages(int stages); /* ring depth of the generic tensor-core kernel: 0 (default) automatic, 3 / 4 forced (diagnostic); returns old */
int avl_set_tc_splitk_cluster(int on); /* 1 (default): the k-slices of a tile form a thread-block cluster, partial tiles are summed in slice order through distributed shared memory inside the kernel (deterministic, no helper launches); 0: atomic partial sums + separate zero / epilogue kernels; returns old */
int avl_set_tc_conv_l1(int on);   /* im2col gathers through L1 (cp.async.ca, default) or L2 only; returns old */
int avl_set_tc_conv_tma(int on);  /* 1 (default): convolutions with >= 16 input channels and enough output tiles are fed by TMA in im2col mode (cuTensorMapEncodeIm2col, cp.async.bulk.tensor...im2col); 0: the cp.async gather everywhere; returns old */
/* Data gradient of a 3x3 stride-2 pad-1 convolution (stage-entry convolutions of the ResNet-18s, smt_resnet.py:132-149 under
 * autograd) as one 2x2-tap TMA convolution of dy with a pixel-shuffle epilogue; w2 [4*Cin][4*Cout] (row (pa, pb, ci), column
 * (u, v, co)); dy (N, OH, OW, Cout) -> dx (N, 2 OH, 2 OW, Cin).  -2: shape not covered (zero-upsample path).                 */
int avl_tc_conv2d_dgrad_s2(const float* dy, int N, int OH, int OW, int Cout, const float* w2, int Cin, float* dx,
                           void* stream);
long long avl_tc_conv_tma_count(void); /* convolutions launched on the TMA im2col kernel so far (diagnostic) */
int avl_set_tc_conv_halo_tma(int on); /* 1 (default): the halo-strip kernel's input strips arrive by TMA (rank-5 tiled map, halo zero-filled by the unit); 0: cp.async gathers; returns old */
int avl_set_pdl(int on);          /* 1 (default): the convolution / GroupNorm kernels of the encoder chains are launched with programmatic stream serialization (their prologues overlap the previous kernel; every kernel waits with griddepcontrol.wait before touching activations); 0: plain launches; returns old */
int avl_set_tc_splitk_fill(int percent); /* split-K of the small-grid convolutions aims at this many CTAs per 100 SMs (25..400; default 40: a rollout step is bound by SM time across its concurrent encoder chains, measured 53.5k -> 57.9k env-steps/s against 200; on the final build 50 / 40 / 30: 62.6k / 63.7k / 63.7k); returns old */
int avl_set_wide_stores(int on); /* 1 (default): 32-byte global stores (STG.256) in the row-per-thread TMEM epilogues of the conv / GEMM kernels; returns old */
int avl_set_tc_conv_halo_small_grid(int percent); /* grid cap of the halo-strip kernel at rollout batches, CTAs per 100 SMs (default 100); 0: none; returns old */
int avl_set_attn_tc(int on);
int avl_set_attn_qsplit(int n); /* CTAs per (sample, head) of the tensor-core self-attention forward: 0 automatic (2 at rollout batches), 1..4 forced; returns old */      /* 1 (default): varlen self-attention (row F, smt_state_encoder.py:160-166) as 3xTF32 warp MMAs whenever the tensor-core level is >= 1; 0: register-tiled fp32 kernels; returns old */

/* ----------------------------------------------------------------------------- row H: GRU state encoder
 * ss_baselines/av_nav/models/rnn_state_encoder.py:80-149 (single_forward T=1 / seq_forward) around
 * nn.GRU(I -> H, 1 layer).  x (T*N, I) time-major; masks (T*N) float, 0 at episode starts (h_{t-1} * mask_t);
 * weights in nn.GRU layout (gate order r, z, n).  Parameter gradients are accumulated.                          */
long long avl_gru_workspace_bytes(int T, int N, int I, int H, int with_backward);
int avl_gru_forward(int T, int N, int I, int H, const float* x, const float* h0, const float* masks,
                    const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, float* out,
                    float* h_last, void* workspace, int with_backward, void* stream);
int avl_gru_backward(int T, int N, int I, int H, const float* x, const float* masks, const float* w_ih,
                     const float* w_hh, const float* dout, const float* dh_last, float* dx, float* dh0, float* dw_ih,
                     float* dw_hh, float* db_ih, float* db_hh, void* workspace, void* stream);

/* ------------------------------------------------------------------- row R: PPO.update_dialog imitation loss
 * ss_baselines/savi/ppo/ppo.py:134-142: rows = nonzero(o_masks); CrossEntropyLoss(weight)(logits[rows],
 * o_actions[rows]) forward + gradient in one launch, no host-side row selection.  out3 = {loss, sum w, #rows}.   */
int avl_masked_weighted_ce(const float* logits, const float* targets, const long long* mask, const float* weight,
                           int B, int A, float* dlogits, float* out3, void* stream);

/* ------------------------------------------------------------------ row K: dialog state encoder (pi_l)
 * ss_baselines/savi/models/dialog_state_encoder.py:114-155 (+ PositionalEncoding :18-40): tokens = valid slots of
 * the K-slot state memory + the current SMT output; with a dialog each token is fused with the dialog embedding
 * (512 -> 256 ReLU -> 256); + pe[agent_step]; nn.Transformer(1+1) decoded with the belief vector.
 * params / grads: avl_dialog_param_count() device pointers (order: avlen_b200/savi/models/
 * dialog_state_encoder.py::DIALOG_PARAM_KEYS).  d_emb NULL = no dialog (fusion skipped, :138).                   */
int avl_dialog_param_count(void);
long long avl_dialog_workspace_bytes(int B, int K, int D, int with_backward);
int avl_dialog_forward(int B, int K, int D, const float* x_att, const float* memory_state, int n_mem_envs,
                       const int* env_index, const float* masks, const float* d_emb, const int* agent_step,
                       const float* pe_table, int pe_len, const float* goal,
                       const float* const* params /* host array */, float* out, void* workspace, int with_backward,
                       void* stream);
int avl_dialog_backward(int B, int K, int D, int has_dialog, const float* goal,
                        const float* const* params /* host array */, float* const* grads /* host array */,
                        const float* gout, float* dx_att, float* dd_emb, float* dgoal, void* workspace,
                        void* stream);

/* --------------------------------------------------------------------------- row L: CLIP text tower (pi_l)
 * ss_baselines/savi/ppo/policy.py:761-762 (clip.load("ViT-B/32")), :844-851 (encode_text(all_dialog).float(),
 * no_grad, frozen).  tokens (B, L<=77) int64 in clip.tokenize layout; out (B, 512) fp32.  dedupe != 0: all-zero
 * rows (envs without an active query) are encoded once.  params: avl_clip_text_param_count(layers) pointers in
 * the order of avlen_b200/savi/models/clip_text.py::clip_param_keys.                                            */
int avl_clip_text_param_count(int layers);
long long avl_clip_text_workspace_bytes(int B, int L);
int avl_clip_text_forward(int B, int L, int vocab, int layers, const long long* tokens,
                          const float* const* params /* host array */, float* out, void* workspace, int dedupe,
                          void* stream);
/* The same tower with its four linears per layer on fp16 operands (tcgen05 kind::f16, fp32 accumulation / residual stream /
 * LayerNorm and softmax statistics) — the dtype `clip.load` gives the tower on CUDA in the reference (policy.py:761).
 * params16: 4 * layers device pointers to fp16 copies of (in_proj_weight, out_proj.weight, c_fc.weight, c_proj.weight). */
int avl_clip_text_forward_f16(int B, int L, int vocab, int layers, const long long* tokens, const float* const* params,
                              const void* const* params16, float* out, void* workspace, int dedupe, void* stream);
int avl_clip_text_status(int B, int L, void* workspace, int* n_sequences /* host */, int* n_rows /* host */);

#ifdef __cplusplus
}
#endif
#endif /* AVLEN_B200_H */
