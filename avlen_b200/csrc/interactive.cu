// Device-side query / option bookkeeping of the AVLEN interactive rollout step: the per-env
// Python loops of PPOTrainer._collect_rollout_step (ss_baselines/savi/ppo/ppo_trainer.py:394-416, :449-460, :487-588,
// :639-694, :769-787) with their .item() / .cpu() round trips, as three small kernels over all envs — integer /
// index / mask work, bit-exact against a trace of the unmodified reference method (tests/golden/interactive_step.npz).
//
// Per-env state (int32, struct of arrays `st[5][n]`): queried, step (inside the current dialog), total_step (of the
// episode), last_query_step, query count; plus the token row of the env's current dialog (n, L) int64.
#include "common.cuh"

#ifndef AVL_HOST_EMUL
namespace {

enum { Q_QUERIED = 0, Q_STEP = 1, Q_TOTAL = 2, Q_LAST = 3, Q_COUNT = 4 };

// :394-416 — episode reset or step advance; query_state = pe[count], last_query_info = pe[diff_step]
__global__ void query_pre_kernel(int n, const unsigned char* __restrict__ new_episode, int* st, const float* __restrict__ pe,
                                 int pe_rows, int emb, float* query_state, float* last_query_info) {
  const int i = blockIdx.x;
  __shared__ int s_count, s_diff;
  if (threadIdx.x == 0) {
    int diff;
    if (new_episode[i]) {
      st[Q_QUERIED * n + i] = 0;
      st[Q_STEP * n + i] = 0;
      st[Q_TOTAL * n + i] = 0;
      st[Q_LAST * n + i] = 0;
      st[Q_COUNT * n + i] = 0;
      diff = 150;
    } else {
      const int total = st[Q_TOTAL * n + i] + 1;
      st[Q_TOTAL * n + i] = total;
      diff = st[Q_COUNT * n + i] >= 2 ? total - st[Q_LAST * n + i] : 150;
    }
    s_count = min(st[Q_COUNT * n + i], pe_rows - 1);
    s_diff = min(max(diff, 0), pe_rows - 1);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < emb; c += blockDim.x) {
    query_state[(size_t)i * emb + c] = pe[(size_t)s_count * emb + c];
    last_query_info[(size_t)i * emb + c] = pe[(size_t)s_diff * emb + c];
  }
}

// :449-460 + :487-588 — a query fires, consecutive-query penalty, rl_mask, the dialog row / agent step pi_l sees
__global__ void query_after_option_kernel(int n, const long long* __restrict__ actions_option,
                                          const float* __restrict__ target_distance,
                                          const long long* __restrict__ pending_dialog, int L, int num_dialog_steps,
                                          float consecutive_reward, int query_within_radius, int* st, long long* dialog_store,
                                          unsigned char* is_queried, long long* query_num, float* cons_reward,
                                          long long* rl_mask, long long* cur_dialog, float* agent_step) {
  const int i = blockIdx.x;
  __shared__ int s_copy_new, s_emit;
  if (threadIdx.x == 0) {
    int queried = st[Q_QUERIED * n + i], count = st[Q_COUNT * n + i], step = st[Q_STEP * n + i];
    if (!queried && actions_option[i] == 1 && (query_within_radius || target_distance[i] > 3.f)) {
      queried = 1;
      ++count;
    }
    float cons = 0.f;
    long long rl = 1;
    int copy_new = 0, emit = 0;
    float astep = 0.f;
    if (queried) {
      if (step == 0) {
        const int total = st[Q_TOTAL * n + i];
        if (count >= 2) {
          const int d = total - (st[Q_LAST * n + i] + 2);
          cons = d > 10 ? 0.f : consecutive_reward / (float)max(d, 1);
        }
        st[Q_LAST * n + i] = total;
        rl = 1;
        copy_new = 1;  // the speaker's instruction becomes the env's dialog
      } else {
        rl = 0;
      }
      if (step < num_dialog_steps) {
        emit = 1;
        astep = (float)step;
        ++step;
      }
    }
    st[Q_QUERIED * n + i] = queried;
    st[Q_COUNT * n + i] = count;
    st[Q_STEP * n + i] = step;
    is_queried[i] = (unsigned char)queried;
    query_num[i] = count;
    cons_reward[i] = cons;
    rl_mask[i] = rl;
    agent_step[i] = astep;
    s_copy_new = copy_new;
    s_emit = emit;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < L; c += blockDim.x) {
    long long tok = dialog_store[(size_t)i * L + c];
    if (s_copy_new) {
      tok = pending_dialog[(size_t)i * L + c];
      dialog_store[(size_t)i * L + c] = tok;
    }
    cur_dialog[(size_t)i * L + c] = s_emit ? tok : 0;
  }
}

// :639-694 + :769-787 — ucnt_gt, action arbitration, o_mask; end of a dialog (masks_vln)
__global__ void option_arbitrate_kernel(int n, const long long* __restrict__ actions_goal,
                                        const long long* __restrict__ actions_vln, const float* __restrict__ probs_goal, int A,
                                        const long long* __restrict__ oracle, int oracle_when_queried, int allow_stop,
                                        int num_dialog_steps, int* st, long long* dialog_store, int L, long long* actions,
                                        long long* o_mask, long long* ucnt_gt, float* masks_vln) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // top-2 of pi_g's probabilities (np.sort(...)[-1] - [-2] < 0.1)
  float m1 = -INFINITY, m2 = -INFINITY;
  for (int a = 0; a < A; ++a) {
    const float v = probs_goal[(size_t)i * A + a];
    if (v > m1) { m2 = m1; m1 = v; }
    else if (v > m2) m2 = v;
  }
  ucnt_gt[i] = (m1 - m2 < 0.1f) ? 1 : 0;
  const int queried = st[Q_QUERIED * n + i];
  long long act, om;
  if (queried) {
    const long long o = oracle[i];
    if (o == 0) {
      act = oracle_when_queried ? (allow_stop ? o : actions_vln[i]) : o;
      om = 0;
    } else {
      act = oracle_when_queried ? o : actions_vln[i];
      om = 1;
    }
  } else {
    act = actions_goal[i];
    om = 1;
  }
  actions[i] = act;
  o_mask[i] = om;
  float mv = 1.f;
  if (queried && st[Q_STEP * n + i] >= num_dialog_steps) {
    st[Q_QUERIED * n + i] = 0;
    st[Q_STEP * n + i] = 0;
    for (int c = 0; c < L; ++c) dialog_store[(size_t)i * L + c] = 0;
    mv = 0.f;
  }
  masks_vln[i] = mv;
}

}  // namespace

AVL_API int avl_query_pre(int n, const unsigned char* new_episode, int* state, const float* pe, int pe_rows, int emb,
                          float* query_state, float* last_query_info, void* stream) {
  if (n < 0 || pe_rows < 151 || emb < 1) return AVL_ERR_ARG;
  if (n == 0) return AVL_OK;
  if (!new_episode || !state || !pe || !query_state || !last_query_info) return AVL_ERR_ARG;
  query_pre_kernel<<<n, 32, 0, (cudaStream_t)stream>>>(n, new_episode, state, pe, pe_rows, emb, query_state, last_query_info);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}

AVL_API int avl_query_after_option(int n, const long long* actions_option, const float* target_distance,
                                   const long long* pending_dialog, int L, int num_dialog_steps, float consecutive_reward,
                                   int query_within_radius, int* state, long long* dialog_store, unsigned char* is_queried,
                                   long long* query_num, float* cons_reward, long long* rl_mask, long long* cur_dialog,
                                   float* agent_step, void* stream) {
  if (n < 0 || L < 1 || num_dialog_steps < 1) return AVL_ERR_ARG;
  if (n == 0) return AVL_OK;
  if (!actions_option || !target_distance || !pending_dialog || !state || !dialog_store || !is_queried || !query_num ||
      !cons_reward || !rl_mask || !cur_dialog || !agent_step)
    return AVL_ERR_ARG;
  query_after_option_kernel<<<n, 96, 0, (cudaStream_t)stream>>>(n, actions_option, target_distance, pending_dialog, L,
                                                               num_dialog_steps, consecutive_reward, query_within_radius,
                                                               state, dialog_store, is_queried, query_num, cons_reward,
                                                               rl_mask, cur_dialog, agent_step);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}

AVL_API int avl_option_arbitrate(int n, const long long* actions_goal, const long long* actions_vln, const float* probs_goal,
                                 int A, const long long* oracle, int oracle_when_queried, int allow_stop,
                                 int num_dialog_steps, int* state, long long* dialog_store, int L, long long* actions,
                                 long long* o_mask, long long* ucnt_gt, float* masks_vln, void* stream) {
  if (n < 0 || A < 2 || L < 1) return AVL_ERR_ARG;
  if (n == 0) return AVL_OK;
  if (!actions_goal || !actions_vln || !probs_goal || !oracle || !state || !dialog_store || !actions || !o_mask || !ucnt_gt ||
      !masks_vln)
    return AVL_ERR_ARG;
  option_arbitrate_kernel<<<avl_div_up(n, 128), 128, 0, (cudaStream_t)stream>>>(
      n, actions_goal, actions_vln, probs_goal, A, oracle, oracle_when_queried, allow_stop, num_dialog_steps, state,
      dialog_store, L, actions, o_mask, ucnt_gt, masks_vln);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}
#endif  // AVL_HOST_EMUL

// ====================================================================================================================
// Graph-walk environment step for all envs in one launch: the SoundSpaces navigation graph
// walk (soundspaces/simulator.py:496-517), the first oracle action of the shortest path (:758-787), the reward
// (ss_baselines/common/environments.py:98-135) and the episode bookkeeping / auto-reset of the VectorEnv, on
// per-scene tables resident in HBM: nbr[V][4] (neighbour in direction 0 / 90 / 180 / 270 degrees or -1), hops[V][V]
// (geodesic distance in edges), next_dir[target][node] (direction of the first edge of the shortest path).
// Angles are kept as quarter turns: rotation r <-> rotation_angle = 90 r; orientation = (270 - 90 r) % 360 (:594-596).
#ifndef AVL_HOST_EMUL
namespace {

struct GraphEnvArgs {
  int n, V;
  const int* nbr;
  const short* hops;
  const signed char* next_dir;
  const float* points;  // (V, 2): x, z
  float grid_size;
  // episode table: per env a ring of E episodes (start node, start rotation, source node)
  const int* ep_start;
  const int* ep_rot;
  const int* ep_source;
  int E;
  // state
  int* node;
  int* rot;
  int* source;
  int* ep_step;
  int* ep_cursor;
  int* start_node;
  int* start_rot;
  float* prev_dist;
  // inputs of this step
  const long long* actions;
  const unsigned char* is_queried;
  const long long* query_num;
  const float* cons_reward;
  // reward switches (RL.* of the yaml)
  float slack_reward, distance_scale, success_reward, query_reward;
  int with_time_penalty, with_distance_reward, with_query_constraint, consecutive_constraint, soft_query_reward;
  int num_total_query, max_steps;
  // outputs
  float* rewards;
  unsigned char* dones;
  float* masks;
  float* pose;           // (n, 4)
  long long* oracle;     // first oracle action from the NEW state
  float* target_distance;
  unsigned char* new_episode;
  int* azimuth;          // (n,) azimuth_angle / 90 of the new state (:598-603), indexes the RIR bank
};

__device__ __forceinline__ int orientation_q(int rot) { return ((3 - rot) % 4 + 4) % 4; }  // (270 - 90 r) % 360, in quarter turns

__device__ __forceinline__ long long first_oracle_action(const GraphEnvArgs& a, int node, int rot, int source) {
  if (node == source) return 0;  // STOP (the path has no edge left, :787)
  const int dir = a.next_dir[(size_t)source * a.V + node];
  if (dir < 0) return 0;
  const int d = ((dir - orientation_q(rot)) % 4 + 4) % 4;  // (direction - orientation) % 360 in quarter turns
  if (d == 0) return 1;   // MOVE_FORWARD
  if (d == 3) return 2;   // 270 -> TURN_LEFT  (:774-776)
  return 3;               // 90 -> TURN_RIGHT ; 180 -> TURN_RIGHT (twice) (:777-783)
}

__global__ void graph_env_step_kernel(GraphEnvArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  int node = a.node[i], rot = a.rot[i];
  const int source = a.source[i];
  const long long act = a.actions[i];
  bool stop = false;
  if (act == 0) {
    stop = true;                                   // STOP: the episode is over (:497-498)
  } else if (act == 1) {                           // MOVE_FORWARD: the neighbour that lies in the facing direction
    const int nb = a.nbr[node * 4 + orientation_q(rot)];
    if (nb >= 0) node = nb;                        // (else: collision, the agent stays)
  } else if (act == 2) {
    rot = (rot + 1) & 3;                           // TURN_LEFT: rotation + 90 (:512-513)
  } else if (act == 3) {
    rot = (rot + 3) & 3;                           // TURN_RIGHT: rotation - 90
  }
  const int step = a.ep_step[i] + 1;
  // ---- reward (environments.py:98-135)
  float reward = 0.f;
  if (a.with_time_penalty) reward += a.slack_reward;
  const float dist = (float)a.hops[(size_t)source * a.V + node] * a.grid_size;
  if (a.with_distance_reward) reward += (a.prev_dist[i] - dist) * a.distance_scale;
  if (stop && node == source) reward += a.success_reward;
  if (a.with_query_constraint && a.is_queried && a.is_queried[i]) {
    const long long qn = a.query_num[i];
    if (qn <= a.num_total_query) {
      if (a.soft_query_reward)
        reward += ((float)qn / (float)a.num_total_query) * (expf(-(float)a.num_total_query) + a.query_reward);
    } else {
      reward += expf(-(float)qn) + a.query_reward;
    }
    if (a.consecutive_constraint) reward += a.cons_reward[i];
  }
  const bool done = stop || step >= a.max_steps;
  a.rewards[i] = reward;
  a.dones[i] = done ? 1 : 0;
  a.masks[i] = done ? 0.f : 1.f;
  int t = step, src = source;
  float pd = dist;
  if (done) {  // VectorEnv auto-reset: the observation that follows belongs to the next episode of this env
    const int c = (a.ep_cursor[i] + 1) % a.E;
    a.ep_cursor[i] = c;
    node = a.ep_start[(size_t)i * a.E + c];
    rot = a.ep_rot[(size_t)i * a.E + c] & 3;
    src = a.ep_source[(size_t)i * a.E + c];
    a.start_node[i] = node;
    a.start_rot[i] = rot;
    a.source[i] = src;
    t = 0;
    pd = (float)a.hops[(size_t)src * a.V + node] * a.grid_size;  // environments.py:66-68
  }
  a.node[i] = node;
  a.rot[i] = rot;
  a.ep_step[i] = t;
  a.prev_dist[i] = pd;
  a.new_episode[i] = done ? 1 : 0;
  a.target_distance[i] = pd;
  a.oracle[i] = first_oracle_action(a, node, rot, src);
  a.azimuth[i] = ((-rot) % 4 + 4) % 4;  // azimuth_angle = -(rotation_angle) % 360 (:598-603)
  // ---- pose (PoseSensor, soundspaces/tasks/nav.py:745-775): position relative to the episode start in the start frame,
  // heading relative to the start rotation, episode time
  const int sn = a.start_node[i], sr = a.start_rot[i];
  const float dx = a.points[node * 2] - a.points[sn * 2], dz = a.points[node * 2 + 1] - a.points[sn * 2 + 1];
  const float th = 1.5707963267948966f * (float)sr;
  const float cs = cosf(th), sn_ = sinf(th);
  // rotate (dx, dz) by the inverse start rotation about +Y
  const float rx = cs * dx - sn_ * dz, rz = sn_ * dx + cs * dz;
  int dr = ((sr - rot) % 4 + 4) % 4;
  float heading = 1.5707963267948966f * (float)(dr > 2 ? dr - 4 : dr);
  a.pose[(size_t)i * 4 + 0] = -rz;
  a.pose[(size_t)i * 4 + 1] = rx;
  a.pose[(size_t)i * 4 + 2] = heading;
  a.pose[(size_t)i * 4 + 3] = (float)t;
}

}  // namespace

// All pointers device.  ``iargs``: V, E, with_time_penalty, with_distance_reward, with_query_constraint,
// consecutive_constraint, soft_query_reward, num_total_query, max_steps ; ``fargs``: grid_size, slack_reward,
// distance_scale, success_reward, query_reward.  state: node, rot, source, ep_step, ep_cursor, start_node, start_rot (int32 n each).
AVL_API int avl_graph_env_step(int n, const int* iargs, const float* fargs, const int* nbr, const short* hops,
                               const signed char* next_dir, const float* points, const int* ep_start, const int* ep_rot,
                               const int* ep_source, int* state, float* prev_dist, const long long* actions,
                               const unsigned char* is_queried, const long long* query_num, const float* cons_reward,
                               float* rewards, unsigned char* dones, float* masks, float* pose, long long* oracle,
                               float* target_distance, unsigned char* new_episode, int* azimuth, void* stream) {
  if (n < 0 || !iargs || !fargs) return AVL_ERR_ARG;
  if (n == 0) return AVL_OK;
  if (!nbr || !hops || !next_dir || !points || !ep_start || !ep_rot || !ep_source || !state || !prev_dist || !actions ||
      !rewards || !dones || !masks || !pose || !oracle || !target_distance || !new_episode || !azimuth)
    return AVL_ERR_ARG;
  if (is_queried && (!query_num || !cons_reward)) return AVL_ERR_ARG;
  GraphEnvArgs a = {};
  a.n = n; a.V = iargs[0]; a.E = iargs[1];
  a.with_time_penalty = iargs[2]; a.with_distance_reward = iargs[3]; a.with_query_constraint = iargs[4];
  a.consecutive_constraint = iargs[5]; a.soft_query_reward = iargs[6]; a.num_total_query = iargs[7]; a.max_steps = iargs[8];
  a.grid_size = fargs[0]; a.slack_reward = fargs[1]; a.distance_scale = fargs[2]; a.success_reward = fargs[3];
  a.query_reward = fargs[4];
  if (a.V < 1 || a.E < 1 || a.max_steps < 1) return AVL_ERR_ARG;
  a.nbr = nbr; a.hops = hops; a.next_dir = next_dir; a.points = points;
  a.ep_start = ep_start; a.ep_rot = ep_rot; a.ep_source = ep_source;
  a.node = state; a.rot = state + n; a.source = state + 2 * n; a.ep_step = state + 3 * n; a.ep_cursor = state + 4 * n;
  a.start_node = state + 5 * n; a.start_rot = state + 6 * n;
  a.prev_dist = prev_dist; a.actions = actions; a.is_queried = is_queried; a.query_num = query_num; a.cons_reward = cons_reward;
  a.rewards = rewards; a.dones = dones; a.masks = masks; a.pose = pose; a.oracle = oracle; a.target_distance = target_distance;
  a.new_episode = new_episode; a.azimuth = azimuth;
  graph_env_step_kernel<<<avl_div_up(n, 128), 128, 0, (cudaStream_t)stream>>>(a);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}
#endif  // AVL_HOST_EMUL
