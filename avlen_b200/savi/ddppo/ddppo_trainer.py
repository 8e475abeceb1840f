"""DD-PPO trainer entry (ss_baselines/savi/ddppo/algo/ddppo_trainer.py:61-1200), registered as ``"ddppo"``:
``trainer = baseline_registry.get_trainer("ddppo")(config); trainer.train()`` (savi/run.py:110-121).

One process per GPU (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* from the environment, as torchrun sets them);
environments, rollout storage, external memories, belief state and audio rendering are rank-local (no exchange
during rollouts); the update exchanges ONE flat gradient all-reduce per minibatch over NCCL.  Straggler
preemption (ddppo_trainer.py:952-959) is kept as an option: ranks stop collecting once more than ``sync_frac`` of
the ranks have finished and at least a quarter of the rollout is done (counter in the torch.distributed store).
"""
from __future__ import annotations

import os
import time
import types

import torch
import torch.distributed as distrib

from ... import _lib
from ... import nn as K
from ...common import spaces
from ...common.baseline_registry import baseline_registry
from ...synth_env import SyntheticVectorEnv
from ..models.belief_predictor import BeliefPredictor
from ..models.rollout_storage import RolloutStorage
from ..ppo.policy import AudioNavDialogPolicy, AudioNavOptionPolicy, AudioNavSMTPolicy
from ..ppo.ppo_trainer import PPOTrainer
from ..ppo.query_state import QueryBookkeeper
from .ddppo import DDPPO


def savi_config(**overrides):
    """Values of ss_baselines/savi/config/semantic_audionav/savi.yaml (:10-66) as a flat namespace."""
    cfg = dict(NUM_PROCESSES=64, NUM_UPDATES=2, clip_param=0.2, ppo_epoch=2, num_mini_batch=2, value_loss_coef=0.5,
               entropy_coef=0.05, lr=2.5e-4, eps=1e-5, max_grad_norm=0.2, num_steps=150, hidden_size=512,
               use_gae=True, gamma=0.99, tau=0.95, use_normalized_advantage=False, policy_type="smt",
               use_belief_predictor=True, use_external_memory=True, memory_size=150, smt_hidden_size=256, nhead=8,
               num_encoder_layers=1, num_decoder_layers=1, dropout=0.0, activation="relu", freeze_encoders=True,
               pretraining=False, use_label_belief=True, use_location_belief=True, online_training=True,
               sync_frac=0.6, distrib_backend="nccl", use_preemption=False, seed=1234, sampling_rate=16000,
               host_buffers=False, has_distractor_sound=False, overlap_belief=True, prefetch_encoders=True,
               # AVLEN interactive stages (savi_interactive_2nd_stage.yaml:19-23,:30-41; config/default.py:183-186)
               NUM_DIALOG_STEPS=3, ORACLE_WHEN_QUERIED=True, QUERY_WITHIN_RADIUS=True, ALLOW_STOP=False,
               CONSECUTIVE_REWARD=-0.5, NUM_TOTAL_QUERY=3, QUERY_COUNT_EMB_SIZE=32, clip_layers=12, graph_env=None,
               # rgb uint8 / depth fp16 in the rollout storage and on the H2D path
               compact_observations=True,
               # whole-rollout-step CUDA graphs (frozen encoders, device-resident env): the ~230 launches of a step
               # replay as ONE graph launch per step from the third rollout on
               step_graphs=True)
    cfg.update(overrides)
    return types.SimpleNamespace(**cfg)


def init_distrib(backend="nccl"):
    """ddp_utils.py:129-182 without SLURM: rank / world from torchrun's environment variables."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not distrib.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        distrib.init_process_group(backend=backend, rank=rank, world_size=world)
    return local_rank, rank, world


@baseline_registry.register_trainer(name="ddppo")
class DDPPOTrainer(PPOTrainer):
    SHORT_ROLLOUT_THRESHOLD: float = 0.25

    def __init__(self, config=None):
        super().__init__(config or savi_config())
        self.rollouts = None
        self.world_size, self.world_rank = 1, 0

    def _policy_kwargs(self, cfg):
        return dict(hidden_size=cfg.smt_hidden_size, nhead=cfg.nhead, num_encoder_layers=cfg.num_encoder_layers,
                    num_decoder_layers=cfg.num_decoder_layers, dropout=cfg.dropout, activation=cfg.activation,
                    pretraining=cfg.pretraining)

    def _setup_belief_predictor(self, cfg):
        if cfg.use_belief_predictor:
            bcfg = types.SimpleNamespace(use_label_belief=cfg.use_label_belief, online_training=cfg.online_training,
                                         use_location_belief=cfg.use_location_belief, weighting_factor=0.5,
                                         current_pred_only=False)
            self.belief_predictor = BeliefPredictor(bcfg, self.device, None, None, cfg.smt_hidden_size,
                                                    cfg.NUM_PROCESSES, cfg.has_distractor_sound).to(self.device)
            self.belief_predictor.freeze_encoders()
            self.belief_predictor.set_eval_encoders()

    def _setup_actor_critic_agent(self, ppo_cfg, observation_space=None):
        cfg = ppo_cfg
        obs_space = observation_space or spaces.savi_observation_space(cfg.sampling_rate)
        self.obs_space = obs_space
        if cfg.policy_type == "interactive":
            return self._setup_interactive(cfg, obs_space)
        self.actor_critic = AudioNavSMTPolicy(
            observation_space=obs_space, action_space=spaces.Discrete(4), use_category_input=cfg.has_distractor_sound,
            **self._policy_kwargs(cfg))
        self.actor_critic.to(self.device)
        if cfg.freeze_encoders:
            self.actor_critic.net.freeze_encoders()
            self.actor_critic.net.set_eval_encoders()
        self.actor_critic.net.smt_state_encoder.rows_per_sample_cap = cfg.memory_size + 1
        self._setup_belief_predictor(cfg)
        self.agent = DDPPO(actor_critic=self.actor_critic, clip_param=cfg.clip_param, ppo_epoch=cfg.ppo_epoch,
                           num_mini_batch=cfg.num_mini_batch, value_loss_coef=cfg.value_loss_coef,
                           entropy_coef=cfg.entropy_coef, lr=cfg.lr, eps=cfg.eps, max_grad_norm=cfg.max_grad_norm,
                           use_normalized_advantage=cfg.use_normalized_advantage)

    def _setup_interactive(self, cfg, obs_space):
        """ddppo_trainer.py:301-512, ``policy_type: "interactive"``: pi_g (goal policy, every parameter frozen, :413-414),
        pi_l (dialog policy, CLIP frozen, :400-403) and pi_q (option policy) — ``self.agent`` trains pi_q, ``agent_vln``
        wraps pi_l for the replay / dialog updates; the sinusoidal query-count table ``pe`` (:506-512)."""
        kw = self._policy_kwargs(cfg)
        act = spaces.Discrete(4)
        dis = cfg.has_distractor_sound
        self.actor_critic_goal = AudioNavSMTPolicy(obs_space, act, use_category_input=dis, **kw).to(self.device)
        self.actor_critic_vln = AudioNavDialogPolicy(obs_space, act, clip_layers=cfg.clip_layers, **kw).to(self.device)
        self.actor_critic_option = AudioNavOptionPolicy(obs_space, act, use_category_input=dis, **kw).to(self.device)
        for p in self.actor_critic_goal.parameters():
            p.requires_grad = False
        for name, p in self.actor_critic_vln.named_parameters():
            if "net.clip" in name:
                p.requires_grad = False
        for pol in (self.actor_critic_goal, self.actor_critic_vln, self.actor_critic_option):
            if cfg.freeze_encoders:
                pol.net.freeze_encoders()
            pol.net.set_eval_encoders()
            pol.net.smt_state_encoder.rows_per_sample_cap = cfg.memory_size + 1
        self.actor_critic = self.actor_critic_option
        self._setup_belief_predictor(cfg)
        mk = lambda ac, head: DDPPO(actor_critic=ac, clip_param=cfg.clip_param, ppo_epoch=cfg.ppo_epoch,  # noqa: E731
                                    num_mini_batch=cfg.num_mini_batch, value_loss_coef=cfg.value_loss_coef,
                                    entropy_coef=cfg.entropy_coef, lr=cfg.lr, eps=cfg.eps, max_grad_norm=cfg.max_grad_norm,
                                    use_normalized_advantage=cfg.use_normalized_advantage, policy_head=head)
        self.agent = mk(self.actor_critic_option, "option")
        self.agent_vln = mk(self.actor_critic_vln, "goal")
        self.query_book = QueryBookkeeper(cfg.NUM_PROCESSES, self.device, num_dialog_steps=cfg.NUM_DIALOG_STEPS,
                                          consecutive_reward=cfg.CONSECUTIVE_REWARD,
                                          query_within_radius=cfg.QUERY_WITHIN_RADIUS,
                                          oracle_when_queried=cfg.ORACLE_WHEN_QUERIED, allow_stop=cfg.ALLOW_STOP,
                                          max_dialog_len=77, emb_size=cfg.QUERY_COUNT_EMB_SIZE)
        self.pe = self.query_book.pe

    def setup(self, envs=None):
        cfg = self.config
        local_rank, self.world_rank, self.world_size = init_distrib(cfg.distrib_backend)
        self.device = torch.device("cuda", local_rank)
        torch.cuda.set_device(self.device)
        if self.world_size > 1 and cfg.host_buffers:
            # one process per GPU on one host: the native frame-gather threads of all ranks share the host's cores
            local = int(os.environ.get("LOCAL_WORLD_SIZE", self.world_size))
            _lib.lib().avl_set_host_gather_threads(max(1, min(8, (os.cpu_count() or 8) // (2 * max(1, local)))))
        torch.manual_seed(cfg.seed + self.world_rank)
        interactive = cfg.policy_type == "interactive"
        if envs is None:
            kw = dict(seed=cfg.seed + self.world_rank, sr=cfg.sampling_rate, host_buffers=cfg.host_buffers,
                      distractor=cfg.has_distractor_sound, compact=cfg.compact_observations,
                      pool=5)  # (a frame pool whose size divides the 150-step rollout: step s always serves frame s % 5)
            if interactive or cfg.graph_env:
                # the interactive step needs what only a navigation graph provides: oracle actions, target distance,
                # episode boundaries, query-aware rewards (ppo_trainer.py:336-345,:642,:706-710)
                from ...graph_env import GraphVectorEnv
                envs = GraphVectorEnv(cfg.NUM_PROCESSES, self.device,
                                      reward=dict(NUM_TOTAL_QUERY=cfg.NUM_TOTAL_QUERY), **kw)
            else:
                envs = SyntheticVectorEnv(cfg.NUM_PROCESSES, self.device, **kw)
        self.envs = envs
        self._setup_actor_critic_agent(cfg)
        if self.world_size > 1:
            self.agent.init_distributed(find_unused_params=True)
            if interactive:
                self.agent_vln.init_distributed(find_unused_params=True)
        em_size = cfg.memory_size + cfg.num_steps  # ddppo_trainer.py:656-657
        if interactive:  # ddppo_trainer.py:640-676: goal / vln / option / dialog memories
            dg, dl = self.actor_critic_goal.net.memory_dim, self.actor_critic_vln.net.memory_dim
            dq = self.actor_critic_option.net.memory_dim
            self.rollouts = RolloutStorage(cfg.num_steps, self.envs.num_envs, self.obs_space, spaces.Discrete(4),
                                           cfg.hidden_size, cfg.use_external_memory, em_size, cfg.memory_size, em_size,
                                           cfg.memory_size, cfg.NUM_DIALOG_STEPS, cfg.NUM_DIALOG_STEPS, dg, dl, dq,
                                           cfg.smt_hidden_size, num_recurrent_layers=1, max_dialog_len=77,
                                           query_count_emb_size=cfg.QUERY_COUNT_EMB_SIZE, use_state_memory=True,
                                           compact_observations=cfg.compact_observations)
        else:
            dim = self.actor_critic.net.memory_dim
            self.rollouts = RolloutStorage(cfg.num_steps, self.envs.num_envs, self.obs_space, spaces.Discrete(4),
                                           cfg.hidden_size, cfg.use_external_memory, em_size, cfg.memory_size, em_size,
                                           cfg.memory_size, 3, 3, dim, dim, dim + 32, 256, num_recurrent_layers=1,
                                           max_dialog_len=77, compact_observations=cfg.compact_observations)
        self.rollouts.to(self.device)
        observations = self.envs.reset()
        if self.belief_predictor is not None:
            self.belief_predictor.update(observations, None)
        for sensor in self.rollouts.observations:
            self.rollouts.observations[sensor][0].copy_(observations[sensor])
        return self

    # ---- whole-step CUDA graphs -----------------------------------------------------------------------------------
    # A rollout step at 64 envs is ~230 small kernels on five streams, issued by ~25 ctypes / torch calls: the host
    # needs ~1.2 ms per step to issue them and its pauses stagger the four ResNet-18 chains by 150-470 us
    # (profiles/r02_trace_rollout_step_timeline.txt).  Every device address a step touches is fixed by its position
    # ``s`` in the rollout (storage slot s / s + 1, persistent env / belief / memory state, the ring position lives in
    # device memory), so step ``s`` is captured ONCE (during the second rollout, after an eager warm-up rollout) and
    # replayed afterwards: one graph launch per step, every dependency resolved on the device.
    # Conditions: plain SMT policy (or the interactive triple), no preemption.  Trainable encoders qualify because their
    # packed weights keep their addresses (re-packed in place, the re-pack after an optimizer step being part of the first
    # step's graph: the capture happens during the second rollout, i.e. right after an update); host frames: two graphs
    # per step (below).
    def _step_graphs_possible(self):
        cfg = self.config
        if not (getattr(cfg, "step_graphs", False) and not cfg.use_preemption and getattr(self, "_step_graphs_ok", True)):
            return False
        if cfg.host_buffers:
            # host frames: a step is TWO graphs around the point where the env worker needs the actions on the host
            # (``_capture_step_split``); plain SMT policy on the synthetic env only
            return (cfg.policy_type == "smt" and getattr(self.envs, "fused_step", False)
                    and getattr(self.envs, "host_buffers", False) and hasattr(self.envs, "stage_frames"))
        if cfg.policy_type == "interactive":
            # pi_g / pi_l are frozen, pi_q never trains its encoders (policy.py:1034-1036): no packed weight changes
            # between rollouts; env / bookkeeping / memory / CLIP-cache state lives in persistent device buffers
            return hasattr(self.envs, "_gstate")
        # (trainable encoders: their packed tensor-core weights are re-packed in place, and the re-pack that follows an
        # optimizer step is itself captured in the first step's graph — see nn._packed_weight)
        return cfg.policy_type == "smt" and getattr(self.envs, "fused_step", False)

    def _capture_step(self, s):
        # (capture_begin / capture_end directly: the torch.cuda.graph context synchronises the device, runs the garbage
        # collector and empties the allocator cache on every entry — 150 times per rollout)
        net = self.actor_critic.net
        g = torch.cuda.CUDAGraph()
        n0 = int(_lib.lib().avl_launch_count())
        kw = {"pool": self._graph_pool} if self._graph_pool is not None else {}
        g.capture_begin(capture_error_mode="relaxed", **kw)
        try:
            self._collect_rollout_step(self.rollouts)
            K.sync_pending()      # deferred belief update: joined inside the captured step
            if hasattr(net, "join_prefetch"):
                net.join_prefetch()   # encoder prefetch of slot s + 1: complete when the step's graph completes
        finally:
            g.capture_end()
        if self._graph_pool is None:
            self._graph_pool = g.pool()
        n1 = int(_lib.lib().avl_launch_count())
        g._avl_launches = n1 - n0
        _lib.lib().avl_launch_count_add(-(n1 - n0))  # counted at capture, but nothing ran yet
        return g

    # Host frames (``host_buffers``: the e2e path): the env worker needs the step's actions on the host and hands back
    # frames that the host has to stage, so a step cannot be one graph.  It is two: A = policy act + D2H of the actions;
    # [host: wait for A, env workers step, batch_obs stages the frames into pinned memory and copies them into FIXED
    # device buffers]; B = device side of the env step (episode bookkeeping, audio rendering), belief networks, encoder
    # prefetch of the next observation, storage insert.  The split is made from inside ``envs.step`` through the env's
    # ``graph_split`` hook, so ``_collect_rollout_step`` is captured unchanged.
    def _capture_step_split(self, s):
        net = self.actor_critic.net
        env = self.envs
        ga, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        n0 = int(_lib.lib().avl_launch_count())
        kw = lambda: ({"pool": self._graph_pool} if self._graph_pool is not None else {})  # noqa: E731
        state = {"in_b": False}

        def split():
            K.sync_pending()
            ga.capture_end()
            if self._graph_pool is None:
                self._graph_pool = ga.pool()
            ga.replay()                                  # the step really runs while it is being captured
            torch.cuda.current_stream().synchronize()    # actions are on the host
            env.stage_frames()
            gb.capture_begin(capture_error_mode="relaxed", **kw())
            state["in_b"] = True

        env.graph_split = split
        env.static_frames = True
        ga.capture_begin(capture_error_mode="relaxed", **kw())
        try:
            self._collect_rollout_step(self.rollouts)
            K.sync_pending()
            if hasattr(net, "join_prefetch"):
                net.join_prefetch()
        finally:
            env.graph_split = None
            (gb if state["in_b"] else ga).capture_end()
        n1 = int(_lib.lib().avl_launch_count())
        gb._avl_launches = n1 - n0   # kernels of A + B; both run exactly once here (A above, B below): the count stands
        ga._avl_launches = 0
        gb.replay()
        return (ga, gb)

    def _replay_step_split(self, pair):
        ga, gb = pair
        env = self.envs
        ga.replay()
        torch.cuda.current_stream().synchronize()
        env._t += 1
        env.stage_frames()
        gb.replay()
        _lib.lib().avl_launch_count_add(gb._avl_launches)
        r = self.rollouts
        r.step += 1
        r.em.advance_host_index()

    def _replay_step(self, g):
        if isinstance(g, tuple):
            return self._replay_step_split(g)
        g.replay()
        _lib.lib().avl_launch_count_add(g._avl_launches)
        # the host-side counters the eager step advances
        r = self.rollouts
        r.step += 1
        r.em.advance_host_index()
        if self.config.policy_type == "interactive":
            for em in (r.em_option, r.em_vln, r.em_vln_dialog):
                em.advance_host_index()
        self.envs._t += 1

    def collect_rollout(self):
        """The 150-step rollout loop (ddppo_trainer.py:894-959)."""
        cfg = self.config
        if self._step_graphs_possible():
            return self._collect_rollout_graphed()
        store = distrib.distributed_c10d._get_default_store() if (self.world_size > 1 and cfg.use_preemption) else None
        for step in range(cfg.num_steps):
            with _lib.nvtx_range("rollout_step"):
                self._collect_rollout_step(self.rollouts)
            if store is not None and step >= cfg.num_steps * self.SHORT_ROLLOUT_THRESHOLD:
                if int(store.add("num_done", 0)) > cfg.sync_frac * self.world_size:
                    break
        if store is not None:
            store.add("num_done", 1)
        K.sync_pending()  # the last step's deferred belief update joins the main stream here
        return self.rollouts.step * self.envs.num_envs

    def _collect_rollout_graphed(self):
        cfg = self.config
        main = torch.cuda.current_stream()
        if getattr(self, "_rollout_stream", None) is None:
            self._rollout_stream = torch.cuda.Stream()
            self._graph_pool = None
            self._step_graphs = None
            self._rollouts_seen = 0
        rs = self._rollout_stream
        rs.wait_stream(main)
        net = self.actor_critic.net
        try:
            with torch.cuda.stream(rs):
                if self._step_graphs is not None:
                    for g in self._step_graphs:
                        self._replay_step(g)
                elif self._rollouts_seen == 0:
                    # warm-up rollout on the rollout stream: allocator, lazy library state, per-stream resources
                    for _ in range(cfg.num_steps):
                        self._collect_rollout_step(self.rollouts)
                    K.sync_pending()
                else:
                    net.drop_prefetch()
                    torch.cuda.synchronize()
                    graphs = []
                    for s in range(cfg.num_steps):
                        if cfg.host_buffers:
                            graphs.append(self._capture_step_split(s))  # (runs the step while capturing it)
                            continue
                        g = self._capture_step(s)
                        g.replay()
                        _lib.lib().avl_launch_count_add(g._avl_launches)
                        graphs.append(g)
                    self._step_graphs = graphs
                    # the prefetch entry the last captured step left behind is owned by its graph: drop the host handle
                    net._prefetched().clear()
            self._rollouts_seen += 1
        except Exception as e:  # capture not possible in this environment: stay eager for good
            import warnings
            warnings.warn(f"avlen_b200: whole-step CUDA graphs disabled ({type(e).__name__}: {e})", RuntimeWarning)
            self._step_graphs_ok = False
            self._step_graphs = None
            torch.cuda.synchronize()
            raise
        main.wait_stream(rs)
        return self.rollouts.step * self.envs.num_envs

    def reset_preemption_counter(self):
        """ddppo_trainer.py:1005 / :1071: rank 0 zeroes ``num_done`` after every update; the barrier keeps a fast rank
        from starting (and finishing a quarter of) the next rollout against the stale count."""
        if self.world_size > 1 and self.config.use_preemption:
            if self.world_rank == 0:
                distrib.distributed_c10d._get_default_store().set("num_done", "0")
            distrib.barrier()

    def train(self):
        if self.rollouts is None:
            self.setup()
        cfg = self.config
        t0 = time.time()
        count_steps = 0
        stats = None
        for _update in range(cfg.NUM_UPDATES):
            count_steps += self.collect_rollout()
            stats = self._update_agent(cfg, self.rollouts)
            self.reset_preemption_counter()
        torch.cuda.synchronize()
        fps = count_steps * self.world_size / max(1e-9, time.time() - t0)
        return {"fps": fps, "value_loss": stats[0], "action_loss": stats[1], "dist_entropy": stats[2]}
