"""The two hot loops of the SAVi trainer (ss_baselines/savi/ppo/ppo_trainer.py:323-897 ``_collect_rollout_step``,
:1045-1093 ``_update_agent``) for the plain SMT policy (``policy_type: "smt"``), device-resident.

The reference's per-env Python loops, ``.item()`` / ``.cpu()`` round trips and pipe traffic to env workers are not
reproduced; the call sequence on the policy / belief predictor / storage objects is the reference's.
"""
from __future__ import annotations

import torch

from ... import _lib
from ... import nn as K


class PPOTrainer:
    def __init__(self, config=None):
        self.config = config
        self.actor_critic = None
        self.agent = None
        self.belief_predictor = None
        self.envs = None
        self.device = None

    @torch.no_grad()
    def _collect_rollout_step(self, rollouts, current_episode_reward=None, running_episode_stats=None, uniforms=None):
        """One environment step for all envs (ppo_trainer.py:323-897, smt policy path :606-623, :714-894)."""
        if getattr(self.config, "policy_type", "smt") == "interactive":
            return self._collect_rollout_step_interactive(rollouts)
        s = rollouts.step
        step_observation = {k: v[s] for k, v in rollouts.observations.items()}
        values, actions, actions_log_probs, recurrent_hidden_states, external_memory_features, _probs = \
            self.actor_critic.act(step_observation, rollouts.recurrent_hidden_states[s], rollouts.prev_actions[s],
                                  rollouts.masks[s], rollouts.external_memory_goal[:, s],
                                  rollouts.external_memory_masks[s], uniforms=uniforms)
        observations, rewards, dones = self.envs.step(actions)
        masks = getattr(self.envs, "last_masks", None)  # (N, 1) not-done floats when the env already produced them
        if masks is None:
            masks = (~dones).float().unsqueeze(1)
        if current_episode_reward is not None:
            current_episode_reward += rewards
            if running_episode_stats is not None:
                running_episode_stats["reward"] += (1 - masks) * current_episode_reward
                running_episode_stats["count"] += 1 - masks
            current_episode_reward *= masks
        box = [observations]

        def belief():  # :890-894: belief for the NEXT observation
            if self.belief_predictor is not None:
                if getattr(self.config, "overlap_belief", False):
                    box[0] = self._belief_update_deferred(rollouts, observations, dones)
                else:
                    self.belief_predictor.update(observations, dones)

        if getattr(self.config, "prefetch_encoders", False) and s + 1 < rollouts.masks.shape[0]:
            # the visual / audio encoders of observation s+1 only need what the environment just returned; the belief
            # networks are enqueued right after the visual encoders (host order = start order of the chains)
            net = self.actor_critic.net
            vs = getattr(self, "_encoder_stream", None)
            if vs is None:
                vs = self._encoder_stream = torch.cuda.Stream()
            slot = {k: rollouts.observations[k][s + 1] for k in net._PREFETCH_KEYS if k in rollouts.observations}
            net.prefetch_observation_features(observations, net.observation_key(slot), vs,
                                              visual_event=getattr(self.envs, "visual_ready_event", None),
                                              between=belief)
        else:
            belief()
        observations = box[0]
        rollouts.insert(observations, recurrent_hidden_states, actions, None, actions_log_probs, values, rewards, masks,
                        masks, external_memory_features, None, None, None, None, None, None, None, None, None, None,
                        None, None)
        return self.envs.num_envs

    @torch.no_grad()
    def _collect_rollout_step_interactive(self, rollouts):
        """The AVLEN interactive step (ppo_trainer.py:323-897, ``DIALOG_TRAINING`` False): pi_q decides whether to query
        (``act_option``), the query bookkeeping runs on the device (``QueryBookkeeper``: :394-416, :449-460, :487-588),
        pi_g (``act``) and pi_l (``act_dialog`` with the CLIP-embedded instruction) both act, the arbitration picks
        the executed action and builds ``o_mask`` / ``ucnt_gt`` (:639-694), the env receives the query flags that shape
        its reward (:706-710), and the four memories are written (:864-888).  No ``.item()`` / ``.cpu()`` anywhere.
        Deviation: pi_q sees THIS step's query-count rows; the reference reads
        ``rollouts.query_state[step]`` before writing it (:440-442 vs :588-590), i.e. the rows of the previous rollout."""
        s = rollouts.step
        envs, book = self.envs, self.query_book
        obs = {k: v[s] for k, v in rollouts.observations.items()}
        h = rollouts.recurrent_hidden_states[s]
        prev = rollouts.prev_actions[s]
        # pi_g needs nothing from pi_q or pi_l: its whole act() (two ResNet-18s, audio CNN, scene-memory transformer,
        # heads) runs on a side stream next to pi_q's and pi_l's, joined where the arbitration reads its outputs
        main = torch.cuda.current_stream()
        gs = getattr(self, "_goal_stream", None)
        if gs is None:
            gs = self._goal_stream = torch.cuda.Stream()
        gs.wait_stream(main)
        with torch.cuda.stream(gs):
            _vg, ag, _lpg, _, xg, pg = self.actor_critic_goal.act(
                obs, h, prev, rollouts.masks[s], rollouts.external_memory_goal[:, s], rollouts.external_memory_masks[s])
            goal_done = torch.cuda.Event()
            goal_done.record(gs)
        # ... and so does the dialog-independent part of pi_l (features + scene-memory transformer)
        ls = getattr(self, "_vln_stream", None)
        if ls is None:
            ls = self._vln_stream = torch.cuda.Stream()
        ls.wait_stream(main)
        with torch.cuda.stream(ls):
            scene = self.actor_critic_vln.net.encode_scene(obs, prev, rollouts.external_memory_vln[:, s],
                                                           rollouts.external_memory_vln_masks[s])
            scene_done = torch.cuda.Event()
            scene_done.record(ls)
        qs, lq = book.pre(envs.is_new_episode())
        vq, unct, aq, lpq, h_out, xq, _pq = self.actor_critic_option.act_option(
            obs, h, prev, rollouts.masks[s], rollouts.external_memory_option[:, s], rollouts.external_memory_masks[s], qs, lq)
        is_q, qnum, cons, rl_mask, dialog, agent_step = book.after_option(aq, envs.target_distance(), envs.pending_dialog())
        main.wait_event(scene_done)
        for t_ in scene:
            t_.record_stream(main)
        _vl, al, _lpl, _, xl, xd, pl = self.actor_critic_vln.act_dialog(
            obs, h, prev, rollouts.masks_vln[s], rollouts.external_memory_vln[:, s],
            rollouts.external_memory_vln_dialog[:, s], rollouts.external_memory_vln_masks[s], dialog, agent_step,
            scene=scene)
        main.wait_event(goal_done)
        for t_ in (ag, xg, pg):
            t_.record_stream(main)
        oracle = envs.compute_oracle_actions()
        o_action = oracle.float()
        actions, o_mask, ucnt_gt, masks_vln = book.arbitrate(ag, al, pg, oracle)
        envs.set_is_queried(is_q)
        envs.set_query_num(qnum)
        envs.set_constraint_reward(cons)
        observations, rewards, dones = envs.step(actions)
        masks = getattr(envs, "last_masks", None)
        if masks is None:
            masks = (~dones).float().unsqueeze(1)
        if self.belief_predictor is not None:
            self.belief_predictor.update(observations, dones)
        rollouts.insert(observations, h_out, actions, aq, lpq, vq, rewards, masks, masks_vln, xg, xq, xl, xd, dialog,
                        o_action, o_mask, rl_mask, ucnt_gt, pl, qs, lq, agent_step)
        return envs.num_envs

    _BELIEF_KEYS = ("location_belief", "category_belief")

    def _belief_update_deferred(self, rollouts, observations, dones):
        """The two belief networks only feed the NEXT step's scene-memory transformer, and the next step's visual
        encoders do not depend on them: enqueue the belief update (networks, belief filter, the copy of the two belief
        vectors into storage slot step+1) on a side stream and let the consumer (``AudioNavSMTNet._belief``, storage
        readers) wait for it with an event — four small-grid ResNet-18 chains share the GPU instead of two.  Same
        kernels on the same data: results are identical to the in-order path.  Returns the observations the main
        stream still has to insert."""
        main = torch.cuda.current_stream()
        side = getattr(self, "_belief_stream", None)
        if side is None:
            side = self._belief_stream = torch.cuda.Stream()
        side.wait_stream(main)
        for v in observations.values():
            if torch.is_tensor(v) and v.is_cuda:
                v.record_stream(side)
        if torch.is_tensor(dones) and dones.is_cuda:
            dones.record_stream(side)
        s = rollouts.step
        with torch.cuda.stream(side):
            self.belief_predictor.update(observations, dones)
            for k in self._BELIEF_KEYS:
                if k in observations and k in rollouts.observations:
                    rollouts.observations[k][s + 1].copy_(observations[k])
        K.defer_join(side)
        return {k: v for k, v in observations.items() if k not in self._BELIEF_KEYS}

    def _update_agent(self, ppo_cfg, rollouts):
        """ppo_trainer.py:1045-1093: bootstrap value, GAE, PPO.update, after_update."""
        with torch.no_grad():
            s = rollouts.step
            last_observation = {k: v[s] for k, v in rollouts.observations.items()}
            if getattr(ppo_cfg, "policy_type", "smt") == "interactive":  # ppo_trainer.py:1058-1068
                next_value = self.actor_critic.get_value_option(
                    last_observation, rollouts.recurrent_hidden_states[s], rollouts.prev_actions[s], rollouts.masks[s],
                    rollouts.external_memory_option[:, s], rollouts.external_memory_masks[s],
                    rollouts.query_state[s - 1], rollouts.last_query_info[s - 1])
            else:
                next_value = self.actor_critic.get_value(last_observation, rollouts.recurrent_hidden_states[s],
                                                         rollouts.prev_actions[s], rollouts.masks[s],
                                                         rollouts.external_memory_goal[:, s],
                                                         rollouts.external_memory_masks[s])
        rollouts.compute_returns(next_value, ppo_cfg.use_gae, ppo_cfg.gamma, ppo_cfg.tau)
        with _lib.nvtx_range("ppo_update"):
            value_loss, action_loss, dist_entropy, _vd, _rd, _ul = self.agent.update(rollouts)
        rollouts.after_update()
        return value_loss, action_loss, dist_entropy
