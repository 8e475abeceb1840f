"""PPO (ss_baselines/savi/ppo/ppo.py:30-303; plain-SAVi loss = av_nav/ppo/ppo.py:60-151) on fused CUDA kernels.

Per minibatch: policy heads -> ONE fused loss kernel that returns every loss term and the gradient of the total loss
w.r.t. logits / value / uncertainty logits -> autograd backward from those -> one flat-buffer global-norm clip + Adam
kernel.  No ``.item()`` inside the minibatch loop; the six scalars are read back once per update.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ... import _lib
from ... import nn as K
from ... import ops

EPS_PPO = 1e-5


def flatten_parameters(module: nn.Module):
    """Re-points every trainable parameter (and its .grad) into one flat fp32 buffer so that the gradient
    all-reduce, the global-norm clip and Adam each touch a single contiguous array."""
    params = [p for p in module.parameters() if p.requires_grad]
    n = sum(p.numel() for p in params)
    dev = params[0].device
    flat_p = torch.empty(n, device=dev, dtype=torch.float32)
    flat_g = torch.zeros(n, device=dev, dtype=torch.float32)
    off = 0
    for p in params:
        k = p.numel()
        flat_p[off:off + k].copy_(p.data.reshape(-1))
        p.data = flat_p[off:off + k].view_as(p.data)
        p.grad = flat_g[off:off + k].view_as(p.data)
        off += k
    return params, flat_p, flat_g


class PPO(nn.Module):
    def __init__(self, actor_critic, clip_param, ppo_epoch, num_mini_batch, value_loss_coef, entropy_coef, lr=None,
                 eps=None, max_grad_norm=None, use_clipped_value_loss=True, use_normalized_advantage=True,
                 unct_coef=0.5, policy_head="goal"):
        super().__init__()
        self.actor_critic = actor_critic
        self.clip_param, self.ppo_epoch, self.num_mini_batch = clip_param, ppo_epoch, num_mini_batch
        self.value_loss_coef, self.entropy_coef, self.unct_coef = value_loss_coef, entropy_coef, unct_coef
        self.max_grad_norm = max_grad_norm
        self.use_clipped_value_loss = use_clipped_value_loss
        self.use_normalized_advantage = use_normalized_advantage
        self.policy_head = policy_head  # "goal": SAVi evaluate_actions ; "option": AVLEN evaluate_actions_option
        self.device = next(actor_critic.parameters()).device
        self._params, self._flat_p, self._flat_g = flatten_parameters(actor_critic)
        # pi_q builds its feature row under no_grad (policy.py:1034-1036): its encoders never see a gradient, so Adam
        # leaves them bit-identical (zero moments) and nothing derived from them has to be refreshed after a step
        net = getattr(actor_critic, "net", None)
        self._encoders_never_trained = policy_head == "option" and net is not None
        untouched = set()
        if self._encoders_never_trained:
            for enc in (net.goal_encoder, net.visual_encoder, net.action_encoder):
                untouched.update(id(p) for p in enc.parameters())
        views = [p for p in self._params if id(p) not in untouched]
        self.optimizer = ops.FlatAdam(self._flat_p, self._flat_g, lr=lr, eps=eps, views=views)
        # dialog pretraining (ppo.py:63, :70-76): its own Adam moments, lr 1e-5, class weights 'balanced'
        self.dialog_optimizer = ops.FlatAdam(self._flat_p, self._flat_g, lr=0.00001, eps=eps, views=self._params)
        self.dialog_class_weight = torch.tensor([0, .33, .33, .33], device=self.device)
        self._loss = ops.PpoLoss(self.device)
        self.world_size = 1

    def forward(self, *x):
        raise NotImplementedError

    def get_advantages(self, rollouts):
        return ops.advantages(rollouts.returns, rollouts.value_preds, rollouts.step, self.use_normalized_advantage,
                              EPS_PPO)

    def _reduce_gradients(self):
        """DD-PPO hook: all-reduce of the flat gradient (overridden in ddppo.DDPPO)."""
        return 1.0

    def update(self, rollouts, perm_fn=None):
        advantages = self.get_advantages(rollouts)
        sums = torch.zeros(8, device=self.device)
        n_updates = 0
        option = self.policy_head == "option"
        for sample in self._minibatches_with_encoder_prefetch(rollouts, advantages, perm_fn, option):
            (obs_batch, hidden_batch, actions_batch, actions_option_batch, prev_actions_batch, value_preds_batch,
             return_batch, masks_batch, old_lp_batch, adv_targ, rl_masks_batch, unct_gt_batch, em_goal, em_option,
             _em_vln, _em_dialog, em_masks, _em_vln_masks, _all_dialog, query_state_batch, last_query_info,
             _agent_step) = sample
            _nvtx = _lib.nvtx_range("ppo_minibatch")
            _nvtx.__enter__()
            self._flat_g.zero_()
            if option:
                logits, values, unct = self.actor_critic.evaluate_heads(
                    "option", obs_batch, hidden_batch, prev_actions_batch, masks_batch, em_option, em_masks,
                    query_state_batch, last_query_info)
                acts, rl_mask, ugt = actions_option_batch, rl_masks_batch.float(), unct_gt_batch
            else:
                logits, values, unct = self.actor_critic.evaluate_heads(
                    "goal", obs_batch, hidden_batch, prev_actions_batch, masks_batch, em_goal, em_masks)
                acts, rl_mask, ugt = actions_batch, None, None
            out, dlogits, dvalues, dunct = self._loss(
                logits.detach(), acts, old_lp_batch, adv_targ, values.detach(), value_preds_batch, return_batch,
                rl_mask, None if unct is None else unct.detach(), ugt, self.clip_param, self.value_loss_coef,
                self.entropy_coef, self.unct_coef, self.use_clipped_value_loss)
            self.before_backward(None)
            heads, grads = [logits, values], [dlogits, dvalues]
            if unct is not None:
                heads.append(unct)
                grads.append(dunct)
            torch.autograd.backward(heads, grads)
            self.after_backward(None)
            scale = self._reduce_gradients()
            self.before_step()
            self.optimizer.step(self.max_grad_norm, grad_scale=scale)
            self.after_step()
            sums += out
            n_updates += 1
            _nvtx.__exit__()
        s = (sums / max(1, n_updates)).tolist()  # the only host synchronisation of the update
        K.check_f16_overflow()  # (device already idle) fp16 activation storage guard, nn.check_f16_overflow
        value_loss, action_loss, entropy, unct_loss = s[0], s[1], s[2], s[3]
        # the reference returns the *sums* of the two debug means (ppo.py:279-280, :289)
        return value_loss, action_loss, entropy, s[5] * n_updates, s[6] * n_updates, unct_loss

    prefetch_encoders = True

    def _minibatches_with_encoder_prefetch(self, rollouts, advantages, perm_fn, option):
        """All ``ppo_epoch x num_mini_batch`` minibatches in the reference's order (ppo.py:163-166), one ahead: while
        minibatch k runs its scene-memory transformer forward / backward and the optimizer step on the current stream,
        the FROZEN encoders of minibatch k+1 (observation-only feature columns: no gradient, independent of the step)
        are already enqueued on a side stream — HBM-bound convolutions / GroupNorms next to FMA- and tensor-bound
        transformer kernels.  Trainable encoders are never prefetched (their weights change with the step)."""
        def minibatches():
            for _e in range(self.ppo_epoch):
                perm = perm_fn(rollouts.rewards.size(1)) if perm_fn is not None else None
                yield from rollouts.recurrent_generator(advantages, self.num_mini_batch, perm=perm)

        net = getattr(self.actor_critic, "net", None)
        can = (self.prefetch_encoders and net is not None and hasattr(net, "prefetch_observation_features")
               and advantages.is_cuda
               and (self._encoders_never_trained or
                    not any(p.requires_grad for enc in (net.goal_encoder, net.visual_encoder, net.action_encoder)
                            for p in enc.parameters())))
        if not can:
            yield from minibatches()
            return
        stream = self.__dict__.get("_encoder_stream")
        if stream is None:
            stream = self.__dict__["_encoder_stream"] = torch.cuda.Stream()
        extra = getattr(net, "_query_count_emb_size", 0) if option else 0
        it = minibatches()
        cur = next(it, None)
        if cur is not None:  # the first minibatch goes through the same stream: one encoder run at a time, in order
            net.prefetch_observation_features(cur[0], net.observation_key(cur[0]), stream, extra_cols=extra)
        while cur is not None:
            nxt = next(it, None)  # its gathers are enqueued on the current stream before minibatch k's kernels
            if nxt is not None:
                net.prefetch_observation_features(nxt[0], net.observation_key(nxt[0]), stream, extra_cols=extra)
            yield cur
            cur = nxt
        net.drop_prefetch()

    def update_dialog(self, rollouts):
        """ppo.py:99-154: one full-batch ``evaluate_actions_dialog`` over the NUM_DIALOG_STEPS x N rows, weighted
        cross-entropy against the oracle actions on the rows with ``o_masks != 0`` (selected on the device), Adam
        (lr 1e-5, no gradient clipping).  Returns the loss as a device scalar (the reference returns a tensor)."""
        (obs_batch, hidden_batch, actions_batch, prev_actions_batch, _, _, masks_batch, _, _, external_memory,
         external_memory_dialog, _em_masks, external_memory_vln_masks, all_dialog_batch, agent_step_batch, _num_steps,
         _num_envs) = rollouts.dialog_batching()
        self._flat_g.zero_()
        (_, _, _, _, _, _, logits) = self.actor_critic.evaluate_actions_dialog(
            obs_batch, hidden_batch, prev_actions_batch, masks_batch, actions_batch, external_memory,
            external_memory_dialog, external_memory_vln_masks, all_dialog_batch, agent_step_batch.detach())
        T = rollouts.step
        dialog_loss, _stats = ops.masked_weighted_ce(logits, rollouts.o_actions[:T].reshape(-1),
                                                     rollouts.o_masks[:T].reshape(-1), self.dialog_class_weight)
        self.before_backward(dialog_loss)
        dialog_loss.backward()
        scale = self._reduce_gradients()
        self.dialog_optimizer.step(None, grad_scale=scale)
        return dialog_loss.detach()

    def before_backward(self, loss):
        pass

    def after_backward(self, loss):
        pass

    def before_step(self):
        pass  # clip_grad_norm_ is fused into the optimizer kernel

    def after_step(self):
        pass

    def state_dict(self, *a, **k):
        return super().state_dict(*a, **k)  # keys prefixed 'actor_critic.' (checkpoint format, ppo_trainer.py:196)
