"""PPO of av_nav (ss_baselines/av_nav/ppo/ppo.py:17-163): same constructor, ``update(rollouts) -> (value_loss,
action_loss, dist_entropy)``, hooks and ``optimizer``.  Per minibatch: policy heads -> ONE fused loss kernel
(forward terms + gradient wrt logits / values) -> autograd backward through the GRU / CNN kernels -> one fused
global-norm clip + Adam kernel over the flat parameter buffer.  The three scalars are read back once per update."""
from __future__ import annotations

import torch
import torch.nn as nn

from ... import ops
from ...savi.ppo.ppo import flatten_parameters

EPS_PPO = 1e-5


class PPO(nn.Module):
    def __init__(self, actor_critic, clip_param, ppo_epoch, num_mini_batch, value_loss_coef, entropy_coef, lr=None,
                 eps=None, max_grad_norm=None, use_clipped_value_loss=True, use_normalized_advantage=True):
        super().__init__()
        self.actor_critic = actor_critic
        self.clip_param, self.ppo_epoch, self.num_mini_batch = clip_param, ppo_epoch, num_mini_batch
        self.value_loss_coef, self.entropy_coef = value_loss_coef, entropy_coef
        self.max_grad_norm = max_grad_norm
        self.use_clipped_value_loss = use_clipped_value_loss
        self.use_normalized_advantage = use_normalized_advantage
        self.device = next(actor_critic.parameters()).device
        self._params, self._flat_p, self._flat_g = flatten_parameters(actor_critic)
        self.optimizer = ops.FlatAdam(self._flat_p, self._flat_g, lr=lr, eps=eps, views=self._params)
        self._loss = ops.PpoLoss(self.device)

    def forward(self, *x):
        raise NotImplementedError

    def get_advantages(self, rollouts):
        return ops.advantages(rollouts.returns, rollouts.value_preds, rollouts.rewards.size(0),
                              self.use_normalized_advantage, EPS_PPO)

    def update(self, rollouts, perm_fn=None):
        advantages = self.get_advantages(rollouts)
        sums = torch.zeros(8, device=self.device)
        n_updates = 0
        for _e in range(self.ppo_epoch):
            perm = perm_fn(rollouts.rewards.size(1)) if perm_fn is not None else None
            for sample in rollouts.recurrent_generator(advantages, self.num_mini_batch, perm=perm):
                (obs_batch, hidden_batch, actions_batch, prev_actions_batch, value_preds_batch, return_batch,
                 masks_batch, old_lp_batch, adv_targ) = sample
                self._flat_g.zero_()
                logits, values = self.actor_critic.evaluate_heads(obs_batch, hidden_batch, prev_actions_batch,
                                                                  masks_batch)
                out, dlogits, dvalues, _ = self._loss(
                    logits.detach(), actions_batch, old_lp_batch, adv_targ, values.detach(), value_preds_batch,
                    return_batch, None, None, None, self.clip_param, self.value_loss_coef, self.entropy_coef, 0.0,
                    self.use_clipped_value_loss)
                self.before_backward(None)
                torch.autograd.backward([logits, values], [dlogits, dvalues])
                self.after_backward(None)
                self.before_step()
                self.optimizer.step(self.max_grad_norm)
                self.after_step()
                sums += out
                n_updates += 1
        s = (sums / max(1, n_updates)).tolist()
        return s[0], s[1], s[2]

    def before_backward(self, loss):
        pass

    def after_backward(self, loss):
        pass

    def before_step(self):
        pass  # nn.utils.clip_grad_norm_ (ppo.py:159-162) is fused into the optimizer kernel

    def after_step(self):
        pass
