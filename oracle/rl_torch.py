"""CPU oracle for the rollout-storage / PPO rows (G, I, M scalar part, N, O, Q) — TEST INFRASTRUCTURE.

Each function restates the cited reference code in plain PyTorch/NumPy on CPU;
``tests/test_oracle_vs_reference.py`` pins them against the unmodified reference
modules (loaded through ``oracle/ref_shim.py``) in the authoring container.
"""
from __future__ import annotations

import numpy as np
import torch


# ---- row N: ss_baselines/savi/models/rollout_storage.py:394-412 ---------------------------------
def compute_returns(rewards, value_preds, masks, next_value, steps, use_gae, gamma, tau):
    """rewards (T,N,1); value_preds/masks (T+1,N,1). Mutates value_preds like the reference; returns `returns`."""
    returns = torch.zeros_like(value_preds)
    if use_gae:
        value_preds[steps] = next_value
        gae = 0
        for step in reversed(range(steps)):
            delta = rewards[step] + gamma * value_preds[step + 1] * masks[step + 1] - value_preds[step]
            gae = delta + gamma * tau * masks[step + 1] * gae
            returns[step] = gae + value_preds[step]
    else:
        returns[steps] = next_value
        for step in reversed(range(steps)):
            returns[step] = returns[step + 1] * gamma * masks[step + 1] + rewards[step]
    return returns


# ---- row O: ss_baselines/savi/ppo/ppo.py:90-95 ---------------------------------------------------
def get_advantages(returns, value_preds, normalize, eps=1e-5):
    adv = returns[:-1] - value_preds[:-1]
    if not normalize:
        return adv
    return (adv - adv.mean()) / (adv.std() + eps)


# ---- row I: ss_baselines/common/utils.py:44-72 ----------------------------------------------------
def categorical_act(logits, uniforms=None):
    """mode() when ``uniforms`` is None, else inverse-CDF sampling on the supplied uniforms
    (``torch.multinomial``'s Philox stream cannot be reproduced; SURVEY.md §7 hard parts)."""
    dist = torch.distributions.Categorical(logits=logits)
    probs = dist.probs
    if uniforms is None:
        action = probs.argmax(dim=-1, keepdim=True)  # CustomFixedCategorical.mode
    else:
        cdf = torch.cumsum(probs, dim=-1)
        action = (cdf <= uniforms[:, None]).sum(-1, keepdim=True).clamp(max=logits.shape[-1] - 1)
    log_probs = dist.log_prob(action.squeeze(-1)).view(action.size(0), -1).sum(-1).unsqueeze(-1)
    return action, log_probs, probs


def categorical_eval(logits, action):
    dist = torch.distributions.Categorical(logits=logits)
    log_probs = dist.log_prob(action.squeeze(-1)).view(action.size(0), -1).sum(-1).unsqueeze(-1)
    return log_probs, dist.entropy(), dist.probs


# ---- row Q: ss_baselines/savi/ppo/ppo.py:219-262, av_nav/ppo/ppo.py:93-131 -------------------------
def ppo_loss(logits, actions, old_lp, adv, values, value_preds, returns, rl_mask, unct, unct_gt, clip,
             value_coef, ent_coef, unct_coef, use_clipped_value=True):
    """Returns dict of losses and autograd gradients of the total loss wrt logits / values / unct."""
    logits = logits.clone().requires_grad_(True)
    values = values.clone().requires_grad_(True)
    unct_r = unct.clone().requires_grad_(True) if unct is not None else None
    dist = torch.distributions.Categorical(logits=logits)
    action_log_probs = dist.log_prob(actions.squeeze(-1)).unsqueeze(-1)
    dist_entropy = dist.entropy().mean()
    ratio = torch.exp(action_log_probs - old_lp)
    if rl_mask is not None:
        surr1 = ratio * adv * rl_mask.unsqueeze(1)
        surr2 = torch.clamp(ratio, 1.0 - clip, 1.0 + clip) * adv * rl_mask.unsqueeze(1)
        action_loss = -torch.min(surr1, surr2).sum() / torch.sum(rl_mask)
    else:
        surr1 = ratio * adv
        surr2 = torch.clamp(ratio, 1.0 - clip, 1.0 + clip) * adv
        action_loss = -torch.min(surr1, surr2).mean()
    if use_clipped_value:
        value_pred_clipped = value_preds + (values - value_preds).clamp(-clip, clip)
        value_losses = (values - returns).pow(2)
        value_losses_clipped = (value_pred_clipped - returns).pow(2)
        value_loss = 0.5 * torch.max(value_losses, value_losses_clipped).mean()
    else:
        value_loss = 0.5 * (returns - values).pow(2).mean()
    total = value_loss * value_coef + action_loss - dist_entropy * ent_coef
    unct_loss = torch.zeros(())
    if unct_r is not None:
        unct_loss = torch.nn.functional.cross_entropy(unct_r, unct_gt.long())
        total = total + unct_coef * unct_loss
    total.backward()
    return dict(value_loss=value_loss.item(), action_loss=action_loss.item(), entropy=dist_entropy.item(),
                unct_loss=float(unct_loss.detach()), total=total.item(), values_mean=values.mean().item(),
                returns_mean=returns.mean().item(), dlogits=logits.grad, dvalues=values.grad,
                dunct=None if unct_r is None else unct_r.grad)


# ---- row G: ss_baselines/savi/models/rollout_storage.py:907-941 ------------------------------------
class ExternalMemory:
    """Restatement WITH the reference's ``num_copies`` layout (total, copies, N, dim)."""

    def __init__(self, num_envs, total_size, capacity, dim, num_copies=1):
        self.num_envs, self.total_size, self.capacity, self.dim = num_envs, total_size, capacity, dim
        self.masks = torch.zeros(num_envs, total_size)
        self.memory = torch.zeros(total_size, num_copies, num_envs, dim)
        self.idx = 0

    def insert(self, em_features, not_done_masks):
        self.memory[self.idx].copy_(em_features.unsqueeze(0))
        capacity_overflow_flag = self.masks.sum(1) == self.capacity
        assert not torch.any(self.masks.sum(1) > self.capacity)
        self.masks[capacity_overflow_flag, self.idx - self.capacity] = 0.0
        self.masks[:, self.idx] = 1.0
        self.masks *= not_done_masks
        self.idx = (self.idx + 1) % self.total_size


# ---- row M (scalar part): ss_baselines/savi/models/belief_predictor.py:139-230 ----------------------
def base_to_odom(pointgoal_base, pose):
    angle = -pose[2]
    d = np.linalg.norm(pointgoal_base)
    theta = np.arctan2(pointgoal_base[1], pointgoal_base[0])
    return np.array([pose[0] + d * np.cos(theta + angle), pose[1] + d * np.sin(theta + angle)])


def odom_to_base(pointgoal_odom, pose):
    angle = -pose[2]
    delta = pointgoal_odom - pose[:2]
    delta_theta = np.arctan2(delta[1], delta[0]) - angle
    d = np.linalg.norm(delta)
    return np.array([d * np.cos(delta_theta), d * np.sin(delta_theta)])


class BeliefState:
    """Per-env EMA state of BeliefPredictor.update (location + label)."""

    def __init__(self, n, weighting_factor=0.5, current_pred_only=False):
        self.last_pointgoal = [None] * n
        self.last_label = [None] * n
        self.w = weighting_factor
        self.current_pred_only = current_pred_only

    def update(self, spectrogram, pose, dones, pointgoals, labels):
        """numpy inputs: spectrogram (N,65,26,2), pose (N,4), dones list/None, pointgoals (N,2), labels (N,21)."""
        n = spectrogram.shape[0]
        loc = np.zeros((n, 2), np.float32)
        cat = np.zeros((n, 21), np.float32)
        for i in range(n):
            ps = pose[i]
            if dones is not None and dones[i]:
                self.last_pointgoal[i] = None
            if float(torch.from_numpy(spectrogram[i]).sum().item()) != 0:
                pointgoal_base = np.array([-pointgoals[i][1], pointgoals[i][0]])
                if self.last_pointgoal[i] is None or self.current_pred_only:
                    avg = pointgoal_base
                else:
                    avg = (1 - self.w) * pointgoal_base + self.w * odom_to_base(self.last_pointgoal[i], ps)
                self.last_pointgoal[i] = base_to_odom(avg, ps)
            else:
                if self.last_pointgoal[i] is None:
                    avg = np.array([10, 10])
                else:
                    avg = odom_to_base(self.last_pointgoal[i], ps)
            loc[i] = avg
        for i in range(n):
            label = labels[i]
            if dones is not None and dones[i]:
                self.last_label[i] = None
            if float(torch.from_numpy(spectrogram[i]).sum().item()) != 0:
                if self.last_label[i] is None or self.current_pred_only:
                    avg = label
                else:
                    avg = (1 - self.w) * label + self.w * self.last_label[i]
                self.last_label[i] = avg
            else:
                avg = np.ones(21) / 21 if self.last_label[i] is None else self.last_label[i]
            cat[i] = avg
        return loc, cat


# ---- ppo.py:62,297-300: clip_grad_norm_ + Adam ------------------------------------------------------
def clip_adam_reference(param, grads, lr, eps, max_norm, betas=(0.9, 0.999)):
    """Runs torch's own clip_grad_norm_ + Adam over a list of gradient tensors; returns params after each step."""
    p = torch.nn.Parameter(param.clone())
    opt = torch.optim.Adam([p], lr=lr, eps=eps, betas=betas)
    outs, norms = [], []
    for g in grads:
        p.grad = g.clone()
        norms.append(float(torch.nn.utils.clip_grad_norm_([p], max_norm)))
        opt.step()
        outs.append(p.detach().clone())
    return outs, norms
