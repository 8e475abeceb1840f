"""Diagnostic: pi_g with use_category_input=True (BASELINE configs 4 / 5) on the GPU against the reference golden vector."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from avlen_b200 import nn as K
from avlen_b200.common import spaces
from avlen_b200.savi.ppo.policy import AudioNavSMTPolicy
from oracle import models_torch as OM

K.set_tensor_cores(False)
g = dict(np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                              "smt_policy_distractor.npz")).items())
d = lambda a: torch.from_numpy(np.asarray(a)).cuda()  # noqa: E731
p = AudioNavSMTPolicy(spaces.savi_observation_space(), spaces.Discrete(4), hidden_size=256, nhead=8, num_encoder_layers=1,
                      num_decoder_layers=1, dropout=0.0, activation="relu", pretraining=False, use_category_input=True)
p.load_state_dict(OM.seeded_state_dict(OM.AudioNavSMTPolicy(pretraining=False, use_category_input=True), int(g["seed"])))
p = p.cuda()
p.net.freeze_encoders()
p.net.set_eval_encoders()
o = {k[4:]: d(v) for k, v in g.items() if k.startswith("obs_")}
o["rgb"] = o["rgb"].float()
n = g["em"].shape[1]
h = torch.zeros(1, n, 512, device="cuda")
rel = lambda a, b: float((a.detach().float().cpu() - torch.from_numpy(np.asarray(b)).float()).abs().max() / max(1e-12, float(np.abs(b).max())))  # noqa: E731
with torch.no_grad():
    v, a, lp, _, x, pr = p.act(o, h, d(g["prev_actions"]), d(g["masks"]), d(g["em"]), d(g["em_masks"]), deterministic=True)
print("act", torch.equal(a.cpu(), torch.from_numpy(g["act_action"])), rel(v, g["act_value"]), rel(lp, g["act_log_probs"]),
      rel(pr, g["act_probs"]), rel(x, g["act_em_feats"]))
v, lp, ent, _, x = p.evaluate_actions(o, h, d(g["prev_actions"]), d(g["masks"]), d(g["action"]), d(g["em"]), d(g["em_masks"]))
print("eval", rel(v, g["eval_value"]), rel(lp, g["eval_log_probs"]), rel(x, g["eval_em_feats"]),
      abs(float(ent.detach()) - float(g["eval_entropy"])))
