"""Shared fixtures for the policy tests: seeded weights + synthetic Habitat-shaped observations."""
import numpy as np
import torch

from avlen_b200 import synth
from oracle import models_torch as OM


def make_obs(n, seed, step=3):
    rng = np.random.default_rng(seed)
    o = synth.make_observations(rng, n, step)
    o["spectrogram"] = np.abs(rng.standard_normal((n, 65, 26, 2))).astype(np.float32)
    o["category_belief"] = rng.random((n, 21)).astype(np.float32)
    o["location_belief"] = rng.normal(0, 3, (n, 2)).astype(np.float32)
    return {k: torch.from_numpy(v) for k, v in o.items()}


def make_memory(M, n, dim, seed, valid_frac=0.4):
    g = torch.Generator().manual_seed(seed)
    mem = torch.randn(M, n, dim, generator=g)
    mem[..., dim - 4:] = torch.cat([torch.randn(M, n, 2, generator=g) * 5, torch.rand(M, n, 1, generator=g) * 6 - 3,
                                    torch.randint(0, 60, (M, n, 1), generator=g).float()], -1)
    masks = (torch.rand(n, M, generator=g) < valid_frac).float()
    return mem, masks


def oracle_and_cuda_policies(seed=5, pretraining=False, freeze=True):
    from avlen_b200.common import spaces
    from avlen_b200.savi.ppo.policy import AudioNavSMTPolicy
    o = OM.AudioNavSMTPolicy(pretraining=pretraining)
    sd = OM.seeded_state_dict(o, seed)
    o.load_state_dict(sd)
    o.eval()
    p = AudioNavSMTPolicy(spaces.savi_observation_space(), spaces.Discrete(4), hidden_size=256, nhead=8,
                          num_encoder_layers=1, num_decoder_layers=1, dropout=0.0, activation="relu",
                          pretraining=pretraining)
    p.load_state_dict(sd)
    p = p.cuda()
    if freeze:
        p.net.freeze_encoders()
        p.net.set_eval_encoders()
        for q in list(o.net.goal_encoder.parameters()) + list(o.net.visual_encoder.parameters()) + \
                list(o.net.action_encoder.parameters()):
            q.requires_grad = False
    return o, p
