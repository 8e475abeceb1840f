"""Minimal stand-ins for the gym spaces the reference constructors read (``.spaces[uuid].shape``, ``.n``);
gym itself is not a dependency of the hot path."""
from __future__ import annotations

import numpy as np


class Box:
    def __init__(self, low=None, high=None, shape=None, dtype=np.float32):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype


class Discrete:
    def __init__(self, n):
        self.n = int(n)


class Dict:
    def __init__(self, spaces):
        self.spaces = dict(spaces)


def savi_observation_space(sr: int = 16000, with_audiogoal: bool = False):
    """Observation space of the SAVi / AVLEN task (configs/semantic_audionav/savi/mp3d/semantic_audiogoal.yaml:6-25)."""
    d = {
        "rgb": Box(0, 255, (128, 128, 3), np.uint8),
        "depth": Box(0, 1, (128, 128, 1), np.float32),
        "spectrogram": Box(-3.4e38, 3.4e38, (65, (1 + sr // 160 + 3) // 4, 2), np.float32),
        "pose": Box(-3.4e38, 3.4e38, (4,), np.float32),
        "category": Box(0, 1, (21,), np.float32),
        "category_belief": Box(0, 1, (21,), np.float32),
        "location_belief": Box(-3.4e38, 3.4e38, (2,), np.float32),
    }
    if with_audiogoal:
        d["audiogoal"] = Box(-3.4e38, 3.4e38, (2, sr), np.float32)
    return Dict(d)
