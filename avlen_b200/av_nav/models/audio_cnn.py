"""AudioCNN of av_nav (ss_baselines/av_nav/models/audio_cnn.py:15-89) — the same network as the SAVi one
(savi/models/audio_cnn.py) without the distractor label planes."""
from ...savi.models.audio_cnn import AudioCNN  # noqa: F401
