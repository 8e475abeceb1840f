"""Loads the UNMODIFIED reference modules from /root/reference under an import shim.

TEST INFRASTRUCTURE ONLY, and only usable in the authoring container (the GPU
box has no /root/reference): used by ``tests/golden/make_golden.py`` to produce
golden vectors and by the ``reference`` CPU tests (skipped when the reference
tree is absent) to pin ``oracle/*`` against the real reference code.

``import ss_baselines...`` cannot be used directly because the package
``__init__``s pull in trainers -> ``habitat`` -> ``yacs`` (absent).  Instead the
needed files are executed by path with stub modules for the missing third-party
imports.  No reference source is copied anywhere.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("AVLEN_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "ss_baselines"))


def _stub(name: str, **attrs):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        m.__path__ = []  # behave like a package
        sys.modules[name] = m
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


class _Anything:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, item):
        return _Anything()


_installed = False


def install():
    """Registers stub modules and empty ``ss_baselines`` package skeletons."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError("reference tree not present at " + REF_ROOT)
    import numpy as np

    _stub("torchsummary", summary=lambda *a, **k: None)
    def _clip_load(*a, **k):
        # openai/CLIP is not installed: the reference's `clip.load("ViT-B/32")` receives the oracle's restatement of
        # its text tower (oracle/clip_torch.py; `.transformer.width` is what policy.py:762 reads)
        from . import clip_torch
        m = clip_torch.CLIPText(layers=int(os.environ.get("AVLEN_SHIM_CLIP_LAYERS", "12")))
        m.transformer.width = 512
        return m, None

    _stub("clip", load=_clip_load, tokenize=lambda *a, **k: None)
    try:
        import pynvml  # noqa: F401
        real_pynvml = sys.modules["pynvml"]
        real_pynvml.nvmlInit = lambda *a, **k: None
    except Exception:
        _stub("pynvml", nvmlInit=lambda *a, **k: None)
    _stub("pynvml.smi", nvidia_smi=_Anything)
    _stub("moviepy")
    _stub("moviepy.editor")
    _stub("moviepy.audio")
    _stub("moviepy.audio.AudioClip", CompositeAudioClip=_Anything)
    _stub("cv2")
    _stub("habitat", logger=_Anything(), Config=_Anything)
    _stub("habitat.utils")
    _stub("habitat.utils.visualizations", maps=_Anything())
    _stub("habitat.utils.visualizations.utils", images_to_video=_Anything(), draw_collision=_Anything())
    _stub("habitat.utils.visualizations.maps")
    _stub("habitat_sim")
    _stub("habitat_sim.utils")
    _stub("habitat_sim.utils.common", d3_40_colors_rgb=np.zeros((40, 3), np.uint8), quat_to_angle_axis=_Anything(),
          quat_to_coeffs=_Anything(), quat_from_angle_axis=_Anything(), quat_from_coeffs=_Anything())

    class _Box:
        def __init__(self, low=None, high=None, shape=None, dtype=None):
            self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

    class _Discrete:
        def __init__(self, n):
            self.n = n

    class _Dict:
        def __init__(self, spaces):
            self.spaces = dict(spaces)

    _stub("gym")
    _stub("gym.spaces", Box=_Box, Discrete=_Discrete, Dict=_Dict)

    def _uuid_cls(u):
        return type("S_" + u, (), {"cls_uuid": u})

    _stub("soundspaces")
    _stub("soundspaces.tasks")
    _stub("soundspaces.tasks.nav", PoseSensor=_uuid_cls("pose"), SpectrogramSensor=_uuid_cls("spectrogram"),
          LocationBelief=_uuid_cls("location_belief"), CategoryBelief=_uuid_cls("category_belief"),
          Category=_uuid_cls("category"))
    for pkg in ["ss_baselines", "ss_baselines.common", "ss_baselines.savi", "ss_baselines.savi.models",
                "ss_baselines.savi.ppo", "ss_baselines.av_nav", "ss_baselines.av_nav.models",
                "ss_baselines.av_nav.ppo", "ss_baselines.savi.ddppo", "ss_baselines.savi.ddppo.algo",
                "ss_baselines.savi.dialog", "ss_baselines.savi.dialog.ques_gen"]:
        m = _stub(pkg)
        m.__path__ = [os.path.join(REF_ROOT, *pkg.split("."))]
    _stub("ss_baselines.common.tensorboard_utils", TensorboardWriter=_Anything)
    _stub("ss_baselines.savi.dialog.ques_gen.utils", get_glove_embedding=_Anything(), load_vocab=_Anything())
    _installed = True


def load(modname: str):
    """``load('ss_baselines.savi.ppo.policy')`` -> the reference module object."""
    install()
    if modname in sys.modules and getattr(sys.modules[modname], "__file__", None):
        return sys.modules[modname]
    path = os.path.join(REF_ROOT, *modname.split(".")) + ".py"
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


def load_simulator():
    """The UNMODIFIED ``soundspaces/simulator.py`` as a module (for ``SoundSpacesSim._compute_audiogoal`` /
    ``get_current_audiogoal_observation``, simulator.py:644-721).  Its imports that are absent here (librosa,
    habitat, habitat_sim, networkx when missing, soundspaces.utils / mp3d_utils) are stubbed; the methods under test
    only use numpy, ``scipy.io.wavfile`` and ``scipy.signal.fftconvolve``, which are the real packages."""
    install()
    name = "soundspaces.simulator"
    if name in sys.modules and getattr(sys.modules[name], "__file__", None):
        return sys.modules[name]
    _stub("librosa")
    try:
        import networkx  # noqa: F401
    except Exception:
        _stub("networkx")

    class _Registry:
        def register_simulator(self, *a, **k):
            return lambda cls: cls

        def __getattr__(self, item):
            return lambda *a, **k: (lambda cls: cls)

    _stub("habitat.core")
    _stub("habitat.core.registry", registry=_Registry())
    _stub("habitat.sims")
    _stub("habitat.sims.habitat_simulator")
    _stub("habitat.sims.habitat_simulator.actions", HabitatSimActions=_Anything())
    base = type("Simulator", (), {})
    _stub("habitat.core.simulator", AgentState=_Anything, Config=_Anything, Observations=_Anything,
          SensorSuite=_Anything, ShortestPathPoint=_Anything, Simulator=base)
    hs = sys.modules["habitat_sim"]
    for attr in ("AgentState", "Simulator", "SimulatorConfiguration", "AgentConfiguration", "Configuration"):
        setattr(hs, attr, _Anything)
    _stub("soundspaces.utils", load_metadata=_Anything())
    _stub("soundspaces.mp3d_utils", HouseReader=_Anything)
    path = os.path.join(REF_ROOT, "soundspaces", "simulator.py")
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_trainer():
    """The UNMODIFIED ``ss_baselines/savi/ppo/ppo_trainer.py`` as a module (for ``PPOTrainer._collect_rollout_step``,
    :323-897, executed on a scripted stand-in for ``self`` by tests/golden/make_golden.py).  Everything it imports
    that is absent here (matplotlib, habitat task / geometry helpers, env construction, slurm helpers, tensorboard)
    is stubbed; ``batch_obs``, ``RolloutStorage`` and the policies it touches are the reference's own."""
    install()
    name = "ss_baselines.savi.ppo.ppo_trainer"
    if name in sys.modules and getattr(sys.modules[name], "__file__", None):
        return sys.modules[name]
    for mod in ("matplotlib", "matplotlib.pyplot"):
        try:
            importlib.import_module(mod)
        except Exception:
            _stub(mod)
    _stub("habitat.tasks")
    _stub("habitat.tasks.nav")
    _stub("habitat.tasks.nav.nav", IntegratedPointGoalGPSAndCompassSensor=_Anything)
    _stub("habitat.tasks.utils", cartesian_to_polar=_Anything())
    _stub("habitat.utils.geometry_utils", quaternion_from_coeff=_Anything(), quaternion_rotate_vector=_Anything())
    base = type("BaseRLTrainer", (), {"__init__": lambda self, config=None: setattr(self, "config", config)})
    _stub("ss_baselines.common.base_trainer", BaseRLTrainer=base, BaseTrainer=base)
    _stub("ss_baselines.common.env_utils", construct_envs=_Anything())
    _stub("ss_baselines.common.environments", get_env_class=_Anything())
    _stub("ss_baselines.savi.ppo.slurm_utils", EXIT=_Anything(), REQUEUE=_Anything(), load_interrupted_state=_Anything(),
          requeue_job=_Anything(), save_interrupted_state=_Anything())
    _stub("ss_baselines.savi.models.belief_predictor", BeliefPredictor=_Anything, BeliefPredictorDDP=_Anything)
    sys.modules["habitat"].Config = _Anything

    class _Reg:  # habitat.core.registry.Registry: only the decorator plumbing the reference's BaselineRegistry uses
        mapping = {}

        @classmethod
        def _register_impl(cls, _type, to_register, name, assert_type=None):
            def wrap(to_register):
                cls.mapping.setdefault(_type, {})[name or to_register.__name__] = to_register
                return to_register
            return wrap if to_register is None else wrap(to_register)

        @classmethod
        def _get_impl(cls, _type, name):
            return cls.mapping.get(_type, {}).get(name)

    core = _stub("habitat.core")
    reg = _stub("habitat.core.registry", Registry=_Reg)
    if not hasattr(reg, "registry"):
        reg.registry = _Reg()
    load("ss_baselines.common.baseline_registry")
    return load(name)


def spaces():
    install()
    return sys.modules["gym.spaces"]


def observation_space(distractor: bool = False):
    """Habitat-shaped observation space of the SAVi/AVLEN task (semantic_audiogoal.yaml:6-25)."""
    import numpy as np
    sp = spaces()
    d = {
        "rgb": sp.Box(0, 255, (128, 128, 3), np.uint8),
        "depth": sp.Box(0, 1, (128, 128, 1), np.float32),
        "spectrogram": sp.Box(-1e9, 1e9, (65, 26, 2), np.float32),
        "pose": sp.Box(-1e9, 1e9, (4,), np.float32),
        "category": sp.Box(0, 1, (21,), np.float32),
        "category_belief": sp.Box(0, 1, (21,), np.float32),
        "location_belief": sp.Box(-1e9, 1e9, (2,), np.float32),
    }
    return sp.Dict(d)
