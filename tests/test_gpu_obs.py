"""Row T (batch_obs, ss_baselines/common/utils.py:129-156) on the GPU path: list of per-env numpy observation dicts ->
pinned double-buffered staging -> asynchronous H2D -> device-side cast; and the e2e environment path that uses it."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _obs_list(rng, n):
    return [{"rgb": rng.integers(0, 256, (128, 128, 3), dtype=np.uint8),
             "depth": rng.random((128, 128, 1), dtype=np.float32),
             "pose": rng.standard_normal(4).astype(np.float32)} for _ in range(n)]


def test_batch_obs_pinned_path_equals_reference_semantics():
    from avlen_b200.common.utils import batch_obs
    rng = np.random.default_rng(0)
    pinned = {}
    dev = torch.device("cuda")
    for _step in range(5):  # > 2 steps: both staging buffers of every sensor are reused
        obs = _obs_list(rng, 6)
        out = batch_obs(obs, device=dev, pinned=pinned)
        torch.cuda.synchronize()
        for k in obs[0]:
            want = torch.stack([torch.from_numpy(np.asarray(o[k])).float() for o in obs])  # utils.py:149-154
            assert out[k].dtype == torch.float32 and out[k].is_cuda
            assert torch.equal(out[k].cpu(), want), k
    assert set(pinned) == {"rgb", "depth", "pose"} and pinned["rgb"][0][0].is_pinned()
    assert pinned["rgb"][0][0].dtype == torch.uint8  # staged in the source dtype: 4x fewer H2D bytes than the reference


def test_batch_obs_keep_dtypes_for_compact_storage():
    from avlen_b200.common.utils import batch_obs
    rng = np.random.default_rng(1)
    obs = _obs_list(rng, 4)
    out = batch_obs(obs, device=torch.device("cuda"), pinned={}, keep_dtypes={"rgb": torch.uint8, "depth": torch.float16})
    assert out["rgb"].dtype == torch.uint8 and out["depth"].dtype == torch.float16 and out["pose"].dtype == torch.float32
    assert torch.equal(out["rgb"].cpu(), torch.from_numpy(np.stack([o["rgb"] for o in obs])))
    assert torch.equal(out["depth"].cpu(), torch.from_numpy(np.stack([o["depth"] for o in obs])).half())


def test_host_buffer_env_goes_through_batch_obs_and_matches_resident_env():
    """The e2e path of bench.py: SyntheticVectorEnv(host_buffers=True) hands batch_obs per-env numpy frames every step;
    the frames that reach the policy are the ones the device-resident env serves (same seed)."""
    from avlen_b200.synth_env import SyntheticVectorEnv
    a = SyntheticVectorEnv(4, "cuda", seed=11, host_buffers=True)
    b = SyntheticVectorEnv(4, "cuda", seed=11, host_buffers=False)
    oa, ob = a.reset(), b.reset()
    for _ in range(3):
        assert torch.equal(oa["rgb"], ob["rgb"]) and torch.equal(oa["depth"], ob["depth"])
        assert oa["rgb"].dtype == torch.float32
        act = torch.ones(4, 1, dtype=torch.int64, device="cuda")
        torch.manual_seed(3)
        oa, _, _ = a.step(act)
        torch.manual_seed(3)
        ob, _, _ = b.step(act)
    assert a._staging and a.h2d_bytes_per_step == 4 * (128 * 128 * 3 + 128 * 128 * 4)
    a.close(); b.close()


def test_resize_half_typed_and_indexed_matches_fp32_path():
    """the encoders' first kernel reads uint8 rgb / fp16 depth from the compact storage, optionally
    through a sample index (no minibatch copies) — bit-identical to converting to fp32 first."""
    from avlen_b200 import nn as K
    g = torch.Generator().manual_seed(0)
    rgb = torch.randint(0, 256, (7, 2, 16, 12, 3), generator=g, dtype=torch.uint8).cuda()      # (T+1, N, H, W, C)
    depth = torch.rand(7, 2, 16, 12, 1, generator=g).half().cuda()
    flat_rgb, flat_depth = rgb.view(-1, 16, 12, 3), depth.view(-1, 16, 12, 1)
    for x, scale, pad in ((flat_rgb, 1.0 / 255.0, 4), (flat_depth, 1.0, 4), (flat_rgb, 1.0 / 255.0, None)):
        want = K.resize_half(x.float().contiguous(), scale, pad)
        assert torch.equal(K.resize_half(x, scale, pad), want)
    idx = torch.tensor([13, 0, 5, 5, 2], device="cuda")
    for store, scale in ((rgb, 1.0 / 255.0), (depth, 1.0)):
        lazy = K.IndexedObservation(store, idx)
        assert lazy.shape == (5,) + tuple(store.shape[2:])
        want = K.resize_half(lazy.materialize().float().contiguous(), scale, 4)
        assert torch.equal(K.resize_half(lazy, scale, 4), want)


def test_compact_storage_update_equals_fp32_storage():
    """The same rollout written into an fp32 store and into the compact store (uint8 rgb / fp16 depth, lazy minibatch
    gather): PPO.update returns the same numbers (depth values chosen fp16-representable, so storage is lossless)."""
    from avlen_b200.common import spaces
    from avlen_b200.savi.models.rollout_storage import RolloutStorage
    from avlen_b200.savi.ppo.policy import AudioNavSMTPolicy
    from avlen_b200.savi.ppo.ppo import PPO
    T, N = 4, 4
    kw = dict(hidden_size=256, nhead=8, num_encoder_layers=1, num_decoder_layers=1, dropout=0.0, activation="relu",
              pretraining=False)
    results = []
    for compact in (False, True):
        torch.manual_seed(0)
        p = AudioNavSMTPolicy(spaces.savi_observation_space(), spaces.Discrete(4), **kw).cuda()
        p.net.freeze_encoders()
        p.net.set_eval_encoders()
        st = RolloutStorage(T, N, spaces.savi_observation_space(), spaces.Discrete(4), 512, True, 8, 4, 8, 4, 3, 3, 276, 276,
                            308, 256, num_recurrent_layers=1, max_dialog_len=77, compact_observations=compact)
        st.to(torch.device("cuda"))
        assert st.observations["rgb"].dtype == (torch.uint8 if compact else torch.float32)
        g = torch.Generator().manual_seed(5)
        z = lambda *s: torch.zeros(*s, device="cuda")  # noqa: E731
        for s in range(T):
            obs = {"rgb": torch.randint(0, 256, (N, 128, 128, 3), generator=g).float().cuda(),
                   "depth": (torch.randint(0, 256, (N, 128, 128, 1), generator=g).float() / 256).cuda(),
                   "spectrogram": torch.rand(N, 65, 26, 2, generator=g).cuda(), "pose": torch.randn(N, 4, generator=g).cuda(),
                   "category": z(N, 21), "category_belief": torch.rand(N, 21, generator=g).cuda(),
                   "location_belief": torch.randn(N, 2, generator=g).cuda()}
            st.insert(obs, z(1, N, 512), torch.randint(0, 4, (N, 1), generator=g).cuda(), None,
                      torch.randn(N, 1, generator=g).cuda(), torch.randn(N, 1, generator=g).cuda(),
                      torch.randn(N, 1, generator=g).cuda(), torch.ones(N, 1).cuda(), torch.ones(N, 1).cuda(),
                      torch.randn(N, 276, generator=g).cuda(), None, None, None, None, None, None, None, None, None, None,
                      None, None)
        st.compute_returns(z(N, 1), True, 0.99, 0.95)
        agent = PPO(p, 0.2, 1, 2, 0.5, 0.05, lr=2.5e-4, eps=1e-5, max_grad_norm=0.2, use_normalized_advantage=False)
        results.append(agent.update(st, perm_fn=lambda n: torch.arange(n)))
    for a, b in zip(*results):
        assert abs(a - b) <= 1e-6 * max(1.0, abs(b)), (results[0], results[1])
