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
