"""Host-side logic of the rollout adapter that needs no GPU."""
import torch

from mapdn_b200.rollout import DeviceReplayBuffer, translate_action


def test_translate_action_matches_reference_formula():
    # reference utilities/util.py:123-132: clamp to [-1,1], then 0.5*(a+1)*(high-low)+low
    a = torch.tensor([[-3.0, -1.0, -0.5, 0.0, 0.25, 1.0, 7.0]])
    out = translate_action(a, action_scale=0.8, action_bias=0.1)
    low, high = 0.1 - 0.8, 0.1 + 0.8
    exp = 0.5 * (torch.clamp(a, -1, 1) + 1.0) * (high - low) + low
    assert out.dtype == torch.float64 and torch.allclose(out, exp.double())
    assert out.min() >= low - 1e-6 and out.max() <= high + 1e-6      # fp32 policy output


def test_replay_ring_buffer_cpu():
    buf = DeviceReplayBuffer(size=10, n_agents=2, obs_dim=3, device=torch.device("cpu"))
    for k in range(4):
        o = torch.full((4, 2, 3), float(k))
        buf.add_batch(o, torch.zeros(4, 2), torch.full((4,), float(k)), o + 1, torch.zeros(4, dtype=torch.bool))
    assert len(buf) == 10 and buf.head == 6                      # 16 transitions through a ring of 10
    s = buf.sample(32)
    assert s["obs"].shape == (10, 2, 3) and set(s["reward"].tolist()) <= {0.0, 1.0, 2.0, 3.0}
    assert torch.all(s["next_obs"] == s["obs"] + 1)
