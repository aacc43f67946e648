"""PPO smoke test on the fused env step."""
import math

import pytest

pytestmark = pytest.mark.gpu


def test_ppo_runs_and_reports_throughput():
  from brax_b200.training import ppo
  agent, m = ppo.train('ant', num_envs=256, episode_length=100, num_timesteps=256 * 5 * 4 * 3, unroll_length=5,
                       batch_size=64, num_minibatches=4, num_update_epochs=2)
  assert m['iterations'] >= 3 and math.isfinite(m['loss']) and m['sps'] > 1000
  for p in agent.parameters():
    assert p.isfinite().all()
