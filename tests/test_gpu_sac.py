"""SAC smoke test on the fused env step (the second trainer BASELINE.json's north_star names)."""
import math

import pytest

pytestmark = pytest.mark.gpu


def test_sac_runs_and_reports_throughput():
  from brax_b200.training import sac
  net, m = sac.train('inverted_pendulum', num_timesteps=64 * 120, episode_length=100, num_envs=64, batch_size=128,
                     min_replay_size=64 * 20, grad_updates_per_step=2, normalize_observations=True, progress_every=20)
  assert m['iterations'] >= 100 and math.isfinite(m['critic_loss']) and math.isfinite(m['actor_loss']) and m['sps'] > 500
  assert m['alpha'] > 0
  for p in net.parameters():
    assert p.isfinite().all()
