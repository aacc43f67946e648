"""PPO arithmetic against the reference's own source.

tests/golden/ref_ppo.npz was written by tools/gen_ppo_golden.py, which runs
`brax.training.agents.ppo.losses.compute_gae / compute_ppo_loss`, `brax.training.distribution`
and `brax.training.acme.running_statistics` unmodified from /root/reference on NumPy (float64).
`brax_b200.training.ppo.Agent` must reproduce them (float64 here, so the comparison is tight)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from brax_b200.training import ppo
from tests.conftest import ROOT

G = np.load(os.path.join(ROOT, 'tests', 'golden', 'ref_ppo.npz'))
T64 = lambda k: torch.as_tensor(G[k], dtype=torch.float64)   # noqa: E731
OBS, ACT, H = 27, 8, 32


def _agent(**kw):
  a = ppo.Agent(OBS, ACT, hidden=(H, H), entropy_cost=1e-2, discounting=0.97, reward_scaling=10.0, lambda_=0.95, epsilon=0.3, **kw).double()
  with torch.no_grad():
    for net, tag in ((a.policy, 'policy'), (a.value, 'value')):
      lin = [m for m in net if isinstance(m, torch.nn.Linear)]
      for i, m in enumerate(lin):
        m.weight.copy_(T64(f'{tag}_{2 * i}').T); m.bias.copy_(T64(f'{tag}_{2 * i + 1}'))
  return a


def _feed_normalizer(a, upto=3):
  for k in range(upto):
    a.update_normalization(T64(f'rs_batch{k}'))


def test_running_statistics_update_and_normalize():
  """acme/running_statistics.py:120-328 (Welford branch, std clipping, normalize)."""
  a = _agent()
  for k in range(3):
    a.update_normalization(T64(f'rs_batch{k}'))
    np.testing.assert_allclose(a.running_mean.numpy(), G[f'rs_mean{k}'], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(a.running_var.numpy(), G[f'rs_summed_variance{k}'], rtol=1e-12, atol=1e-10)
    np.testing.assert_allclose(a.running_std.numpy(), G[f'rs_std{k}'], rtol=1e-6)   # the reference divides by a float32 count
    assert float(a.num_steps) == float(G[f'rs_count{k}'])
  np.testing.assert_allclose(a.normalize(T64('rs_probe')).numpy(), G['rs_probe_normalized'], rtol=1e-6, atol=1e-9)
  a.clip_obs = 5.0
  np.testing.assert_allclose(a.normalize(T64('rs_probe')).numpy(), G['rs_probe_normalized_clip5'], rtol=1e-6, atol=1e-9)


def test_gae_matches_compute_gae():
  """agents/ppo/losses.py:38-101."""
  a = _agent()
  vs, adv = a.gae(T64('gae_truncation'), T64('gae_termination'), T64('gae_rewards'), T64('gae_values'), T64('gae_bootstrap'))
  np.testing.assert_allclose(vs.numpy(), G['gae_vs'], rtol=1e-12, atol=1e-12)
  np.testing.assert_allclose(adv.numpy(), G['gae_advantages'], rtol=1e-12, atol=1e-12)
  assert G['gae_truncation'].sum() > 0 and G['gae_termination'].sum() > 0      # both masks exercised


@pytest.mark.parametrize('tag,norm', [('adv_norm', True), ('no_adv_norm', False)])
def test_loss_matches_compute_ppo_loss(tag, norm):
  """agents/ppo/losses.py:143-303 incl. NormalTanhDistribution.log_prob / entropy (distribution.py:72-91,142-172)."""
  a = _agent(normalize_advantage=norm)
  _feed_normalizer(a)
  tm = lambda k: T64(k).transpose(0, 1).contiguous()     # noqa: E731  golden data is [B, T, ...]; ours is time-major
  td = {'obs': tm('loss_obs'), 'logits': tm('loss_behaviour_logits'), 'pre': tm('loss_raw_action'), 'reward': tm('loss_reward'),
        'done': tm('loss_done'), 'truncation': tm('loss_truncation')}
  # the behaviour log-prob the reference stores equals ours recomputed from the behaviour logits
  loc, scale = a.dist_create(td['logits'])
  np.testing.assert_allclose(a.dist_log_prob(loc, scale, td['pre']).numpy(), G['loss_behaviour_log_prob'].T, rtol=1e-9, atol=1e-9)
  total, pl, vl, el = a.loss(td, entropy_noise=T64('loss_entropy_noise'), parts=True)
  ref = {k: float(G[f'loss_{tag}_{k}']) for k in ('total', 'policy_loss', 'v_loss', 'entropy_loss')}
  # std in the normaliser comes through a float32 count in the reference: 1e-6 relative
  assert abs(float(pl.detach()) - ref['policy_loss']) <= 2e-6 * max(1.0, abs(ref['policy_loss']))
  assert abs(float(vl.detach()) - ref['v_loss']) <= 2e-6 * abs(ref['v_loss'])
  assert abs(float(el.detach()) - ref['entropy_loss']) <= 2e-6 * abs(ref['entropy_loss'])
  assert abs(float(total.detach()) - ref['total']) <= 2e-6 * abs(ref['total'])
  total.backward()      # differentiable end to end
  assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in a.parameters())


def _norm_worker(rank, world, port, ret):
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
  dist.init_process_group('gloo', rank=rank, world_size=world)
  try:
    a = _agent()
    for k in range(3):
      full = T64(f'rs_batch{k}')                                  # [T, B, obs]: each rank sees half of the envs
      a.update_normalization(full[:, rank::world].contiguous())
    ret[rank] = (a.running_mean.numpy().copy(), a.running_var.numpy().copy(), a.running_std.numpy().copy(), float(a.num_steps))
  finally:
    dist.destroy_process_group()


def test_normalizer_statistics_are_identical_across_ranks_and_match_the_global_batch():
  """The reference psums the step increment, the mean update and the variance update over the device axis
  (acme/running_statistics.py:174,262-264,270-271): every rank holds the statistics of the GLOBAL batch."""
  world = 2
  mgr = mp.Manager(); ret = mgr.dict()
  mp.spawn(_norm_worker, args=(world, 29531 + os.getpid() % 200, ret), nprocs=world, join=True)
  for j in range(3):
    np.testing.assert_array_equal(ret[0][j], ret[1][j])
  assert ret[0][3] == ret[1][3] == float(G['rs_count2'])
  np.testing.assert_allclose(ret[0][0], G['rs_mean2'], rtol=1e-10, atol=1e-10)
  np.testing.assert_allclose(ret[0][1], G['rs_summed_variance2'], rtol=1e-9, atol=1e-8)
  np.testing.assert_allclose(ret[0][2], G['rs_std2'], rtol=1e-6)
