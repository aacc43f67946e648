"""SAC losses (reference brax/training/agents/sac/losses.py:30-131) on CPU tensors: shapes, the
truncation mask of the q error, and which parameters each loss reaches."""
import torch

from brax_b200.training import sac


def _batch(n, obs, act):
  g = torch.Generator().manual_seed(0)
  return {'obs': torch.randn(n, obs, generator=g), 'next_obs': torch.randn(n, obs, generator=g),
          'action': torch.rand(n, act, generator=g) * 2 - 1, 'reward': torch.randn(n, generator=g),
          'discount': torch.ones(n), 'truncation': torch.zeros(n)}


def test_losses_reach_the_right_parameters():
  torch.manual_seed(0)
  net = sac.SACNetworks(5, 2, hidden=(16, 16))
  a, c, p = sac.losses(net, _batch(32, 5, 2), reward_scaling=1.0, discounting=0.9, act_size=2)
  assert a.dim() == c.dim() == p.dim() == 0
  a.backward(inputs=[net.log_alpha]); c.backward(inputs=list(net.q.parameters())); p.backward(inputs=list(net.policy.parameters()))
  assert net.log_alpha.grad is not None and all(q.grad is not None for q in net.q.parameters())
  assert all(q.grad is not None for q in net.policy.parameters())
  assert all(q.grad is None for q in net.target_q.parameters())


def test_truncated_transitions_do_not_contribute_to_the_critic_loss():
  torch.manual_seed(0)
  net = sac.SACNetworks(5, 2, hidden=(16, 16))
  tr = _batch(32, 5, 2)
  tr['truncation'] = torch.ones(32)
  torch.manual_seed(1)
  _, c, _ = sac.losses(net, tr, 1.0, 0.9, 2)
  assert float(c.detach()) == 0.0


def test_alpha_loss_matches_eq_18():
  """alpha_loss = mean(alpha * (-log_prob - target_entropy)), target_entropy = -0.5 * act (losses.py:37,59)."""
  torch.manual_seed(0)
  net = sac.SACNetworks(5, 2, hidden=(16, 16))
  tr = _batch(64, 5, 2)
  torch.manual_seed(3)
  a, _, _ = sac.losses(net, tr, 1.0, 0.9, 2)
  torch.manual_seed(3)
  loc, scale = net.dist_params(tr['obs'])
  lp = net.log_prob(loc, scale, net.sample_pre_tanh(loc, scale))
  assert torch.allclose(a, (net.log_alpha.exp() * (-lp + 1.0)).mean())


def test_replay_buffer_is_a_ring():
  buf = sac.ReplayBuffer(10, 3, 1, 'cpu')
  for k in range(4):
    o = torch.full((4, 3), float(k))
    buf.insert(o, torch.zeros(4, 1), torch.zeros(4), torch.ones(4), o, torch.zeros(4))
  assert buf.size == 10 and buf.pos == 6
  assert set(buf.obs[:, 0].tolist()) == {1.0, 2.0, 3.0}      # the oldest batch was overwritten
  assert buf.sample(5)['obs'].shape == (5, 3)
