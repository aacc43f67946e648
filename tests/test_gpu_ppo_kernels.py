"""The hand-written PPO kernels (bxg_gae, bxg_policy_act) against the framework-op statements of the same functions,
which tests/test_ppo_reference.py pins to the reference's own source."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _exact_fp32_framework_gemms():
  """The trainer switches the framework's GEMMs to TF32 (ppo.train); the statements these kernels are compared with
  must run in plain float32."""
  import torch
  old = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
  torch.backends.cuda.matmul.allow_tf32 = False
  torch.backends.cudnn.allow_tf32 = False
  yield
  torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


def test_gae_kernel_matches_compute_gae():
  import torch
  from brax_b200.training import fused, ppo
  dev = torch.device('cuda', 0)
  g = torch.Generator(device='cpu').manual_seed(0)
  for T, B in ((5, 2048), (1, 7), (20, 333)):
    trunc = (torch.rand((T, B), generator=g) < 0.15).float()
    done = torch.maximum(trunc, (torch.rand((T, B), generator=g) < 0.2).float())
    term = done * (1 - trunc)
    rew, val, boot = torch.randn((T, B), generator=g), torch.randn((T, B), generator=g), torch.randn((B,), generator=g)
    a = ppo.Agent(27, 8)
    vs_ref, adv_ref = a.gae_reference(trunc, term, rew, val, boot)
    vs, adv = fused.gae(*(t.to(dev) for t in (trunc, term, rew, val, boot)), a.lambda_, a.discounting)
    np.testing.assert_allclose(vs.cpu().numpy(), vs_ref.numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(adv.cpu().numpy(), adv_ref.numpy(), rtol=1e-5, atol=1e-5)
  # the golden vectors of the reference source too
  import os
  from tests.conftest import ROOT
  G = np.load(os.path.join(ROOT, 'tests', 'golden', 'ref_ppo.npz'))
  f = lambda k: torch.as_tensor(G[k], dtype=torch.float32, device=dev)   # noqa: E731
  vs, adv = fused.gae(f('gae_truncation'), f('gae_termination'), f('gae_rewards'), f('gae_values'), f('gae_bootstrap'), 0.95, 0.97)
  np.testing.assert_allclose(vs.cpu().numpy(), G['gae_vs'], rtol=1e-5, atol=1e-5)
  np.testing.assert_allclose(adv.cpu().numpy(), G['gae_advantages'], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize('obs_size,act_size', [(27, 8), (244, 17), (11, 3)])
def test_policy_act_kernel_matches_the_framework_ops(obs_size, act_size):
  import torch
  from brax_b200.training import fused, ppo
  dev = torch.device('cuda', 0)
  torch.manual_seed(1)
  a = ppo.Agent(obs_size, act_size).to(dev)
  a.update_normalization(torch.randn((5, 64, obs_size), device=dev) * 3 + 1)
  assert fused.supports(a.policy)
  for n in (1, 63, 4096):
    obs = torch.randn((n, obs_size), device=dev) * 3 + 1
    noise = torch.randn((n, act_size), device=dev)
    act, logits, pre = fused.policy_act(a.policy, a.running_mean, a.running_std, obs, noise)
    ref_logits = a.policy(a.normalize(obs))
    loc, scale = a.dist_create(ref_logits)
    ref_pre = loc + scale * noise
    np.testing.assert_allclose(logits.cpu().numpy(), ref_logits.detach().cpu().numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(pre.cpu().numpy(), ref_pre.detach().cpu().numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(act.cpu().numpy(), torch.tanh(ref_pre).detach().cpu().numpy(), rtol=1e-4, atol=1e-5)
  # a clipped normaliser and a policy the kernel is not compiled for
  a.clip_obs = 5.0
  obs = torch.randn((33, obs_size), device=dev) * 30
  act, logits, _ = fused.policy_act(a.policy, a.running_mean, a.running_std, obs, torch.zeros((33, act_size), device=dev), clip=5.0)
  np.testing.assert_allclose(logits.cpu().numpy(), a.policy(a.normalize(obs)).detach().cpu().numpy(), rtol=1e-4, atol=1e-4)
  assert not fused.supports(ppo.Agent(obs_size, act_size, hidden=(32, 32)).to(dev).policy)


@pytest.mark.parametrize('normalize_advantage', [True, False])
def test_ppo_head_kernel_matches_the_framework_loss_and_its_autograd_gradients(normalize_advantage):
  """bxg_ppo_head against Agent.loss in framework ops (which tests/test_ppo_reference.py pins to compute_ppo_loss run
  from the reference source): loss terms and the gradient of every network parameter."""
  import torch
  from brax_b200.training import ppo
  dev = torch.device('cuda', 0)
  torch.manual_seed(3)
  T, B, OBS, ACT = 5, 1536, 27, 8
  a = ppo.Agent(OBS, ACT, normalize_advantage=normalize_advantage).to(dev)
  a.update_normalization(torch.randn((T, B, OBS), device=dev) * 2 + 1)
  obs = torch.randn((T + 1, B, OBS), device=dev) * 2 + 1
  with torch.no_grad():
    beh = a.policy(a.normalize(obs[:-1])) + 0.05 * torch.randn((T, B, 2 * ACT), device=dev)
    loc, scale = a.dist_create(beh)
    pre = loc + scale * torch.randn_like(loc)
  trunc = (torch.rand((T, B), device=dev) < 0.15).float()
  done = torch.maximum(trunc, (torch.rand((T, B), device=dev) < 0.2).float())
  td = {'obs': obs, 'logits': beh, 'pre': pre, 'reward': torch.randn((T, B), device=dev), 'done': done, 'truncation': trunc}
  noise = torch.randn((T, B, ACT), device=dev)
  res = {}
  for fused_head in (False, True):
    a.fused_head = fused_head
    a.zero_grad(set_to_none=True)
    parts = a.loss(td, entropy_noise=noise, parts=True)
    parts[0].backward()
    res[fused_head] = ([float(p.detach()) for p in parts], [p.grad.detach().clone() for p in a.parameters()])
  for x, y in zip(*[res[k][0] for k in (False, True)]):
    assert abs(x - y) <= 2e-5 * max(1.0, abs(x)), (res[False][0], res[True][0])
  for gx, gy in zip(res[False][1], res[True][1]):
    scale = float(gx.abs().max()) + 1e-12
    assert float((gx - gy).abs().max()) <= 2e-4 * scale + 1e-7, (float((gx - gy).abs().max()), scale)
