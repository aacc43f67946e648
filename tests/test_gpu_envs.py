"""Fused env step on the GPU (brax_b200.envs) vs the NumPy restatement of the
reference envs + wrappers (oracle/env_oracle.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


CLASSIC = ('inverted_pendulum', 'inverted_double_pendulum', 'reacher', 'swimmer', 'humanoidstandup', 'pusher')


def _oracle(name, sys, **kw):
  from oracle.env_oracle import EnvOracle
  if name in CLASSIC:
    return EnvOracle(sys, name, ctrl_cost_weight=1e-4 if name == 'swimmer' else 0.0,
                     n_frames={'swimmer': 4, 'humanoidstandup': 5, 'pusher': 5}.get(name, 2), **kw)
  if name == 'ant':
    return EnvOracle(sys, 'ant', ctrl_cost_weight=0.5, healthy_reward=1.0, healthy_z_range=(0.2, 1.0), **kw)
  if name == 'hopper':
    return EnvOracle(sys, 'hopper', forward_reward_weight=1.0, ctrl_cost_weight=1e-3, healthy_reward=1.0, n_frames=4,
                     healthy_z_range=(0.7, np.inf), healthy_angle_range=(-0.2, 0.2), healthy_state_range=(-100.0, 100.0), **kw)
  if name == 'walker2d':
    return EnvOracle(sys, 'walker2d', forward_reward_weight=1.0, ctrl_cost_weight=1e-3, healthy_reward=1.0, n_frames=4,
                     healthy_z_range=(0.8, 2.0), healthy_angle_range=(-1.0, 1.0), **kw)
  if name == 'halfcheetah':
    return EnvOracle(sys, 'halfcheetah', forward_reward_weight=1.0, ctrl_cost_weight=0.1, healthy_reward=0.0,
                     terminate_when_unhealthy=False, **kw)
  return EnvOracle(sys, 'humanoid', forward_reward_weight=1.25, ctrl_cost_weight=0.1, healthy_reward=5.0,
                   healthy_z_range=(1.0, 2.0), **kw)


def _state_from_oracle(torch, env_state_cls, ps_cls, o_env, dev, first=None):
  from oracle import oracle as O
  ps = ps_cls.from_flat({k: torch.as_tensor(o_env['ps'][k], device=dev) for k in O.STATE_FIELDS})
  info = {'steps': torch.as_tensor(o_env['steps'], device=dev), 'truncation': torch.as_tensor(o_env['truncation'], device=dev)}
  if first is not None:
    info.update(first)
  return env_state_cls(ps, torch.as_tensor(o_env['obs'], device=dev), torch.as_tensor(o_env['reward'], device=dev),
                       torch.as_tensor(o_env['done'], device=dev), {}, info)


@pytest.mark.parametrize('name', ['ant', 'humanoid', 'halfcheetah', 'hopper', 'walker2d', *CLASSIC])
def test_env_reset_and_step_match_reference_restatement(name):
  import torch
  from brax_b200 import envs
  from brax_b200.envs.base import State
  from brax_b200.generalized.base import State as PS
  n = 64
  env = envs.create(name, episode_length=1000, auto_reset=True, batch_size=n)
  assert env.action_size == env.sys.nu and env.observation_size == {
      'ant': 27, 'humanoid': 244, 'halfcheetah': 17, 'hopper': 11, 'walker2d': 17, 'inverted_pendulum': 4,
      'inverted_double_pendulum': 8, 'reacher': 11, 'swimmer': 8, 'humanoidstandup': 244, 'pusher': 23}[name]
  st = env.reset(0)
  dev = st.obs.device
  orc = _oracle(name, env.sys, episode_length=1000, auto_reset=True)
  o_env = orc.reset(st.pipeline_state.q.cpu().numpy(), st.pipeline_state.qd.cpu().numpy())
  np.testing.assert_allclose(st.obs.cpu().numpy(), o_env['obs'], rtol=1e-5, atol=1e-5)
  first = {'first_pipeline_state': st.info['first_pipeline_state'], 'first_obs': st.info['first_obs']}
  gen = torch.Generator(device='cpu').manual_seed(0)
  within = []
  for k in range(6):
    act = (torch.rand((n, env.action_size), generator=gen) * 2 - 1).to(dev)
    s_in = _state_from_oracle(torch, State, PS, o_env, dev, first)      # one-step map from the oracle's state
    s_out = env.step(s_in, act)
    o_env = orc.step(o_env, act.cpu().numpy())
    gq, gqd = s_out.pipeline_state.q.cpu().numpy(), s_out.pipeline_state.qd.cpu().numpy()
    e = np.maximum((np.abs(gq - o_env['ps']['q']) / (1e-5 + 1e-4 * np.abs(o_env['ps']['q']))).max(1),
                   (np.abs(gqd - o_env['ps']['qd']) / (1e-5 + 1e-4 * np.abs(o_env['ps']['qd']))).max(1))
    ok = e <= 1.0   # envs inside the stated physics tolerance (same solver branch)
    within.append(ok.mean())
    np.testing.assert_allclose(s_out.obs.cpu().numpy()[ok], o_env['obs'][ok], rtol=2e-3, atol=2e-3)
    np.testing.assert_allclose(s_out.reward.cpu().numpy()[ok], o_env['reward'][ok], rtol=1e-3, atol=5e-3)
    np.testing.assert_array_equal(s_out.done.cpu().numpy()[ok], o_env['done'][ok])
    np.testing.assert_array_equal(s_out.info['steps'].cpu().numpy(), o_env['steps'])
    for nm, v in s_out.metrics.items():
      np.testing.assert_allclose(v.cpu().numpy()[ok], o_env['metrics'][nm][ok], rtol=1e-3, atol=5e-3, err_msg=nm)
  assert np.mean(within) >= 0.8


def test_auto_reset_on_gpu():
  import torch
  from brax_b200 import envs
  n = 32
  env = envs.create('ant', episode_length=3, auto_reset=True, batch_size=n)
  st = env.reset(1)
  first_q = st.info['first_pipeline_state'].q.clone()
  act = torch.zeros((n, 8), device=st.obs.device)
  for k in range(4):
    st = env.step(st, act)
    if k == 2:
      assert (st.done == 1).all() and (st.info['truncation'] == 1).all()
      assert torch.equal(st.pipeline_state.q, first_q) and torch.equal(st.obs, st.info['first_obs'])
  assert (st.info['steps'] == 1).all() and (st.done == 0).all()


def test_training_wrap_mirrors_the_reference_entry_point():
  """envs.get_environment + wrappers.training.wrap (reference envs/__init__.py:51-107) give the same
  fused env as envs.create."""
  import torch
  from brax_b200 import envs
  from brax_b200.envs.wrappers import training
  n = 16
  a = envs.create('ant', episode_length=3, auto_reset=True, batch_size=n)
  b = training.wrap(envs.get_environment('ant', batch_size=n), episode_length=3)
  sa, sb = a.reset(5), b.reset(5)
  act = torch.zeros((n, 8), device=sa.obs.device)
  for _ in range(4):
    sa, sb = a.step(sa, act), b.step(sb, act)
  assert torch.equal(sa.obs, sb.obs) and torch.equal(sa.done, sb.done) and torch.equal(sa.info['steps'], sb.info['steps'])


def test_env_step_equals_pipeline_step_plus_obs():
  """The fused env step leaves exactly the pipeline state pipeline.step produces."""
  import torch
  from brax_b200 import envs
  from brax_b200.generalized import pipeline
  env = envs.get_environment('ant', batch_size=50)
  st = env.reset(3)
  act = torch.rand((50, 8), device=st.obs.device) * 2 - 1
  ps = pipeline.step(env.sys, st.pipeline_state, act, n_frames=5)
  st2 = env.step(st, act)
  assert torch.equal(st2.pipeline_state.q, ps.q) and torch.equal(st2.pipeline_state.mass_mx_inv, ps.mass_mx_inv)
  assert torch.equal(st2.obs, torch.cat([ps.q[:, 2:], ps.qd], 1))


def test_action_repeat_is_the_episode_wrappers_scan():
  """action_repeat = 2 (EpisodeWrapper, reference wrappers/training.py:98-112): two inner env steps with one
  action, rewards summed, `steps` advanced by 2, episode logic applied once."""
  import torch
  from brax_b200 import envs
  n = 64
  a = envs.create('ant', episode_length=5, action_repeat=2, auto_reset=True, batch_size=n)
  b = envs.create('ant', episode_length=1000, action_repeat=1, auto_reset=True, batch_size=n)
  sa, sb = a.reset(2), b.reset(2)
  gen = torch.Generator(device='cpu').manual_seed(1)
  ok = torch.ones(n, dtype=torch.bool, device=sa.obs.device)   # envs the reference env never terminated (b resets those at once, a after the pair)
  for k in range(3):
    act = 0.3 * (torch.rand((n, 8), generator=gen) * 2 - 1).to(sa.obs.device)
    sb1 = b.step(sb, act); sb = b.step(sb1, act)
    sa = a.step(sa, act)
    ok &= (sb1.done == 0) & ((sb.done == 0) | (k == 2))
    torch.testing.assert_close(sa.reward[ok], (sb1.reward + sb.reward)[ok])
    if k < 2:
      assert torch.equal(sa.pipeline_state.q[ok], sb.pipeline_state.q[ok]) and torch.equal(sa.obs[ok], sb.obs[ok])
      assert (sa.info['steps'][ok] == 2 * (k + 1)).all() and (sa.done[ok] == sb.done[ok]).all()
  assert ok.sum() >= n // 2
  # third call: steps = 6 >= episode_length = 5 -> truncated, state snaps back to the first state
  assert (sa.done[ok] == 1).all() and (sa.info['truncation'][ok] == (1 - sb.done)[ok]).all() and (sa.info['steps'][ok] == 6).all()
  assert torch.equal(sa.obs[ok], sa.info['first_obs'][ok])


@pytest.mark.parametrize('name', ['ant', 'hopper'])
def test_episode_metrics_follow_the_episode_wrapper(name):
  """info['episode_done'] / info['episode_metrics'] as EpisodeWrapper maintains them (wrappers/training.py:83-127):
  sums restart after an episode that ended, `length` counts steps, every env metric is accumulated."""
  import torch
  from brax_b200 import envs
  n, ep_len = 32, 3
  env = envs.create(name, episode_length=ep_len, auto_reset=True, batch_size=n)
  st = env.reset(1)
  em = st.info['episode_metrics']
  assert set(em) == {'sum_reward', 'length', *env.metric_names} and float(st.info['episode_done'].sum()) == 0
  exp = {k: np.zeros(n) for k in em}
  prev_done = np.zeros(n)
  gen = torch.Generator(device='cpu').manual_seed(0)
  for k in range(8):
    act = (torch.rand((n, env.action_size), generator=gen) * 2 - 1).to(st.obs.device)
    st = env.step(st, act)
    keep = 1.0 - prev_done
    exp['sum_reward'] = exp['sum_reward'] * keep + st.reward.cpu().numpy()
    exp['length'] = exp['length'] * keep + 1
    for m in env.metric_names:
      if m != 'reward':
        exp[m] = exp[m] * keep + st.metrics[m].cpu().numpy()
    prev_done = st.done.cpu().numpy()
    for m, v in exp.items():
      np.testing.assert_allclose(st.info['episode_metrics'][m].cpu().numpy(), v, rtol=1e-5, atol=1e-5, err_msg=f'{name} step {k} {m}')
    np.testing.assert_array_equal(st.info['episode_done'].cpu().numpy(), prev_done)
  assert float(st.info['episode_metrics']['length'].max()) <= ep_len
