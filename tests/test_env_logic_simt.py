"""Env epilogue logic (obs / reward / done / metrics / episode + auto-reset) on the
CPU: the kernel source through the host lane-group emulator vs the NumPy
restatement of the reference envs (oracle/env_oracle.py)."""
import numpy as np
import pytest

from oracle import oracle as O
from oracle.env_oracle import EnvOracle
from tests.simt.sim import SimEnv


CLASSIC = ('inverted_pendulum', 'inverted_double_pendulum', 'reacher', 'swimmer', 'humanoidstandup', 'pusher')


class _EnvStub:
  """What SimEnv needs from an env, built without touching CUDA."""

  def __init__(self, name, episode_length=None):
    from brax_b200 import envs_assets, native
    self.sys = envs_assets.load(name)
    sp = native.EnvSpecC()
    if name == 'ant':
      sp.kind, sp.forward_reward_weight, sp.ctrl_cost_weight, sp.healthy_reward = native.ENV_ROOT_VELOCITY, 1.0, 0.5, 1.0
      sp.healthy_z_min, sp.healthy_z_max = 0.2, 1.0
    elif name == 'halfcheetah':
      sp.kind, sp.forward_reward_weight, sp.ctrl_cost_weight, sp.healthy_reward = native.ENV_ROOT_VELOCITY, 1.0, 0.1, 0.0
      sp.healthy_z_min, sp.healthy_z_max = -3.0e38, 3.0e38
    elif name in ('hopper', 'walker2d'):
      sp.kind, sp.forward_reward_weight, sp.ctrl_cost_weight, sp.healthy_reward = native.ENV_PLANAR, 1.0, 1e-3, 1.0
      sp.healthy_z_min, sp.healthy_z_max = (0.7, 3.0e38) if name == 'hopper' else (0.8, 2.0)
      sp.healthy_angle_min, sp.healthy_angle_max = (-0.2, 0.2) if name == 'hopper' else (-1.0, 1.0)
      sp.healthy_state_min, sp.healthy_state_max = (-100.0, 100.0) if name == 'hopper' else (-3.0e38, 3.0e38)
    elif name in CLASSIC:
      sp.kind = {'inverted_pendulum': native.ENV_CARTPOLE, 'inverted_double_pendulum': native.ENV_DOUBLE_CARTPOLE,
                 'reacher': native.ENV_REACHER, 'swimmer': native.ENV_SWIMMER, 'humanoidstandup': native.ENV_STANDUP, 'pusher': native.ENV_PUSHER}[name]
      sp.forward_reward_weight, sp.ctrl_cost_weight, sp.healthy_reward = 1.0, (1e-4 if name == 'swimmer' else 0.0), 0.0
      sp.healthy_z_min, sp.healthy_z_max, sp.healthy_angle_max = -3.0e38, 3.0e38, 0.2
      if name == 'inverted_double_pendulum':
        sp.healthy_reward, sp.healthy_z_min, sp.tip_link = 10.0, 1.0, 2
        sp.tip_pos[2] = 0.6
      if name == 'reacher':
        sp.tip_link, sp.target_link = 1, 2
        sp.tip_pos[0] = 0.11
      if name == 'humanoidstandup':
        sp.ctrl_cost_weight, sp.healthy_reward = 0.01, 1.0
      if name == 'pusher':
        names = list(self.sys.link_names)
        sp.tip_link, sp.object_link, sp.target_link = names.index('r_wrist_flex_link'), names.index('object'), names.index('goal')
    else:
      sp.kind, sp.forward_reward_weight, sp.ctrl_cost_weight, sp.healthy_reward = native.ENV_COM_VELOCITY, 1.25, 0.1, 5.0
      sp.healthy_z_min, sp.healthy_z_max = 1.0, 2.0
    sp.obs_skip, sp.terminate_when_unhealthy = {'halfcheetah': (1, 0), 'hopper': (1, 1), 'walker2d': (1, 1), 'inverted_pendulum': (0, 0),
                                                'inverted_double_pendulum': (0, 0), 'reacher': (0, 0), 'swimmer': (2, 0), 'humanoidstandup': (2, 0), 'pusher': (0, 0)}.get(name, (2, 1))
    sp.episode_length = episode_length or 0
    self.n_frames = 4 if name in ('hopper', 'walker2d', 'swimmer') else (2 if name in CLASSIC[:3] else 5)
    sp.env_dt = float(np.float32(self.sys.opt.timestep) * np.float32(self.n_frames))
    self.spec = sp


def _oracle(name, sys, **kw):
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


_PLANAR_SLOTS = {'reward_forward': 0, 'reward_healthy': 1, 'reward_ctrl': 2, 'x_position': 4, 'x_velocity': 7}
_SLOTS = {'halfcheetah': {'reward_run': 0, 'reward_ctrl': 2, 'x_position': 4, 'x_velocity': 7},
          'hopper': _PLANAR_SLOTS, 'walker2d': _PLANAR_SLOTS,
          'swimmer': {'reward_fwd': 0, 'reward_ctrl': 2, 'x_position': 4, 'y_position': 5, 'distance_from_origin': 6,
                      'x_velocity': 7, 'y_velocity': 8, 'forward_reward': 9}}


@pytest.mark.parametrize('name', ['ant', 'humanoid', 'halfcheetah', 'hopper', 'walker2d', *CLASSIC])
def test_reset_obs_and_step_outputs(name):
  from brax_b200 import workloads
  stub = _EnvStub(name, episode_length=1000)
  n = 8
  if name in ('halfcheetah', 'hopper', 'walker2d'):   # reset noise as the reference envs, dropped towards the floor so that the capsules touch
    rng0 = np.random.default_rng(3)
    noise = 0.1 if name == 'halfcheetah' else 5e-3
    q = (np.asarray(stub.sys.init_q)[None] + rng0.uniform(-noise, noise, (n, stub.sys.nq))).astype(np.float32)
    q[:, 1] -= {'halfcheetah': 0.3, 'hopper': 0.03, 'walker2d': 0.03}[name]
    if name == 'hopper':
      q[0, 2] = 0.5      # root angle outside (-0.2, 0.2): unhealthy, terminates
    qd = (noise * rng0.standard_normal((n, stub.sys.nv))).astype(np.float32)
  elif name in CLASSIC:   # reset noise of the reference envs; one pendulum env starts tipped over (terminates at once)
    rng0 = np.random.default_rng(3)
    noise = {'inverted_pendulum': 0.01, 'inverted_double_pendulum': 0.01, 'reacher': 0.1, 'swimmer': 0.1, 'humanoidstandup': 0.01, 'pusher': 0.005}[name]
    q = (np.asarray(stub.sys.init_q)[None] + rng0.uniform(-noise, noise, (n, stub.sys.nq))).astype(np.float32)
    qd = rng0.uniform(-noise, noise, (n, stub.sys.nv)).astype(np.float32)
    if name == 'inverted_pendulum':
      q[0, 1] = 0.5
    if name == 'inverted_double_pendulum':
      q[0, 1] = 1.5
    if name == 'reacher':
      qd[:, 2:] = 0.0
    if name == 'pusher':   # arm lowered onto the table, the object next to the wrist (both contact kinds active for some envs)
      q[:, 1] = 0.40 + 0.05 * rng0.uniform(size=n)
      w = O.Oracle(stub.sys, np.float32).init(q, qd)['x_pos'][:, 6]
      q[:, 7], q[:, 8] = w[:, 1] + 0.07, w[:, 0] - 0.40
  else:
    _, q, qd = workloads.reset(name, 0, n, 0, 'cpu')
    q, qd = q.numpy(), qd.numpy()
  sim, orc = SimEnv(stub), _oracle(name, stub.sys, episode_length=1000)
  st, obs = sim.reset(q, qd)
  env = orc.reset(q, qd)
  np.testing.assert_allclose(obs, env['obs'], rtol=1e-5, atol=1e-6)
  done, steps = np.zeros(n, np.float32), np.zeros(n, np.float32)
  rng = np.random.default_rng(0)
  for k in range(4):
    act = rng.uniform(-1, 1, (n, stub.sys.nu)).astype(np.float32)
    # one-step map from the oracle's pipeline state (physics sensitivity: test_fp_sensitivity.py)
    st_in = {f: env['ps'][f].copy() for f in st}
    out, io = sim.step(st_in, act, env['done'], env['steps'])
    env = orc.step(env, act)
    e = np.maximum((np.abs(out['q'] - env['ps']['q']) / (1e-5 + 1e-4 * np.abs(env['ps']['q']))).max(1),
                   (np.abs(out['qd'] - env['ps']['qd']) / (1e-5 + 1e-4 * np.abs(env['ps']['qd']))).max(1))
    ok = e <= 1.0     # envs inside the stated physics tolerance (same solver branch)
    assert ok.mean() >= 0.6      # (8 envs: float32 rounding decides a line-search branch for a few of them)
    np.testing.assert_allclose(io['obs'][ok], env['obs'][ok], rtol=2e-3, atol=2e-3)
    np.testing.assert_allclose(io['reward'][ok], env['reward'][ok], rtol=1e-3, atol=5e-3)
    np.testing.assert_array_equal(io['done'][ok], env['done'][ok])
    np.testing.assert_array_equal(io['steps'], env['steps'])
    names = list(env['metrics'])
    for i, nm in enumerate(names):
      slot = _SLOTS.get(name, {}).get(nm, i)
      np.testing.assert_allclose(io['metrics'][ok, slot], env['metrics'][nm][ok], rtol=1e-3, atol=5e-3, err_msg=nm)


def test_episode_truncation_and_auto_reset():
  """episode_length = 3: step 3 sets done with truncation = 1 - done_env, the pipeline state
  and obs snap back to the first state, and `steps` restarts on the following step."""
  from brax_b200 import workloads
  stub = _EnvStub('ant', episode_length=3)
  n = 4
  _, q, qd = workloads.reset('ant', 0, n, 0, 'cpu')
  sim, orc = SimEnv(stub), _oracle('ant', stub.sys, episode_length=3, auto_reset=True)
  st, obs = sim.reset(q.numpy(), qd.numpy())
  first, first_obs = {k: v.copy() for k, v in st.items()}, obs.copy()
  env = orc.reset(q.numpy(), qd.numpy())
  done, steps = np.zeros(n, np.float32), np.zeros(n, np.float32)
  act = np.zeros((n, 8), np.float32)
  for k in range(5):
    st, io = sim.step(st, act, done, steps, first=first, first_obs=first_obs)
    env = orc.step(env, act)
    done, steps = io['done'], io['steps']
    np.testing.assert_array_equal(steps, env['steps'])
    np.testing.assert_array_equal(done, env['done'])
    np.testing.assert_array_equal(io['truncation'], env['truncation'])
    if k == 2:
      assert (done == 1).all() and (io['truncation'] == 1).all() and (steps == 3).all()
      for f in ('q', 'qd', 'mass_mx_inv', 'con_jac', 'x_pos'):
        np.testing.assert_array_equal(st[f], first[f])
      np.testing.assert_array_equal(io['obs'], first_obs)
    if k == 3:
      assert (steps == 1).all() and (done == 0).all()


def test_unhealthy_env_terminates():
  from brax_b200 import workloads
  stub = _EnvStub('ant')
  _, q, qd = workloads.reset('ant', 0, 2, 0, 'cpu')
  q = q.numpy(); q[1, 2] = 3.0          # far above healthy_z_range -> done
  sim = SimEnv(stub)
  st, _ = sim.reset(q, qd.numpy())
  _, io = sim.step(st, np.zeros((2, 8), np.float32), np.zeros(2, np.float32), np.zeros(2, np.float32))
  assert io['done'][0] == 0 and io['done'][1] == 1
  assert io['metrics'][0, 1] == 1.0     # reward_survive
