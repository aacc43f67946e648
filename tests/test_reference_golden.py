"""The oracle against golden vectors produced by the REFERENCE'S OWN SOURCE.

tests/golden/ref_<model>.npz were written by tools/gen_reference_golden.py, which runs
`brax.generalized.pipeline.init / step` unmodified from /root/reference on NumPy (float64)
through the stand-ins in tools/refshim/.  That pins, against the reference's code itself,
every function of SURVEY.md section 8(a) the oracle restates: kinematics.forward,
dynamics.transform_com / inverse / forward, mass.matrix / matrix_inv (math.inv_approximate),
constraint.jacobian (jac_contact, jac_limit, _imp_aref, point_jacobian), the A and b of
constraint.force, integrator.integrate, actuator.to_tau, scan.tree / link_types ordering and
pipeline orchestration.  Still restated on both sides (see tools/refshim/README.md): the mjx
colliders, jaxopt's projected gradient, the MuJoCo model compiler."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests.conftest import ROOT

MODELS = ['ant', 'humanoid', 'halfcheetah', 'hopper', 'walker2d', 'humanoidstandup', 'pusher', 'triple_pendulum_motor', 'inverted_pendulum',
          'inverted_double_pendulum', 'reacher', 'swimmer', 'two_trees']
ASSETS = ('ant', 'humanoid', 'halfcheetah', 'hopper', 'walker2d', 'humanoidstandup', 'pusher')


def _load(name):
  from brax_b200 import envs_assets
  from brax_b200.io import model_json
  g = np.load(os.path.join(ROOT, 'tests', 'golden', f'ref_{name}.npz'))
  if name == 'two_trees':
    from brax_b200.io import mjcf
    from tests.synthetic_models import TWO_TREES_XML
    return mjcf.loads(TWO_TREES_XML), g
  if name not in ASSETS:
    return model_json.load(os.path.join(ROOT, 'tests', 'golden', f'{name}.json')), g
  return envs_assets.load(name), g


def _close(a, b, what, rtol=1e-9, atol=1e-9):
  scale = max(1.0, float(np.abs(b).max())) if b.size else 1.0
  np.testing.assert_allclose(a, b, rtol=rtol, atol=atol * scale, err_msg=what)


@pytest.mark.parametrize('name', MODELS)
def test_init_matches_reference_source(name):
  s, g = _load(name)
  st = O.Oracle(s, np.float64).init(g['q0'], g['qd0'])
  for f in O.STATE_FIELDS:
    _close(st[f], g[f'init_{f}'].reshape(st[f].shape), f'{name} init {f}')


@pytest.mark.parametrize('name', MODELS)
def test_every_step_is_a_one_step_map_of_the_reference_state(name):
  """step k of the oracle from the REFERENCE's state k-1 equals the reference's state k,
  for every leaf (float64, 1e-9): no room for a restatement error to hide in chaos."""
  s, g = _load(name)
  o = O.Oracle(s, np.float64)
  steps = g['act'].shape[0]
  st = o.init(g['q0'], g['qd0'])
  active = 0
  for k in range(steps):
    prev = 'init' if k == 0 else f'step{k - 1}'
    for f in O.STATE_FIELDS:
      st[f][...] = g[f'{prev}_{f}'].reshape(st[f].shape)
    o.step(st, g['act'][k], 1)
    for f in O.STATE_FIELDS:
      _close(st[f], g[f'step{k}_{f}'].reshape(st[f].shape), f'{name} step {k} {f}', rtol=1e-8, atol=1e-8)
    active += int((g[f'step{k}_con_diag'] != 0).sum())
  if name in ASSETS or name == 'two_trees':
    assert active > 0      # contact / limit rows were exercised


@pytest.mark.parametrize('name', MODELS)
def test_free_running_rollout_tracks_the_reference(name):
  s, g = _load(name)
  o = O.Oracle(s, np.float64)
  st = o.init(g['q0'], g['qd0'])
  steps = g['act'].shape[0]
  for k in range(steps):
    o.step(st, g['act'][k], 1)
  for f in ('q', 'qd', 'x_pos', 'x_rot', 'xd_ang', 'xd_vel'):
    _close(st[f], g[f'step{steps - 1}_{f}'].reshape(st[f].shape), f'{name} rollout {f}', rtol=1e-6, atol=1e-6)


# ---------------------------------------------------------------- envs + wrappers
ENVS = {
    'ant': dict(kind='ant', ctrl_cost_weight=0.5, healthy_reward=1.0, healthy_z_range=(0.2, 1.0)),
    'humanoid': dict(kind='humanoid', forward_reward_weight=1.25, ctrl_cost_weight=0.1, healthy_reward=5.0, healthy_z_range=(1.0, 2.0)),
    'halfcheetah': dict(kind='halfcheetah', forward_reward_weight=1.0, ctrl_cost_weight=0.1, healthy_reward=0.0,
                        terminate_when_unhealthy=False),
    'hopper': dict(kind='hopper', forward_reward_weight=1.0, ctrl_cost_weight=1e-3, healthy_reward=1.0, n_frames=4,
                   healthy_z_range=(0.7, np.inf), healthy_angle_range=(-0.2, 0.2), healthy_state_range=(-100.0, 100.0)),
    'walker2d': dict(kind='walker2d', forward_reward_weight=1.0, ctrl_cost_weight=1e-3, healthy_reward=1.0, n_frames=4,
                     healthy_z_range=(0.8, 2.0), healthy_angle_range=(-1.0, 1.0)),
    'humanoidstandup': dict(kind='humanoidstandup', n_frames=5),
    'pusher': dict(kind='pusher', n_frames=5),
    'inverted_pendulum': dict(kind='inverted_pendulum', n_frames=2),
    'inverted_double_pendulum': dict(kind='inverted_double_pendulum', n_frames=2),
    'reacher': dict(kind='reacher', n_frames=2),
    'swimmer': dict(kind='swimmer', forward_reward_weight=1.0, ctrl_cost_weight=1e-4, n_frames=4),
}


@pytest.mark.parametrize('name', sorted(ENVS))
def test_env_oracle_matches_reference_envs_and_wrappers(name):
  """oracle/env_oracle.py (float32) against `brax.envs.<env>` wrapped by `training.wrap`
  (VmapWrapper + EpisodeWrapper + AutoResetWrapper), all run from the reference's source in
  float64: reset observation, then every env step as a one-step map from the reference's
  state.  Episode bookkeeping (done, steps, truncation, auto-reset selection) must be exact;
  obs / reward / metrics agree to float32 physics accuracy."""
  from brax_b200 import envs_assets
  from oracle.env_oracle import EnvOracle
  g = np.load(os.path.join(ROOT, 'tests', 'golden', f'ref_env_{name}.npz'))
  s = envs_assets.load(name)
  kw = dict(ENVS[name]); kind = kw.pop('kind')
  orc = EnvOracle(s, kind, episode_length=int(g['episode_length']), auto_reset=True, **kw)
  f32 = np.float32
  env = orc.reset(g['q0'].astype(f32), g['qd0'].astype(f32))
  np.testing.assert_allclose(env['obs'], g['obs0'], rtol=1e-4, atol=1e-5)
  first_ps, first_obs = env['first_ps'], env['first_obs']
  n, steps = g['q0'].shape[0], g['act'].shape[0]
  saw_done = saw_trunc = False
  for k in range(steps):
    if k > 0:   # continue from the reference's state
      p = f'step{k - 1}_'
      for f in O.STATE_FIELDS:
        env['ps'][f] = g[p + 'ps_' + f].reshape(env['ps'][f].shape).astype(f32)
      env.update(obs=g[p + 'obs'].astype(f32), done=g[p + 'done'].astype(f32), steps=g[p + 'steps'].astype(f32),
                 truncation=g[p + 'truncation'].astype(f32), first_ps=first_ps, first_obs=first_obs)
    env = orc.step(env, g['act'][k].astype(f32))
    p = f'step{k}_'
    np.testing.assert_array_equal(env['done'], g[p + 'done'], err_msg=f'{name} step {k} done')
    np.testing.assert_array_equal(env['steps'], g[p + 'steps'], err_msg=f'{name} step {k} steps')
    np.testing.assert_array_equal(env['truncation'], g[p + 'truncation'], err_msg=f'{name} step {k} truncation')
    np.testing.assert_allclose(env['ps']['q'], g[p + 'q'], rtol=2e-3, atol=2e-4, err_msg=f'{name} step {k} q')
    np.testing.assert_allclose(env['obs'], g[p + 'obs'], rtol=5e-3, atol=5e-3, err_msg=f'{name} step {k} obs')
    np.testing.assert_allclose(env['reward'], g[p + 'reward'], rtol=2e-3, atol=5e-3, err_msg=f'{name} step {k} reward')
    for m, v in env['metrics'].items():
      np.testing.assert_allclose(v, g[p + 'metric_' + m], rtol=2e-3, atol=5e-3, err_msg=f'{name} step {k} {m}')
    saw_done |= bool(g[p + 'done'].any()); saw_trunc |= bool(g[p + 'truncation'].any())
  assert saw_trunc and (saw_done or name in ('reacher', 'swimmer', 'humanoidstandup', 'pusher'))


@pytest.mark.parametrize('name', ['ant', 'humanoid', 'halfcheetah', 'hopper', 'two_trees', 'swimmer', 'reacher', 'inverted_double_pendulum', 'humanoidstandup', 'pusher'])
def test_kernel_source_against_reference_source_golden(name):
  """brax_b200/csrc/bxg_core.cuh (float32, through the host lane emulator) directly against the
  reference-source golden (float64): one-step maps inside the stated 1e-4 / 1e-5 tolerance."""
  from tests.simt.sim import Sim
  s, g = _load(name)
  sim = Sim(s)
  f32 = np.float32
  st = sim.init(g['q0'].astype(f32), g['qd0'].astype(f32))
  for f in O.STATE_FIELDS:
    ref = g[f'init_{f}'].reshape(st[f].shape)
    # capsule-capsule: mjx picks between two near-equal closest-point candidates (d1 < d2 differ by ~1e-11 when both
    # points are interior); float32 cannot resolve the tie, the float32 oracle lands on the same side as the kernel
    # and the contact normal moves by ~4e-5 rad
    loose = 10.0 if name == 'pusher' and f.startswith('con_') else 1.0
    np.testing.assert_allclose(st[f], ref, rtol=1e-4 * loose, atol=2e-5 * loose * max(1.0, float(np.abs(ref).max()) if ref.size else 1.0), err_msg=f)
  inside = []
  for k in range(g['act'].shape[0]):
    prev = 'init' if k == 0 else f'step{k - 1}'
    st_in = {f: np.ascontiguousarray(g[f'{prev}_{f}'].reshape(st[f].shape).astype(f32)) for f in O.STATE_FIELDS}
    out = sim.step(st_in, g['act'][k].astype(f32), 1)
    e = np.zeros(g['q0'].shape[0])
    for f in ('q', 'qd', 'x_pos', 'x_rot', 'xd_ang', 'xd_vel'):
      r = g[f'step{k}_{f}'].reshape(out[f].shape)
      e = np.maximum(e, (np.abs(out[f] - r) / (1e-5 + 1e-4 * np.abs(r))).reshape(len(e), -1).max(1))
    inside.append(e <= 1.0)
  assert np.mean(inside) >= 0.85, (name, np.mean(inside))
