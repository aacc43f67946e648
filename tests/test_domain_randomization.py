"""Domain randomisation: one System per env (reference envs/wrappers/training.py:223-260, DomainRandomizationVmapWrapper).

tests/golden/ref_dr_<env>.npz (tools/gen_reference_golden_dr.py) is the reference's own wrapper stack --
AutoReset(Episode(DomainRandomizationVmap(env))) -- run from the reference source in float64 with per-env link centres
of mass, masses, inertias, actuator gears, joint damping and armature.  CPU side: the host logic that turns
(`sys_v`, `in_axes`) into per-env Systems, the kernel source in double precision with env e's System (exact), the
float32 env epilogue through the emulator, and the float32 oracle.  GPU side: tests/test_gpu_domain_randomization.py.
"""
import os

import numpy as np
import pytest

from brax_b200 import base, envs_assets
from oracle import oracle as O
from tests.conftest import ROOT
from tests.simt import sim as S
from tests.test_env_logic_simt import _EnvStub, _oracle

f32 = np.float32
NAMES = ['ant', 'humanoid']


def load(name):
  return np.load(os.path.join(ROOT, 'tests', 'golden', f'ref_dr_{name}.npz'))


def randomization_fn_from_golden(g):
  """A `randomization_fn(sys) -> (sys_v, in_axes)` in the reference's form (ppo/train_test.py:229-239) that reproduces
  the golden file's randomised leaves."""
  leaves = {k[len('rand_'):]: g[k].astype(f32) for k in g.files if k.startswith('rand_')}

  def fn(sys):
    sys_v = sys.tree_replace(leaves)
    in_axes = base.tree_map(lambda x: None, sys)
    in_axes = in_axes.tree_replace({k: 0 for k in leaves})
    return sys_v, in_axes
  return fn, leaves


def systems(name, g):
  s = envs_assets.load(name)
  fn, leaves = randomization_fn_from_golden(g)
  return s, base.unbatch(*fn(s)), leaves


def ps_at(g, k, sl, dtype, shapes):
  """The reference's pipeline state before env-step k (k = 0: after reset) for the envs in `sl`."""
  p = 'init_ps_' if k == 0 else f'step{k - 1}_ps_'
  return {f: np.ascontiguousarray(g[p + f][sl].reshape((-1,) + shapes[f][1:]).astype(dtype)) for f in O.STATE_FIELDS}


def test_unbatch_slices_only_the_mapped_leaves():
  g = load('ant')
  s, per_env, leaves = systems('ant', g)
  assert len(per_env) == g['q0'].shape[0]
  for e, se in enumerate(per_env):
    np.testing.assert_array_equal(se.link.inertia.mass, leaves['link.inertia.mass'][e])
    np.testing.assert_array_equal(se.link.inertia.transform.pos, leaves['link.inertia.transform.pos'][e])
    np.testing.assert_array_equal(se.actuator.gear, leaves['actuator.gear'][e])
    assert se.link.inertia.i.shape == s.link.inertia.i.shape
    assert se.dof.stiffness is s.dof.stiffness and se.link.transform.pos is s.link.transform.pos   # shared, not copied
    assert se.link_parents == s.link_parents and se.nv == s.nv
  with pytest.raises(ValueError):
    base.unbatch(s.tree_replace({'dof.damping': np.zeros((3, s.nv), f32), 'dof.armature': np.zeros((4, s.nv), f32)}),
                 base.tree_map(lambda x: None, s).tree_replace({'dof.damping': 0, 'dof.armature': 0}))
  with pytest.raises(NotImplementedError):
    base.unbatch(s.tree_replace({'dof.damping': np.zeros((s.nv, 3), f32)}), base.tree_map(lambda x: None, s).tree_replace({'dof.damping': 1}))


@pytest.mark.parametrize('name', NAMES)
def test_float64_kernel_source_with_each_envs_system_is_exact(name):
  """Every env-step of the reference wrapper's run as a one-step map, env e stepped with ITS System: the kernel source
  in double precision reproduces q, qd, x, xd and mass_mx_inv to 1e-9; with the nominal System instead it does not."""
  g = load(name)
  s, per_env, _ = systems(name, g)
  n, steps = g['q0'].shape[0], g['act'].shape[0]
  n_frames = 5
  nominal_err = 0.0
  act = g['act']
  if name == 'humanoid':   # the env maps [-1, 1] onto the actuators' ctrl_range before pipeline_step (envs/humanoid.py:262-264)
    lo, hi = np.asarray(s.actuator.ctrl_range, np.float64).reshape(-1, 2).T
    act = (act + 1) * (hi - lo) * 0.5 + lo
  for e in range(n):
    sim, sim_nom = S.Sim(per_env[e], dtype=np.float64), S.Sim(s, dtype=np.float64)
    shapes = {f: v.shape for f, v in sim.alloc(1).items()}
    for k in range(steps):
      if g[f'step{k}_done'][e]:      # auto-reset replaced the state
        continue
      st_in = ps_at(g, k, slice(e, e + 1), np.float64, shapes)
      out = sim.step(st_in, act[k, e:e + 1], n_frames)
      out_nom = sim_nom.step(st_in, act[k, e:e + 1], n_frames)
      for leaf in ('q', 'qd', 'x_pos', 'x_rot', 'xd_ang', 'xd_vel', 'mass_mx_inv', 'qf_smooth', 'con_aref'):
        ref = g[f'step{k}_ps_{leaf}'][e].reshape(out[leaf].shape)
        scale = max(1.0, float(np.abs(ref).max()))
        assert np.abs(out[leaf] - ref).max() <= 1e-9 * scale, (name, e, k, leaf, np.abs(out[leaf] - ref).max())
      nominal_err = max(nominal_err, float(np.abs(out_nom['qd'] - g[f'step{k}_ps_qd'][e]).max()))
  assert nominal_err > 1e-3, nominal_err    # the randomisation matters: the nominal System is measurably wrong


@pytest.mark.parametrize('name', NAMES)
def test_float32_env_step_with_each_envs_system(name):
  """obs / reward / done / episode bookkeeping of the float32 kernel source (emulator) and of the float32 oracle, env e
  with its own System, against the reference wrapper stack; one-env-step maps from the reference's state."""
  g = load(name)
  s, per_env, _ = systems(name, g)
  n, steps, ep_len = g['q0'].shape[0], g['act'].shape[0], int(g['episode_length'])
  for e in range(n):
    stub = _EnvStub(name, episode_length=ep_len)
    stub.sys = per_env[e]
    env = S.SimEnv(stub)
    orc = _oracle(name, per_env[e], episode_length=ep_len, auto_reset=True)
    sl = slice(e, e + 1)
    first, first_obs = env.reset(g['q0'][sl].astype(f32), g['qd0'][sl].astype(f32))
    np.testing.assert_allclose(first_obs, g['obs0'][sl], rtol=1e-4, atol=1e-5)
    oenv = orc.reset(g['q0'][sl].astype(f32), g['qd0'][sl].astype(f32))
    np.testing.assert_allclose(oenv['obs'], g['obs0'][sl], rtol=1e-4, atol=1e-5)
    shapes = {f: v.shape for f, v in first.items()}
    for k in range(steps):
      st_in = ps_at(g, k, sl, f32, shapes)
      done = np.zeros(1, f32) if k == 0 else g[f'step{k - 1}_done'][sl].astype(f32)
      nsteps = np.zeros(1, f32) if k == 0 else g[f'step{k - 1}_steps'][sl].astype(f32)
      out, io = env.step(st_in, g['act'][k, sl].astype(f32), done, nsteps, first=first, first_obs=first_obs)
      p = f'step{k}_'
      np.testing.assert_array_equal(io['done'], g[p + 'done'][sl]); np.testing.assert_array_equal(io['steps'], g[p + 'steps'][sl])
      np.testing.assert_array_equal(io['truncation'], g[p + 'truncation'][sl])
      np.testing.assert_allclose(out['q'], g[p + 'q'][sl], rtol=2e-3, atol=2e-4)
      np.testing.assert_allclose(io['obs'], g[p + 'obs'][sl], rtol=5e-3, atol=5e-3)
      np.testing.assert_allclose(io['reward'], g[p + 'reward'][sl], rtol=2e-3, atol=5e-3)
      # the oracle
      for f in O.STATE_FIELDS:
        oenv['ps'][f] = st_in[f].copy()
      oenv.update(obs=(g['obs0'] if k == 0 else g[f'step{k - 1}_obs'])[sl].astype(f32), done=done.copy(), steps=nsteps.copy(),
                  truncation=np.zeros(1, f32), first_ps=oenv['first_ps'], first_obs=oenv['first_obs'])
      oenv = orc.step(oenv, g['act'][k, sl].astype(f32))
      np.testing.assert_array_equal(oenv['done'], g[p + 'done'][sl])
      np.testing.assert_allclose(oenv['ps']['q'], g[p + 'q'][sl], rtol=2e-3, atol=2e-4)
      np.testing.assert_allclose(oenv['reward'], g[p + 'reward'][sl], rtol=2e-3, atol=5e-3)
