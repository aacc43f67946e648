"""One System per env on the GPU (bxg_model_create_batched; reference DomainRandomizationVmapWrapper,
envs/wrappers/training.py:223-260).

The per-env kernels read each env's constants from its own model in global memory.  They must be, env by env, BIT FOR BIT
the host emulation of the same source with that env's System (whose double-precision build equals the reference wrapper's
run exactly: tests/test_domain_randomization.py), and the wrapped env must reproduce the reference wrapper stack's golden
run (tests/golden/ref_dr_<env>.npz) as one-env-step maps."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
f32 = np.float32
NAMES = ['ant', 'humanoid']


def _golden(name):
  from tests import test_domain_randomization as T
  g = T.load(name)
  s, per_env, _ = T.systems(name, g)
  return T, g, s, per_env


def _same_bits(got, ref, what):
  from tests.test_gpu_bitexact import _assert_same_bits
  _assert_same_bits(got, ref, what)


@pytest.mark.parametrize('name', NAMES)
def test_per_env_models_equal_host_emulation_with_each_envs_system_bit_for_bit(name):
  import torch
  from brax_b200 import native
  from tests.simt import sim as S
  T, g, s, per_env = _golden(name)
  n = len(per_env)
  dev = torch.device('cuda', 0)
  nm = native.BatchedNativeModel(per_env, 0)
  assert nm.num_models == n and nm.kernel_id == {'ant': 0, 'humanoid': 1}[name]
  q, qd = g['q0'].astype(f32), g['qd0'].astype(f32)
  q[:, 2] -= {'ant': 0.03, 'humanoid': 0.1}[name]      # feet on the floor: contacts active
  sims = [S.Sim(se) for se in per_env]
  st = nm.init(torch.as_tensor(q, device=dev), torch.as_tensor(qd, device=dev))
  hs = [sim.init(q[e:e + 1], qd[e:e + 1]) for e, sim in enumerate(sims)]
  for f in native.STATE_FIELDS:
    _same_bits(st[f].cpu().numpy(), np.concatenate([h[f] for h in hs]), f'{name} init {f}')
  lean = {k: st[k].clone() for k in native.LEAN_FIELDS}
  active = 0
  for k in range(4):
    act = g['act'][k].astype(f32) * (0.4 if name == 'humanoid' else 1.0)
    a = torch.as_tensor(act, device=dev)
    diag = nm.alloc_diag(n)
    st = nm.step(st, a, 5, diag=diag)
    hs = [sim.step(h, act[e:e + 1], 5, diag=True) for e, (sim, h) in enumerate(zip(sims, hs))]
    for f in native.STATE_FIELDS:
      _same_bits(st[f].cpu().numpy(), np.concatenate([h[f] for h in hs]), f'{name} env-step {k} {f}')
    _same_bits(diag['stats'].cpu().numpy(), np.concatenate([h['stats'] for h in hs]), f'{name} env-step {k} branch counters')
    active += int(sum((h['con_dist'] < 0).sum() for h in hs))
    hs = [{f: h[f] for f in native.STATE_FIELDS} for h in hs]
    # lean state I/O on the per-env kernels: the same bits
    lean = nm.step(lean, a, 5, lean=True)
    for f in native.LEAN_FIELDS:
      assert torch.equal(lean[f], st[f]), (name, k, f)
  assert active > 0
  # the models differ: the nominal System gives other numbers
  nom = native.NativeModel(s, 0)
  a0 = nom.init(torch.as_tensor(q, device=dev), torch.as_tensor(qd, device=dev))
  b0 = nm.init(torch.as_tensor(q, device=dev), torch.as_tensor(qd, device=dev))
  assert not torch.equal(a0['mass_mx'], b0['mass_mx'])


@pytest.mark.parametrize('name', NAMES)
def test_identical_per_env_models_equal_the_shared_model(name):
  """n copies of one System through the per-env kernels = the default kernels (shared-memory model, specialised build)."""
  import torch
  from brax_b200 import envs_assets, native
  from tests.test_gpu_bitexact import _inputs
  n = 37
  s, q, qd, acts, nf = _inputs(name, n, seed=6)
  dev = torch.device('cuda', 0)
  shared, per_env = native.NativeModel(s, 0), native.BatchedNativeModel([s] * n, 0)
  a = shared.init(torch.as_tensor(q, device=dev), torch.as_tensor(qd, device=dev))
  b = per_env.init(torch.as_tensor(q, device=dev), torch.as_tensor(qd, device=dev))
  for act in acts:
    t = torch.as_tensor(act, device=dev)
    a, b = shared.step(a, t, nf), per_env.step(b, t, nf)
    for f in native.STATE_FIELDS:
      assert torch.equal(a[f], b[f]), (name, f)
  with pytest.raises(RuntimeError, match='one env per model'):
    per_env.init(torch.as_tensor(q[:5], device=dev), torch.as_tensor(qd[:5], device=dev))


def test_models_of_another_topology_are_rejected():
  from brax_b200 import envs_assets, native
  s = envs_assets.load('ant')
  other = s.tree_replace({'opt.timestep': np.float32(0.5) * np.float32(s.opt.timestep)})
  with pytest.raises(RuntimeError, match='differs from model 0'):
    native.BatchedNativeModel([s, other], 0)


@pytest.mark.parametrize('name', NAMES)
def test_wrapped_env_reproduces_the_reference_wrapper_stack(name):
  """training.wrap(env, randomization_fn=...) against AutoReset(Episode(DomainRandomizationVmap(env))) run from the
  reference source: reset observation, then every env-step as a one-step map from the reference's state."""
  import torch
  from brax_b200 import envs
  from brax_b200.envs.base import State
  from brax_b200.envs.wrappers import training
  from brax_b200.generalized.base import State as PS
  from oracle import oracle as O
  T, g, s, per_env = _golden(name)
  n, steps, ep_len = g['q0'].shape[0], g['act'].shape[0], int(g['episode_length'])
  fn, _ = T.randomization_fn_from_golden(g)
  env = training.wrap(envs.create(name), episode_length=ep_len, randomization_fn=fn)
  assert env.batch_size == n and env.systems is not None
  dev = env.device
  model = env._model()
  assert model.num_models == n
  bufs, obs0 = model.env_reset(env.spec, torch.as_tensor(g['q0'].astype(f32), device=dev), torch.as_tensor(g['qd0'].astype(f32), device=dev))
  np.testing.assert_allclose(obs0.cpu().numpy(), g['obs0'], rtol=1e-4, atol=1e-5)
  first = {'first_pipeline_state': PS.from_flat(bufs), 'first_obs': obs0}
  shapes = {f: tuple(v.shape) for f, v in bufs.items()}
  inside = []
  for k in range(steps):
    ps = T.ps_at(g, k, slice(0, n), f32, shapes)
    prev = 'obs0' if k == 0 else f'step{k - 1}_obs'
    done = np.zeros(n, f32) if k == 0 else g[f'step{k - 1}_done'].astype(f32)
    nsteps = np.zeros(n, f32) if k == 0 else g[f'step{k - 1}_steps'].astype(f32)
    info = {'steps': torch.as_tensor(nsteps, device=dev), 'truncation': torch.zeros(n, device=dev)}
    info.update(first)
    s_in = State(PS.from_flat({f: torch.as_tensor(ps[f], device=dev) for f in O.STATE_FIELDS}), torch.as_tensor(g[prev].astype(f32), device=dev),
                 torch.zeros(n, device=dev), torch.as_tensor(done, device=dev), {}, info)
    out = env.step(s_in, torch.as_tensor(g['act'][k].astype(f32), device=dev))
    p = f'step{k}_'
    np.testing.assert_array_equal(out.done.cpu().numpy(), g[p + 'done'])
    np.testing.assert_array_equal(out.info['steps'].cpu().numpy(), g[p + 'steps'])
    np.testing.assert_array_equal(out.info['truncation'].cpu().numpy(), g[p + 'truncation'])
    gq, gqd = out.pipeline_state.q.cpu().numpy(), out.pipeline_state.qd.cpu().numpy()
    e = np.maximum((np.abs(gq - g[p + 'q']) / (1e-5 + 1e-4 * np.abs(g[p + 'q']))).max(1), (np.abs(gqd - g[p + 'qd']) / (1e-5 + 1e-4 * np.abs(g[p + 'qd']))).max(1))
    inside.append(e <= 1.0)
    np.testing.assert_allclose(gq, g[p + 'q'], rtol=2e-3, atol=2e-4)
    np.testing.assert_allclose(out.obs.cpu().numpy(), g[p + 'obs'], rtol=5e-3, atol=5e-3)
    np.testing.assert_allclose(out.reward.cpu().numpy(), g[p + 'reward'], rtol=2e-3, atol=5e-3)
  # float32 against the float64 reference run: ~0.87 of Humanoid env-steps take the reference's solver branches
  # (tests/test_reference_golden_big.py); 24 env-steps here
  assert np.mean(inside) >= 0.75, np.mean(inside)


def test_ppo_trains_with_domain_randomization():
  """The reference's testPPOWithDomainRandomization (agents/ppo/train_test.py:226-258): a short PPO run on an env whose
  links' centre-of-mass offsets are randomised per env."""
  from brax_b200 import base
  from brax_b200.training import ppo

  def rand_fn(sys, rng):
    n = len(rng)
    off = np.stack([np.random.default_rng(int(r)).uniform(-0.1, 0.1, 3) for r in rng]).astype(f32)
    pos = np.repeat(np.asarray(sys.link.inertia.transform.pos, f32)[None], n, 0)
    pos[:, 0] += off
    sys_v = sys.tree_replace({'link.inertia.transform.pos': pos})
    in_axes = base.tree_map(lambda x: None, sys).tree_replace({'link.inertia.transform.pos': 0})
    return sys_v, in_axes

  agent, metrics = ppo.train('ant', num_envs=64, episode_length=100, num_timesteps=2 ** 13, unroll_length=5, batch_size=64,
                             num_minibatches=8, num_update_epochs=4, normalize_advantage=False, randomization_fn=rand_fn, seed=2)
  assert metrics['sps'] > 0 and np.isfinite(metrics.get('loss', 0.0))
