"""The CUDA path against the SAME SOURCE run on the host: bit for bit.

`brax_b200/csrc/bxg_core.cuh` fixes every rounding in the source: fused multiply-adds are
written out (`r_fma`, `__ffma2_rn`), the device build has no implicit contraction
(`-fmad=false`), division and square root are IEEE on both sides, sine / cosine come from one
polynomial (`r_sincos`), and group reductions use the butterfly order of the shuffles.  The host
emulator (tests/simt/, float build, `-ffp-contract=off`) therefore reproduces what the GPU
computes EXACTLY: every State leaf, the solver / Newton-Schulz branch counters, the contact
distances, for every kernel variant, `torch.equal`-style, no tolerance and no percentile.

Together with tests/test_kernel_logic_f64.py (the same source in double precision equals the
reference-source goldens to 1e-9 on every leaf) this leaves float32 rounding as the only
difference between the CUDA path and the reference's algorithm.  A logic bug in a rarely taken
branch (a Newton-Schulz reject, a limit row at exactly zero) cannot hide in either test.
(`powf` is the one libm call left: only models with a solimp power other than 2 reach it; none here.)
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

MODELS = ['ant', 'humanoid', 'humanoid_falls', 'hopper', 'halfcheetah', 'walker2d', 'humanoidstandup', 'pusher', 'swimmer',
          'inverted_pendulum', 'inverted_double_pendulum', 'reacher']


def _inputs(name, n, seed):
  from brax_b200 import envs_assets, workloads
  if name in ('ant', 'humanoid', 'humanoid_falls'):
    s, q, qd = workloads.reset(name, 0, n, seed, 'cpu')
    q, qd = q.numpy(), qd.numpy()
    if name == 'ant':
      q[: n // 2, 2] = 0.45 + 0.1 * np.random.default_rng(seed).uniform(size=n // 2).astype(np.float32)   # feet on the floor
    if name == 'humanoid':
      q[: n // 2, 2] = 1.29 + 0.02 * np.random.default_rng(seed).uniform(size=n // 2).astype(np.float32)
    acts = [workloads.action(name, 0, n, seed, k, 'cpu').numpy() for k in range(4)]
    return s, q, qd, acts, workloads.N_FRAMES[name]
  s = envs_assets.load(name)
  rng = np.random.default_rng(seed)
  q = (np.asarray(s.init_q, np.float32)[None] + rng.uniform(-0.1, 0.1, (n, s.nq))).astype(np.float32)
  if name in ('hopper', 'walker2d', 'halfcheetah'):
    q[:, 1] -= {'hopper': 0.05, 'walker2d': 0.1, 'halfcheetah': 0.45}[name] * rng.uniform(0.5, 1.0, n).astype(np.float32)
  qd = (0.1 * rng.standard_normal((n, s.nv))).astype(np.float32)
  acts = [rng.uniform(-1, 1, (n, s.nu)).astype(np.float32) for _ in range(4)]
  return s, q, qd, acts, 4


def _assert_same_bits(got, ref, what):
  """Bitwise equality (so that -0.0 / +0.0 and NaN payloads count too), with a readable report."""
  g = np.ascontiguousarray(got); r = np.ascontiguousarray(ref).reshape(g.shape)
  if g.dtype == np.float32:
    same = g.view(np.uint32) == r.astype(np.float32).view(np.uint32)
  else:
    same = g == r
  if not same.all():
    bad = np.argwhere(~same)
    i = tuple(bad[0])
    raise AssertionError(f'{what}: {len(bad)} of {g.size} elements differ; first at {i}: gpu {g[i]!r} vs host {r[i]!r} '
                         f'(max abs diff {np.abs(g.astype(np.float64) - r.astype(np.float64)).max():.3e})')


@pytest.mark.parametrize('minv', ['newton_schulz', 'cholesky'])
@pytest.mark.parametrize('name', MODELS)
def test_cuda_equals_host_emulation_of_the_same_source_bit_for_bit(name, minv):
  import torch
  from brax_b200 import native
  from tests.simt import sim as S
  if minv == 'cholesky' and name not in ('ant', 'humanoid', 'hopper', 'inverted_pendulum'):
    pytest.skip('exact-inverse mode: one model per lane-group width')
  mode = native.MINV_NEWTON_SCHULZ if minv == 'newton_schulz' else native.MINV_CHOLESKY
  n = 40                      # not a multiple of any CTA size: the last pass is ragged
  s, q, qd, acts, nf = _inputs(name, n, seed=3)
  dev = torch.device('cuda', 0)
  nm = native.NativeModel(s, 0, mode)
  sim = S.Sim(s, minv_mode=mode)
  assert native.plan(s, mode)['variant'] >= 0
  g = nm.init(torch.as_tensor(q, device=dev), torch.as_tensor(qd, device=dev))
  h = sim.init(q, qd)
  for f in native.STATE_FIELDS:
    _assert_same_bits(g[f].cpu().numpy(), h[f], f'{name} init {f}')
  active = 0
  for k, act in enumerate(acts):
    # one env-step (n_frames substeps in ONE launch) and, from the same state, the substeps one launch each
    diag = nm.alloc_diag(n)
    g2 = nm.step(g, torch.as_tensor(act, device=dev), nf, diag=diag)
    h2 = sim.step(h, act, nf, diag=True)
    for f in native.STATE_FIELDS:
      _assert_same_bits(g2[f].cpu().numpy(), h2[f], f'{name} env-step {k} {f}')
    _assert_same_bits(diag['stats'].cpu().numpy(), h2['stats'], f'{name} env-step {k} branch counters')
    if nm.ncon:
      _assert_same_bits(diag['con_dist'].cpu().numpy(), h2['con_dist'], f'{name} env-step {k} contact distances')
      active += int((h2['con_dist'] < 0).sum())
    g, h = g2, {f: h2[f] for f in native.STATE_FIELDS}
  if name in ('ant', 'humanoid', 'humanoid_falls', 'hopper', 'halfcheetah', 'walker2d', 'humanoidstandup'):
    assert active > 0      # contacts were exercised


@pytest.mark.parametrize('name', ['ant', 'humanoid', 'hopper', 'inverted_double_pendulum', 'reacher', 'swimmer', 'pusher'])
def test_fused_env_step_equals_host_emulation_bit_for_bit(name):
  """bxg_env_reset / bxg_env_step (physics + obs, reward, done, metrics, Episode / AutoReset arithmetic)."""
  import torch
  from brax_b200 import envs
  from tests.simt import sim as S
  n = 24
  env = envs.create(name, episode_length=3, auto_reset=True, batch_size=n)
  st = env.reset(5)
  sim = S.SimEnv(env)
  q, qd = st.pipeline_state.q.cpu().numpy(), st.pipeline_state.qd.cpu().numpy()
  hs, hobs = sim.reset(q, qd)
  _assert_same_bits(st.obs.cpu().numpy(), hobs, f'{name} reset obs')
  first, first_obs = {f: v.copy() for f, v in hs.items()}, hobs.copy()
  done, steps = np.zeros(n, np.float32), np.zeros(n, np.float32)
  gen = np.random.default_rng(0)
  saw_done = False
  for k in range(5):       # crosses the episode end: truncation and auto-reset happen
    act = gen.uniform(-1, 1, (n, env.action_size)).astype(np.float32)
    st = env.step(st, torch.as_tensor(act, device=st.obs.device))
    hs, io = sim.step(hs, act, done, steps, first=first, first_obs=first_obs)
    done, steps = io['done'], io['steps']
    saw_done |= bool(done.any())
    _assert_same_bits(st.obs.cpu().numpy(), io['obs'], f'{name} step {k} obs')
    _assert_same_bits(st.reward.cpu().numpy(), io['reward'], f'{name} step {k} reward')
    _assert_same_bits(st.done.cpu().numpy(), io['done'], f'{name} step {k} done')
    _assert_same_bits(st.info['steps'].cpu().numpy(), io['steps'], f'{name} step {k} steps')
    _assert_same_bits(st.info['truncation'].cpu().numpy(), io['truncation'], f'{name} step {k} truncation')
    _assert_same_bits(st.pipeline_state.q.cpu().numpy(), hs['q'], f'{name} step {k} q')
    _assert_same_bits(st.pipeline_state.mass_mx_inv.cpu().numpy(), hs['mass_mx_inv'], f'{name} step {k} mass_mx_inv')
  assert saw_done


@pytest.mark.parametrize('name', ['ant', 'humanoid', 'hopper', 'swimmer', 'pusher'])
def test_lean_state_io_gives_the_same_bits(name):
  """BXG_STEP_LEAN (include/bxg.h): only q, qd, x and mass_mx_inv are read, the derived leaves are recomputed on
  chip; q, qd, x, xd, mass_mx_inv come out bit-identical to the default step, the other leaves are never touched."""
  import torch
  from brax_b200 import native
  n = 70
  s, q, qd, acts, nf = _inputs(name, n, seed=9)
  dev = torch.device('cuda', 0)
  nm = native.NativeModel(s, 0)
  full = nm.init(torch.as_tensor(q, device=dev), torch.as_tensor(qd, device=dev))
  lean = {k: full[k].clone() for k in native.LEAN_FIELDS}
  for act in acts:
    a = torch.as_tensor(act, device=dev)
    full = nm.step(full, a, nf)
    poisoned = {k: torch.full_like(v, float('nan')) for k, v in nm.alloc(n).items() if k not in native.LEAN_FIELDS}
    out = dict(poisoned, **nm.alloc(n, lean=True))
    nm.step(dict(poisoned, **lean), a, nf, out=out, lean=True)      # derived input leaves are garbage: never read
    for k in native.LEAN_FIELDS:
      assert torch.equal(out[k], full[k]), (name, k)
    assert all(torch.isnan(out[k]).all() for k in poisoned)           # and never written
    lean = {k: out[k] for k in native.LEAN_FIELDS}


@pytest.mark.parametrize('name', ['ant', 'humanoid'])
def test_lean_fused_env_gives_the_same_bits(name):
  import torch
  from brax_b200 import envs
  n = 96
  e_full = envs.create(name, episode_length=4, auto_reset=True, batch_size=n)
  e_lean = envs.create(name, episode_length=4, auto_reset=True, batch_size=n, lean=True)
  a, b = e_full.reset(2), e_lean.reset(2)
  assert b.pipeline_state.cdof.ang is None and torch.equal(a.obs, b.obs)
  gen = torch.Generator(device='cpu').manual_seed(0)
  for k in range(7):
    act = (torch.rand((n, e_full.action_size), generator=gen) * 2 - 1).to(a.obs.device)
    a, b = e_full.step(a, act), e_lean.step(b, act)
    for x, y in ((a.obs, b.obs), (a.reward, b.reward), (a.done, b.done), (a.pipeline_state.q, b.pipeline_state.q),
                 (a.pipeline_state.x.rot, b.pipeline_state.x.rot), (a.info['steps'], b.info['steps'])):
      assert torch.equal(x, y), (name, k)


@pytest.mark.parametrize('name', ['ant', 'humanoid'])
def test_model_specialised_kernels_equal_the_generic_ones_bit_for_bit(name):
  """Kernel ids 10 / 11 (constexpr Dims, bxg_inst.cu) are what Ant / Humanoid run on; BXG_NO_SPECIALISE keeps the
  blob-driven kernel of the variant.  Same source, same arithmetic: identical bits."""
  import os
  import torch
  from brax_b200 import native
  n = 50
  s, q, qd, acts, nf = _inputs(name, n, seed=4)
  assert native.plan(s)['kernel_id'] == {'ant': 10, 'humanoid': 11}[name]
  dev = torch.device('cuda', 0)
  spec = native.NativeModel(s, 0)
  os.environ['BXG_NO_SPECIALISE'] = '1'
  try:
    gen = native.NativeModel(s, 0)
  finally:
    del os.environ['BXG_NO_SPECIALISE']
  a = spec.init(torch.as_tensor(q, device=dev), torch.as_tensor(qd, device=dev))
  b = gen.init(torch.as_tensor(q, device=dev), torch.as_tensor(qd, device=dev))
  for act in acts:
    t = torch.as_tensor(act, device=dev)
    a, b = spec.step(a, t, nf), gen.step(b, t, nf)
    for f in native.STATE_FIELDS:
      assert torch.equal(a[f], b[f]), (name, f)
