"""Reference-sized rollouts from the reference's own source: 32 envs x 50 env-steps (250 physics
substeps per env) of Ant, Humanoid and Humanoid-falls, with the reference run's BRANCH DECISIONS.

tests/golden/ref_big_<model>.npz (tools/gen_reference_golden_big.py) hold, per physics substep of the
reference run, the projected-gradient iteration and line-search counts, the Newton-Schulz accepted
candidates and cold starts, and the contact active set -- observed from the reference's own
`jp.where` selections (math.py:297,302) and objective calls -- plus q, qd after every env-step and
mass_mx_inv around five checkpoint env-steps.

CPU side (this file): the oracle (float64) and the kernel source in double precision (host
emulator) must reproduce EVERY branch decision of the whole free-running rollout exactly and track
q, qd; at the checkpoints they are exact one-env-step maps from the reference's state.  The float32
builds are then measured against the same reference run (how often float32 takes another branch).
GPU side: tests/test_gpu_reference_big.py.
"""
import os

import numpy as np
import pytest

from brax_b200 import envs_assets, native
from oracle import oracle as O
from tests.conftest import ROOT
from tests.simt import sim as S

MODELS = ['ant', 'humanoid', 'humanoid_falls']


def load(name):
  g = np.load(os.path.join(ROOT, 'tests', 'golden', f'ref_big_{name}.npz'))
  return envs_assets.load('humanoid' if name == 'humanoid_falls' else name), g


def state_at(g, k):
  """(q, qd) after k env-steps of the reference run (k = 0: the reset state)."""
  return (g['q0'], g['qd0']) if k == 0 else (g['q'][k - 1], g['qd'][k - 1])


def checkpoint_state(stepper_init, g, i):
  """The reference's full State before env-step ck_steps[i]: every leaf but mass_mx_inv is a pure function of
  (q, qd) (pipeline.py:51-61,78-94), mass_mx_inv carries the Newton-Schulz history and comes from the golden file."""
  k = int(g['ck_steps'][i])
  q, qd = state_at(g, k)
  st = stepper_init(q, qd)
  st['mass_mx_inv'][...] = g['ck_minv'][i, 0].astype(st['mass_mx_inv'].dtype)
  return k, st


@pytest.mark.parametrize('name', MODELS)
def test_golden_files_exercise_the_branches(name):
  s, g = load(name)
  st = g['stats'].astype(np.int64)
  T, F, E, _ = st.shape
  assert (T, E) == (50, 32) and F == 5
  assert np.isfinite(g['q']).all() and np.isfinite(g['qd']).all()
  assert st[..., 1].max() > st[..., 0].max()            # the line search backtracks
  assert len(np.unique(st[..., 2])) > 1                 # Newton-Schulz accepts vary
  if name != 'ant':
    assert st[..., 3].sum() > 0                         # cold starts happen
  assert g['con_active'].any() and not g['con_active'].all()


@pytest.mark.parametrize('name', MODELS)
def test_float64_oracle_reproduces_every_branch_of_the_reference_rollout(name):
  """Free-running 50 env-steps x 32 envs: all 8000 substeps take the reference run's branches, q / qd track."""
  s, g = load(name)
  o = O.Oracle(s, np.float64)
  st = o.init(g['q0'], g['qd0'])
  T, F = g['stats'].shape[:2]
  for t in range(T):
    for f in range(F):
      prev = st['stats'].copy()
      o.step(st, g['act'][t], 1)
      np.testing.assert_array_equal(st['stats'] - prev, g['stats'][t, f], err_msg=f'{name} env-step {t} substep {f}: branch statistics')
      np.testing.assert_array_equal(st['con_dist'] < 0, g['con_active'][t, f], err_msg=f'{name} env-step {t} substep {f}: contact active set')
    # two float64 evaluation orders drift apart at the system's own (chaotic) rate: 1e-16 -> 1e-12 in ten env-steps,
    # 1e-5 by env-step 50 for the most agitated Ant, with no branch ever differing
    tol = 1e-9 if t < 10 else 1e-3
    np.testing.assert_allclose(st['q'], g['q'][t], rtol=tol, atol=tol, err_msg=f'{name} q after env-step {t}')
    np.testing.assert_allclose(st['qd'], g['qd'][t], rtol=10 * tol, atol=10 * tol, err_msg=f'{name} qd after env-step {t}')


@pytest.mark.parametrize('name', MODELS)
def test_float64_kernel_source_reproduces_every_branch_at_the_checkpoints(name):
  """The kernel source in double precision (host emulator, the variant the library picks): exact one-env-step maps
  from the reference's state at the five checkpoints, substep by substep, branch statistics equal."""
  s, g = load(name)
  sim = S.Sim(s, dtype=np.float64)
  for i in range(len(g['ck_steps'])):
    k, st = checkpoint_state(lambda q, qd: sim.init(q, qd), g, i)
    for f in range(g['stats'].shape[1]):
      st = sim.step(st, g['act'][k], 1, diag=True)
      np.testing.assert_array_equal(st['stats'], g['stats'][k, f], err_msg=f'{name} checkpoint {k} substep {f}')
      np.testing.assert_array_equal(st['con_dist'] < 0, g['con_active'][k, f])
    for leaf, ref in (('q', g['q'][k]), ('qd', g['qd'][k]), ('mass_mx_inv', g['ck_minv'][i, 1])):
      scale = max(1.0, float(np.abs(ref).max()))
      assert np.abs(st[leaf] - ref).max() <= 1e-9 * scale, (name, k, leaf, np.abs(st[leaf] - ref).max())


# float32 measured against the reference run.  The Newton-Schulz ACCEPT count is precision dependent by construction
# (in float64 a warm-started iteration reaches its residual floor ~1e-16 after about three steps and every later
# `err_next < err` is false, math.py:297; in float32 the floor is ~1e-6 and the comparison keeps flipping on rounding
# noise), so the float32 branch signature is (PG iterations, line-search trials, cold start, contact active set).
SIG = (0, 1, 3)
# Measured here, float32 oracle | float32 kernel source, one env-step (five substeps) from the reference's state:
#   same-branch env-steps      Ant 0.74 | 0.76   Humanoid 0.87 | 0.88   Humanoid-falls 0.87 | 0.88
#   of those inside 1e-4/1e-5  Ant 0.98 | 0.98   Humanoid 0.99 | 0.99   Humanoid-falls 0.96 | 0.96   (p50 0.02-0.08 of the tolerance)
# The tail is the reference algorithm's own sensitivity: the float64 oracle fed the float32-ROUNDED checkpoint state
# leaves the tolerance by 8x for the worst Ant env (test below), without any branch changing.  So float32 results
# cannot be held to a max; that is what the double-precision instantiation above (exact) and the bit-for-bit
# CUDA == float32-emulator tests (tests/test_gpu_bitexact.py) are for.  The gates below sit just under the measured values.
F32_GATES = {'ant': (0.70, 0.95), 'humanoid': (0.82, 0.96), 'humanoid_falls': (0.82, 0.93)}


def f32_one_env_step_maps(name, stepper):
  """stepper(st, act) -> (st, per-substep stats, con_dist).  Returns (same-branch mask, scaled error) per checkpoint env-step."""
  s, g = load(name)
  out = []
  for i in range(len(g['ck_steps'])):
    k, st = checkpoint_state(stepper.init, g, i)
    ok = np.ones(g['q0'].shape[0], bool)
    act = g['act'][k].astype(np.float32)
    for f in range(g['stats'].shape[1]):
      st, stats, dist = stepper.substep(st, act)
      ok &= (stats[:, SIG] == g['stats'][k, f][:, SIG]).all(1)
      ok &= ((dist < 0) == g['con_active'][k, f]).all(1)
    ref_q, ref_qd = g['q'][k], g['qd'][k]
    err = np.maximum((np.abs(st['q'] - ref_q) / (1e-5 + 1e-4 * np.abs(ref_q))).max(1), (np.abs(st['qd'] - ref_qd) / (1e-5 + 1e-4 * np.abs(ref_qd))).max(1))
    out.append((ok, err))
  return np.concatenate([o for o, _ in out]), np.concatenate([e for _, e in out])


class OracleStepper:
  def __init__(self, s):
    self.o = O.Oracle(s, np.float32)

  def init(self, q, qd):
    return self.o.init(q.astype(np.float32), qd.astype(np.float32))

  def substep(self, st, act):
    prev = st['stats'].copy()
    self.o.step(st, act, 1)
    return st, st['stats'] - prev, st['con_dist']


class SimStepper:
  def __init__(self, s):
    self.sim = S.Sim(s, dtype=np.float32)

  def init(self, q, qd):
    return self.sim.init(q, qd)

  def substep(self, st, act):
    st = self.sim.step(st, act, 1, diag=True)
    return st, st['stats'], st['con_dist']


@pytest.mark.parametrize('name', MODELS)
@pytest.mark.parametrize('which', ['oracle', 'kernel_source'])
def test_float32_builds_against_the_reference_run(name, which):
  """float32 oracle and float32 kernel source (emulator) as one-env-step maps from the reference's checkpoints."""
  s, _ = load(name)
  ok, err = f32_one_env_step_maps(name, OracleStepper(s) if which == 'oracle' else SimStepper(s))
  frac_gate, inside_gate = F32_GATES[name]
  assert ok.mean() >= frac_gate, (name, which, ok.mean())
  assert (err[ok] <= 1.0).mean() >= inside_gate, (name, which, (err[ok] <= 1.0).mean())
  assert np.median(err[ok]) <= 0.15 and np.percentile(err[ok], 90) <= 0.5, (name, which, np.percentile(err[ok], [50, 90]))


def test_the_tail_is_the_reference_algorithms_own_sensitivity():
  """The float64 oracle on the float32-ROUNDED reference state (a 6e-8 relative input perturbation, nothing else)
  already leaves the 1e-4 / 1e-5 tolerance after one env-step for the most agitated Ant env, with every branch equal."""
  s, g = load('ant')
  o = O.Oracle(s, np.float64)
  r32 = lambda a: a.astype(np.float32).astype(np.float64)   # noqa: E731
  worst = 0.0
  for i in range(len(g['ck_steps'])):
    k = int(g['ck_steps'][i]); q, qd = state_at(g, k)
    st = o.init(r32(q), r32(qd)); st['mass_mx_inv'][...] = r32(g['ck_minv'][i, 0])
    same = np.ones(q.shape[0], bool)
    for f in range(g['stats'].shape[1]):
      prev = st['stats'].copy()
      o.step(st, r32(g['act'][k]), 1)
      same &= ((st['stats'] - prev)[:, SIG] == g['stats'][k, f][:, SIG]).all(1)
    err = np.maximum((np.abs(st['q'] - g['q'][k]) / (1e-5 + 1e-4 * np.abs(g['q'][k]))).max(1),
                     (np.abs(st['qd'] - g['qd'][k]) / (1e-5 + 1e-4 * np.abs(g['qd'][k]))).max(1))
    worst = max(worst, float(err[same].max()))
  assert worst > 1.0, worst
