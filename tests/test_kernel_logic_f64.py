"""The kernel's algorithm source in DOUBLE precision against the reference-source goldens.

`brax_b200/csrc/bxg_core.cuh` is written against a scalar type `real` (float in the product).
tests/simt/ instantiates the same source with real = double and runs the lanes of a group as a
loop.  tests/golden/ref_<model>.npz hold every `generalized.State` leaf after `init` and after
every `step`, produced by the reference's own source in float64 (tools/gen_reference_golden.py).
So the LOGIC of every kernel variant (indexing, tree tables, phase order, slab aliasing, the
Newton-Schulz accept / reject / cold-start state machine of `math.py:292-302`, active-row
compaction, the solver's line search) is held to 1e-9 on EVERY leaf of EVERY env and step:
no percentile, no float32 noise floor.  What float32 adds on top is rounding only, and
tests/test_gpu_parity.py bounds that on the device.

Matches the intent of `brax/generalized/pipeline_test.py:37-47` (the step agrees with the reference
engine on every state field), with the reference's own code as the other side.
"""
import os

import numpy as np
import pytest

from brax_b200 import native
from tests.conftest import ROOT
from tests.simt import sim as S
from tests.test_reference_golden import MODELS, _load

# variants each golden model is run on: the one the library picks (-1) plus every other compiled
# variant the model fits (bxg_model.h variant()); 3 is the generic any-size kernel
FITS = {
    'ant': (-1, 1, 2, 3, 4, 5, 6), 'humanoid': (-1, 2, 3, 4, 6), 'halfcheetah': (-1, 3), 'hopper': (-1, 0, 1, 3, 5),
    'walker2d': (-1, 3), 'humanoidstandup': (-1, 3), 'pusher': (-1, 3), 'triple_pendulum_motor': (-1, 0, 1, 3, 8, 9),
    'inverted_pendulum': (-1, 0, 3, 8), 'inverted_double_pendulum': (-1, 0, 3, 8), 'reacher': (-1, 0, 3),
    'swimmer': (-1, 3), 'two_trees': (-1, 1, 3),
}


def _close(a, b, what, tol=1e-9):
  b = np.asarray(b, np.float64).reshape(a.shape)
  scale = max(1.0, float(np.abs(b).max())) if b.size else 1.0
  err = float(np.abs(a - b).max()) if b.size else 0.0
  assert np.isfinite(a).all(), f'{what}: non-finite'
  assert err <= tol * scale, f'{what}: max abs error {err:.3e} (scale {scale:.3g})'


def _variants(name):
  return [pytest.param(name, v, id=f'{name}-v{"auto" if v < 0 else v}') for v in FITS[name]]


ALL = [p for n in MODELS for p in _variants(n)]


@pytest.mark.parametrize('name,variant', ALL)
def test_init_every_leaf_1e9(name, variant):
  s, g = _load(name)
  sim = S.Sim(s, variant=variant, dtype=np.float64)
  st = sim.init(g['q0'], g['qd0'])
  for f in native.STATE_FIELDS:
    _close(st[f], g[f'init_{f}'], f'{name} v{variant} init {f}')


@pytest.mark.parametrize('name,variant', ALL)
def test_every_step_every_leaf_1e9(name, variant):
  """One-step maps from the REFERENCE's state: leaf-exact to 1e-9 for every env and step, and the same
  again with the lanes of every phase run in reverse order (an intra-phase dependency would differ)."""
  s, g = _load(name)
  steps = g['act'].shape[0]
  for reverse in (False, True):
    sim = S.Sim(s, variant=variant, dtype=np.float64, reverse=reverse)
    for k in range(steps):
      prev = 'init' if k == 0 else f'step{k - 1}'
      shapes = sim.shapes
      n = g['q0'].shape[0]
      st = {f: np.ascontiguousarray(g[f'{prev}_{f}'].reshape((n,) + shapes[f]), np.float64) for f in native.STATE_FIELDS}
      out = sim.step(st, g['act'][k], 1)
      for f in native.STATE_FIELDS:
        _close(out[f], g[f'step{k}_{f}'], f'{name} v{variant} rev={reverse} step {k} {f}', tol=1e-8)


@pytest.mark.parametrize('name', ['ant', 'humanoid', 'hopper'])
def test_cholesky_mode_differs_only_through_minv(name):
  """BXG_MINV_CHOLESKY is a stated deviation: everything the step computes
  BEFORE mass_mx_inv is used again is unchanged, and M Minv = I to double precision."""
  s, g = _load(name)
  sim = S.Sim(s, dtype=np.float64, minv_mode=native.MINV_CHOLESKY)
  n = g['q0'].shape[0]
  st = {f: np.ascontiguousarray(g[f'init_{f}'].reshape((n,) + sim.shapes[f]), np.float64) for f in native.STATE_FIELDS}
  out = sim.step(st, g['act'][0], 1)
  for f in ('q', 'qd', 'x_pos', 'x_rot', 'xd_ang', 'xd_vel', 'mass_mx', 'con_jac', 'qf_smooth', 'qf_constraint', 'qdd'):
    _close(out[f], g[f'step0_{f}'], f'{name} cholesky step0 {f}', tol=1e-8)
  eye = np.einsum('eij,ejk->eik', out['mass_mx'], out['mass_mx_inv'])
  assert np.abs(eye - np.eye(s.nv)).max() < 1e-9


def test_double_build_is_the_same_source():
  """The two emulator libraries are one translation unit compiled twice; the float one is what the
  CPU suite compares bit-for-bit with the float oracle (tests/test_kernel_logic_simt.py)."""
  S.build()
  import ctypes
  assert ctypes.CDLL(S._SO).sim_sizeof_real() == 4 and ctypes.CDLL(S._SO64).sim_sizeof_real() == 8
  src = open(os.path.join(ROOT, 'tests', 'simt', 'bxg_sim.cpp')).read()
  assert 'bxg_core.cuh' in src and 'BXG_SIM_F64' in src
