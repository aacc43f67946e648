"""The exact checks have teeth: deliberately broken copies of the kernel source FAIL them.

Each mutation below is a small, plausible bug in a rarely taken branch of brax_b200/csrc/bxg_core.cuh (or a one-ulp
class perturbation).  The mutated source is compiled as the double-precision host emulator and run through the same
comparison tests/test_kernel_logic_f64.py makes against the reference-source goldens; every mutation must push some
State leaf outside the 1e-9 gate (the unmutated source sits at 1e-13).  A percentile-based float32 comparison would
pass most of them: that is why the double-precision instantiation exists."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np
import pytest

from brax_b200 import native
from tests.conftest import ROOT
from tests.simt import sim as S
from tests.test_reference_golden import _load

CORE = os.path.join(ROOT, 'brax_b200', 'csrc', 'bxg_core.cuh')

MUTATIONS = {
    # Newton-Schulz: after a REJECTED candidate the next I + r' must replace the loser (math.py:292-302)
    'newton_schulz_reject_path': ("real* dst = accept ? Xc : Q;", "real* dst = accept ? Q : Xc;", ('ant', 'hopper')),
    # jac_limit: the upper joint limit is ignored (constraint.py:112-118)
    'upper_joint_limit_ignored': ("real pos = r_min(r_min(pos_min, pos_max), R(0.));", "real pos = r_min(pos_min, R(0.));", ('ant', 'humanoid', 'walker2d')),
    # a one-ulp-class error in the Newton-Schulz cold start 0.5 M^T / tr(M M^T) (math.py:302): 0.5 -> 0.5 (1 + 2^-23)
    'cold_start_off_by_one_float_ulp': ("real x[4] = {R(0.5) * m.x, R(0.5) * m.y, R(0.5) * m.z, R(0.5) * m.w}, q[4];",
                                        "real x[4] = {R(0.50000006) * m.x, R(0.50000006) * m.y, R(0.50000006) * m.z, R(0.50000006) * m.w}, q[4];", ('humanoid',)),
    # point_jacobian: the ancestor mask of a contact shifted by one dof (constraint.py:89-95)
    'contact_ancestor_mask_shifted': ("uint32_t bit = d < 32 ? (lo >> d) & 1u : (hi >> (d - 32)) & 1u;",
                                      "uint32_t bit = d < 31 ? (lo >> (d + 1)) & 1u : (hi >> (d - 31)) & 1u;", ('ant', 'humanoid')),
    # _imp_aref: the upper branch of the impedance curve (x >= mid) is never taken (constraint.py:44-49)
    'impedance_upper_branch_dropped': ("real imp_y = imp_x < mid ? imp_a : imp_b;", "real imp_y = imp_a;", ('ant', 'humanoid', 'hopper')),
}


def _worst_error(lib_path, name):
  """max over init + every step + every leaf of |emulator - reference| / scale, as in tests/test_kernel_logic_f64.py"""
  s, g = _load(name)
  sim = S.Sim(s, dtype=np.float64)
  sim.lib = ctypes.CDLL(lib_path)
  n, steps, worst = g['q0'].shape[0], g['act'].shape[0], 0.0

  def upd(out, prefix):
    nonlocal worst
    for f in native.STATE_FIELDS:
      ref = np.asarray(g[f'{prefix}_{f}'], np.float64).reshape(out[f].shape)
      if ref.size:
        err = np.abs(out[f] - ref).max() / max(1.0, np.abs(ref).max())
        worst = max(worst, float(err) if np.isfinite(err) else np.inf)
  upd(sim.init(g['q0'], g['qd0']), 'init')
  for k in range(steps):
    prev = 'init' if k == 0 else f'step{k - 1}'
    st = {f: np.ascontiguousarray(g[f'{prev}_{f}'].reshape((n,) + sim.shapes[f]), np.float64) for f in native.STATE_FIELDS}
    upd(sim.step(st, g['act'][k], 1), f'step{k}')
  return worst


def _build(src_text, tmp):
  csrc = os.path.join(tmp, 'brax_b200', 'csrc')
  os.makedirs(csrc)
  os.makedirs(os.path.join(tmp, 'include'))
  os.makedirs(os.path.join(tmp, 'tests', 'simt'))
  for f in ('bxg_model.h',):
    open(os.path.join(csrc, f), 'w').write(open(os.path.join(ROOT, 'brax_b200', 'csrc', f)).read())
  open(os.path.join(csrc, 'bxg_core.cuh'), 'w').write(src_text)
  open(os.path.join(tmp, 'include', 'bxg.h'), 'w').write(open(os.path.join(ROOT, 'include', 'bxg.h')).read())
  open(os.path.join(tmp, 'tests', 'simt', 'bxg_sim.cpp'), 'w').write(open(os.path.join(ROOT, 'tests', 'simt', 'bxg_sim.cpp')).read())
  so = os.path.join(tmp, 'sim_f64.so')
  subprocess.run(['g++', '-O1', '-ffp-contract=off', '-fPIC', '-shared', '-std=c++17', '-Wno-unknown-pragmas', '-DBXG_REAL=double', '-DBXG_SIM_F64',
                  os.path.join(tmp, 'tests', 'simt', 'bxg_sim.cpp'), '-o', so], check=True)
  return so


def test_the_unmutated_source_passes_with_four_orders_of_margin():
  with tempfile.TemporaryDirectory() as tmp:
    so = _build(open(CORE).read(), tmp)
    for name in ('ant', 'humanoid', 'hopper', 'walker2d'):
      assert _worst_error(so, name) < 1e-12, name


@pytest.mark.parametrize('mutation', sorted(MUTATIONS))
def test_mutation_fails_the_double_precision_check(mutation):
  old, new, models = MUTATIONS[mutation]
  src = open(CORE).read()
  assert src.count(old) == 1, f'mutation site of {mutation} not found exactly once'
  with tempfile.TemporaryDirectory() as tmp:
    so = _build(src.replace(old, new), tmp)
    worst = max(_worst_error(so, name) for name in models)
  assert worst > 1e-9, f'{mutation}: the broken kernel source still passes (worst error {worst:.2e})'
