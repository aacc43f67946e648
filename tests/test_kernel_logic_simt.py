"""Kernel LOGIC on the CPU: brax_b200/csrc/bxg_core.cuh (the source the CUDA
kernel is compiled from) run through the host lane-group emulator in
tests/simt/ and compared with the float32 oracle.  The kernel source writes its fused
multiply-adds out (the device build has no implicit contraction, so the emulator and
the GPU agree bit for bit: tests/test_gpu_bitexact.py); the oracle is unfused C, so the
two float32 results agree to rounding.  The exact check of the kernel's logic is the
double-precision instantiation against the reference goldens (tests/test_kernel_logic_f64.py)."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.conftest import golden
from tests.simt.sim import Sim


def _inputs(sys_, model, n, seed=0):
  from brax_b200 import workloads
  _, q, qd = workloads.reset(model, 0, n, seed, 'cpu')
  return q.numpy(), qd.numpy()


def _acts(model, n, k):
  from brax_b200 import workloads
  return workloads.action(model, 0, n, 0, k, 'cpu').numpy()


def _close32(a, b, what, rtol=2e-5):
  scale = max(1.0, float(np.abs(b).max())) if b.size else 1.0
  np.testing.assert_allclose(a, b, rtol=rtol, atol=rtol * scale, err_msg=str(what))


def test_ant_generic_kernels_agree_with_oracle_to_rounding(ant):
  """The generic (any-size) kernels sum in the oracle's order: float32 rounding apart (fused vs unfused products)."""
  n = 16
  q, qd = _inputs(ant, 'ant', n)
  for G in (3,):   # the generic (any-size) kernel variant
    sim, o = Sim(ant, variant=G), O.Oracle(ant)
    a, b = sim.init(q, qd), o.init(q, qd)
    for f in O.STATE_FIELDS:
      _close32(a[f], b[f], ('init', G, f))
    for k in range(12):
      act = _acts('ant', n, k)
      st_in = {f: b[f].copy() for f in O.STATE_FIELDS}      # one-substep map from the oracle's state
      a = sim.step(st_in, act, 1, diag=True)
      prev = b['stats'].copy()
      o.step(b, act, 1)
      same = ((b['stats'] - prev)[:, :2] == a['stats'][:, :2]).all(1)
      for f in ('mass_mx', 'con_jac', 'con_diag', 'con_aref', 'cdof_ang', 'cdofd_vel', 'cinr_i', 'qf_smooth'):
        _close32(a[f], b[f], (k, G, f), rtol=2e-4)
      for f in ('q', 'qd', 'x_pos', 'xd_vel'):
        _close32(a[f][same], b[f][same], (k, G, f), rtol=1e-4)
      assert same.mean() >= 0.8
      o.step(b, act, 4)


@pytest.mark.parametrize('model,G', [('ant', 0), ('ant', 1), ('ant', 2), ('ant', 4), ('humanoid', 1), ('humanoid', 2), ('humanoid', 3), ('humanoid', 4)])
def test_register_row_kernels_match_oracle(model, G, ant, humanoid):
  """The register-row kernels (Newton-Schulz, constraint solve) replace sequential
  sums by group reductions, so they agree with the oracle to rounding, not to the
  bit: one-step maps from the oracle's state stay inside 1e-4/1e-5 whenever the
  solver took the same discrete branches."""
  s = {'ant': ant, 'humanoid': humanoid}[model]
  n = 16
  q, qd = _inputs(s, model, n)
  sim, o = Sim(s, variant=G), O.Oracle(s)
  b = o.init(q, qd)
  checked = 0
  for k in range(10):
    act = _acts(model, n, k)
    st_in = {f: b[f].copy() for f in O.STATE_FIELDS}
    a = sim.step(st_in, act, 1, diag=True)
    prev = b['stats'].copy()
    o.step(b, act, 1)
    e = np.zeros(n)
    for f in ('q', 'qd', 'x_pos', 'x_rot', 'xd_ang', 'xd_vel'):
      ee = np.abs(a[f] - b[f]) / (1e-5 + 1e-4 * np.abs(b[f]))
      e = np.maximum(e, ee.reshape(n, -1).max(1))
    same = ((b['stats'] - prev)[:, :2] == a['stats'][:, :2]).all(1)
    checked += int(same.sum())
    assert e[same].max() <= 1.0, (k, e)
    for f in ('mass_mx', 'con_jac', 'con_diag', 'cdof_ang', 'cinr_i'):   # not solver dependent beyond q
      np.testing.assert_allclose(a[f][same], b[f][same], rtol=1e-3, atol=1e-4, err_msg=f)
    o.step(b, act, 4)
  assert checked >= 0.8 * n * 10


def test_humanoid_matches_oracle(humanoid):
  """Humanoid takes the Newton-Schulz cold-start path every substep, where the
  group reduction of the residual norm differs from the oracle's sequential sum
  in the last bits; everything else is identical."""
  n = 8
  q, qd = _inputs(humanoid, 'humanoid', n)
  sim, o = Sim(humanoid), O.Oracle(humanoid)
  a, b = sim.init(q, qd), o.init(q, qd)
  for f in O.STATE_FIELDS:
    _close32(a[f], b[f], ('init', f))
  for k in range(6):
    act = _acts('humanoid', n, k)
    st_in = {f: b[f].copy() for f in O.STATE_FIELDS}      # one-step map from the oracle's state
    a = sim.step(st_in, act, 1, diag=True)
    prev = b['stats'].copy()
    o.step(b, act, 1)
    e = np.zeros(n)
    for f in ('q', 'qd', 'x_pos', 'x_rot', 'xd_ang', 'xd_vel'):
      ee = np.abs(a[f] - b[f]) / (1e-5 + 1e-4 * np.abs(b[f]))
      e = np.maximum(e, ee.reshape(n, -1).max(1))
    same = ((b['stats'] - prev)[:, :2] == a['stats'][:, :2]).all(1)
    assert e[same].max() <= 1.0, (k, e)
    o.step(b, act, 4)


@pytest.mark.parametrize('model,variant', [('ant', -1), ('humanoid', -1), ('humanoid', 4)])
def test_no_intra_phase_lane_dependency(model, variant, ant, humanoid):
  """Forward vs reverse lane order inside every phase must give identical bits:
  a difference would be a shared-memory race on the device.  The emulator also
  poisons the slab with NaN per env, so stale reads would surface here."""
  s = {'ant': ant, 'humanoid': humanoid}[model]
  n = 4
  q, qd = _inputs(s, model, n)
  fw, rv = Sim(s, variant=variant), Sim(s, variant=variant, reverse=True)
  a, b = fw.init(q, qd), rv.init(q, qd)
  for k in range(5):
    act = _acts(model, n, k)
    a, b = fw.step(a, act, 5), rv.step(b, act, 5)
  for f in O.STATE_FIELDS:
    assert np.isfinite(a[f]).all(), f
    assert np.array_equal(a[f], b[f]), f


def test_pendulums_no_free_joint_no_constraints():
  s = golden('triple_pendulum')   # nc == 0, nu == 0
  sim, o = Sim(s), O.Oracle(s)
  q = np.array([[0.3, -0.2, 0.1]], np.float32); qd = np.zeros((1, 3), np.float32)
  a, b = sim.init(q, qd), o.init(q, qd)
  for _ in range(20):
    a = sim.step(a, np.zeros((1, 0), np.float32), 1); o.step(b, np.zeros((1, 0), np.float32), 1)
  for f in O.STATE_FIELDS:   # 20 free-running substeps, float32 rounding apart (fused vs unfused)
    _close32(a[f], b[f], f, rtol=1e-4)


def test_matrix_inv_iterations_zero_path():
  """matrix_inv_iterations == 0: exact inverse each step and implicit damping in
  integrate (integrator.py:58-60)."""
  s = golden('double_pendulum').replace(matrix_inv_iterations=0)
  sim, o = Sim(s), O.Oracle(s)
  q = np.array([[0.4, -0.3]], np.float32); qd = np.array([[0.1, 0.2]], np.float32)
  a, b = sim.init(q, qd), o.init(q, qd)
  nu = s.nu
  act = np.zeros((1, nu), np.float32)
  for _ in range(20):
    a = sim.step(a, act, 1); o.step(b, act, 1)
  np.testing.assert_allclose(a['q'], b['q'], rtol=1e-6, atol=1e-7)


def test_cholesky_mode_differs_only_through_minv(humanoid):
  """BXG_MINV_CHOLESKY replaces Newton-Schulz by the exact inverse.  It must give
  M Minv = I; it is NOT parity with the reference for Humanoid, whose
  Newton-Schulz never converges (documented in DESIGN.md)."""
  from brax_b200 import native
  n = 4
  q, qd = _inputs(humanoid, 'humanoid', n)
  sim = Sim(humanoid, minv_mode=native.MINV_CHOLESKY)
  a = sim.init(q, qd)
  a = sim.step(a, _acts('humanoid', n, 0), 5)
  for e in range(n):
    r = a['mass_mx'][e].astype(np.float64) @ a['mass_mx_inv'][e].astype(np.float64) - np.eye(humanoid.nv)
    assert np.abs(r).max() < 2e-3
  ns = Sim(humanoid).step(Sim(humanoid).init(q, qd), _acts('humanoid', n, 0), 5)
  r = ns['mass_mx'][0].astype(np.float64) @ ns['mass_mx_inv'][0].astype(np.float64) - np.eye(humanoid.nv)
  assert np.linalg.norm(r) > 1.0   # the reference algorithm's inverse is far from converged here
