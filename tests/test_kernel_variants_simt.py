"""Kernel variants beyond Ant / Humanoid, and the size limits of the boundary
(CPU: planning query + host lane emulator vs the oracle)."""
import ctypes

import numpy as np
import pytest

from brax_b200 import native
from brax_b200.io import mjcf
from oracle import oracle as O
from tests.simt.sim import Sim
from tests.synthetic_models import centipede_xml


def _build():
  import __graft_entry__ as g
  g.build()


@pytest.mark.parametrize('n_legs,variant', [(3, 0), (5, 2), (7, 6), (10, 3)])
def test_variant_selection_and_parity(n_legs, variant):
  """3 legs: half-warp Ant class; 5 legs: nc = 30 -> the 32/32 tile variant (4x8 tiles);
  7 legs: nv = 20, nc = 42 -> the 24-dof / 80-row variant (rows of A in shared memory, 128-bit active mask);
  10 legs: nv = 26 > 24 -> generic any-size kernel."""
  _build()
  s = mjcf.loads(centipede_xml(n_legs))
  plan = native.plan(s)
  assert plan['variant'] == variant, plan
  assert plan['nc'] == 4 * n_legs + 2 * n_legs
  rng = np.random.default_rng(0)
  n = 6
  q = (np.asarray(s.init_q)[None] + rng.uniform(-0.05, 0.05, (n, s.nq))).astype(np.float32)
  q[:, 2] = 0.2 + 0.1 * rng.uniform(size=n)          # low enough for the feet to touch
  qd = (0.1 * rng.standard_normal((n, s.nv))).astype(np.float32)
  sim, o = Sim(s), O.Oracle(s)
  a, b = sim.init(q, qd), o.init(q, qd)
  for f in O.STATE_FIELDS:
    np.testing.assert_allclose(a[f], b[f], rtol=1e-5, atol=1e-6, err_msg=f)
  active = 0
  for k in range(8):
    act = rng.uniform(-1, 1, (n, s.nu)).astype(np.float32)
    st_in = {f: b[f].copy() for f in O.STATE_FIELDS}
    a = sim.step(st_in, act, 1, diag=True)
    prev = b['stats'].copy()
    o.step(b, act, 1)
    same = ((b['stats'] - prev)[:, :2] == a['stats'][:, :2]).all(1)
    e = np.zeros(n)
    for f in ('q', 'qd', 'x_pos', 'x_rot', 'xd_ang', 'xd_vel'):
      ee = np.abs(a[f] - b[f]) / (1e-5 + 1e-4 * np.abs(b[f]))
      e = np.maximum(e, ee.reshape(n, -1).max(1))
    # same-branch envs: tight in the bulk; a lone env may sit just outside (fp sensitivity)
    assert same.sum() >= n // 2 and np.median(e[same]) <= 0.2 and (e[same] > 5.0).sum() <= 1, (k, e, same)
    active += int((b['con_dist'] < 0).sum())
    o.step(b, act, 4)
  assert active > 0          # contacts were exercised
  # reverse lane order: no intra-phase dependency in this variant either
  r = Sim(s, reverse=True)
  x, y = sim.step(sim.init(q, qd), np.zeros((n, s.nu), np.float32), 3), r.step(r.init(q, qd), np.zeros((n, s.nu), np.float32), 3)
  for f in O.STATE_FIELDS:
    assert np.array_equal(x[f], y[f]), f


def test_models_beyond_the_limits_are_rejected():
  """Maximum sizes: 32 links / 64 dofs / 64 constraint rows.  Larger models must fail
  loudly at planning time (BXG_E_UNSUPPORTED), never run a wrong kernel."""
  _build()
  lib = native.lib()
  big = mjcf.loads(centipede_xml(16))       # 33 links, nv = 38, nc = 96
  desc, keep = native.make_desc(big)
  info = (ctypes.c_int32 * 8)()
  rc = lib.bxg_plan(ctypes.byref(desc), info)
  assert rc == 3 and b'not supported' in lib.bxg_last_error()
  ok = mjcf.loads(centipede_xml(10))        # 21 links, nv = 26, nc = 60: the generic kernel takes it
  assert native.plan(ok)['variant'] == 3


@pytest.mark.parametrize('model', ['inverted_pendulum', 'inverted_double_pendulum', 'reacher'])
def test_four_lane_variant(model):
  """Classic-control models (<= 4 links / dofs / rows) take the 4-lane variant: eight envs per warp, 4 x 4
  register tiles.  Parity with the oracle as one-substep maps, forward and reverse lane order."""
  _build()
  from brax_b200 import envs_assets
  s = envs_assets.load(model)
  plan = native.plan(s)
  assert plan['variant'] == 7 and plan['lanes_per_env'] == 4, plan
  rng = np.random.default_rng(0)
  n = 11                                   # not a multiple of the eight envs of a warp
  q = (np.asarray(s.init_q)[None] + rng.uniform(-0.3, 0.3, (n, s.nq))).astype(np.float32)
  if model == 'inverted_double_pendulum':
    q[:, 1] = 1.6 * rng.uniform(-1, 1, n)  # beyond the joint range (+-1.57?): limit rows active for some envs
  qd = rng.standard_normal((n, s.nv)).astype(np.float32)
  sim, o = Sim(s), O.Oracle(s)
  a, b = sim.init(q, qd), o.init(q, qd)
  for f in O.STATE_FIELDS:
    np.testing.assert_allclose(a[f], b[f], rtol=1e-5, atol=1e-6 * max(1.0, float(np.abs(b[f]).max()) if b[f].size else 1.0), err_msg=f)
  inside = []
  for k in range(8):
    act = rng.uniform(-1, 1, (n, s.nu)).astype(np.float32)
    st_in = {f: b[f].copy() for f in O.STATE_FIELDS}
    a = sim.step(st_in, act, 1)
    o.step(b, act, 1)
    e = np.zeros(n)
    for f in ('q', 'qd', 'x_pos', 'x_rot', 'xd_ang', 'xd_vel'):
      e = np.maximum(e, (np.abs(a[f] - b[f]) / (1e-5 + 1e-4 * np.abs(b[f]))).reshape(n, -1).max(1))
    inside.append(e <= 1.0)
  assert np.mean(inside) >= 0.9, np.mean(inside)
  r = Sim(s, reverse=True)
  z = np.zeros((n, s.nu), np.float32)
  x, y = sim.step(sim.init(q, qd), z, 3), r.step(r.init(q, qd), z, 3)
  for f in O.STATE_FIELDS:
    assert np.array_equal(x[f], y[f]), f
