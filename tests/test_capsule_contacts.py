"""Plane-capsule contacts: the oracle's restatement of
mjx `plane_capsule` against hand-derived geometry, and the kernel source (host lane
emulator) against the oracle on HalfCheetah, Hopper and Walker2d.

The reference's own tests hold no numbers for this pair type (capsule.xml is only
used for `elasticity` and frame round trips: io/mjcf_test.py:91-93, com_test.py:30-93),
so the geometry below is derived by hand from the published mjx formulas."""
import numpy as np
import pytest

from brax_b200 import envs_assets, native
from brax_b200.io import mjcf
from oracle import oracle as O
from tests.simt.sim import Sim

# one free body carrying a capsule of radius 0.25 along its x axis, half length 0.5
LYING_CAPSULE = """
<mujoco model="lying_capsule">
  <compiler angle="radian" inertiafromgeom="true"/>
  <option timestep="0.01"/>
  <default><geom conaffinity="0" contype="1" friction="0.7 0.1 0.1"/></default>
  <worldbody>
    <geom conaffinity="1" contype="0" name="floor" pos="0 0 0" size="40 40 40" type="plane"/>
    <body name="b" pos="0 0 0.25">
      <joint name="root" type="free"/>
      <geom name="cap" fromto="-0.5 0 0 0.5 0 0" size="0.25" type="capsule"/>
    </body>
  </worldbody>
</mujoco>
"""


def _build():
  import __graft_entry__ as g
  g.build()


def test_capsule_pair_table():
  s = mjcf.loads(LYING_CAPSULE)
  cp = s.contact_pairs()
  assert list(cp.kind) == [1, 1] and list(cp.link_a) == [-1, -1] and list(cp.link_b) == [0, 0]
  np.testing.assert_allclose(cp.half_len, [0.5, -0.5])          # +axis end first (mjx: [segment, -segment])
  np.testing.assert_allclose(cp.radius, [0.25, 0.25])
  np.testing.assert_allclose(cp.friction, [0.7, 0.7])
  # fromto along +x: the geom's z axis is rotated onto x
  ax = np.array([2 * (cp.geom_quat[0][1] * cp.geom_quat[0][3] + cp.geom_quat[0][0] * cp.geom_quat[0][2]),
                 2 * (cp.geom_quat[0][2] * cp.geom_quat[0][3] - cp.geom_quat[0][0] * cp.geom_quat[0][1]),
                 1 - 2 * (cp.geom_quat[0][1] ** 2 + cp.geom_quat[0][2] ** 2)])
  np.testing.assert_allclose(np.abs(ax), [1, 0, 0], atol=1e-6)


@pytest.mark.parametrize('pitch', [0.0, 0.3, -0.7])
def test_oracle_plane_capsule_geometry(pitch):
  """dist = z_end - r, pos = end - n (r + dist / 2) for both end spheres of a capsule
  pitched about y; the order of the two contacts follows the geom's own axis."""
  _build()
  s = mjcf.loads(LYING_CAPSULE)
  cp = s.contact_pairs()
  o = O.Oracle(s, np.float64)
  q = np.array([0.1, -0.2, 0.4, np.cos(pitch / 2), 0, np.sin(pitch / 2), 0])
  dist, pos = o.contact(q)
  # world axis of the geom = R_y(pitch) applied to its local axis (+x or -x after fromto)
  gq = np.asarray(cp.geom_quat[0], np.float64)
  local = np.array([2 * (gq[1] * gq[3] + gq[0] * gq[2]), 2 * (gq[2] * gq[3] - gq[0] * gq[1]), 1 - 2 * (gq[1] ** 2 + gq[2] ** 2)])
  c, sn = np.cos(pitch), np.sin(pitch)
  axis = np.array([c * local[0] + sn * local[2], local[1], -sn * local[0] + c * local[2]])
  for k, half in enumerate((0.5, -0.5)):
    end = q[:3] + axis * half
    d = end[2] - 0.25
    np.testing.assert_allclose(dist[k], d, atol=1e-6)      # System constants are float32
    np.testing.assert_allclose(pos[k], end - np.array([0, 0, 1]) * (0.25 + 0.5 * d), atol=1e-6)


def test_lying_capsule_rests_on_the_floor():
  """Converged behaviour pins the contact rows end to end: a capsule dropped flat from 5 mm
  settles with both end spheres carrying its weight (sub-millimetre penetration, no drift)."""
  _build()
  s = mjcf.loads(LYING_CAPSULE)
  s = s.replace(solver_iterations=200, matrix_inv_iterations=0)
  o = O.Oracle(s, np.float64)
  st = o.init(np.array([[0, 0, 0.255, 1, 0, 0, 0]]), np.zeros((1, 6)))
  for _ in range(300):
    o.step(st, np.zeros((1, 0)), 1)
  assert abs(st['q'][0, 2] - 0.25) < 2e-3, st['q'][0]
  assert np.abs(st['qd'][0]).max() < 2e-2, st['qd'][0]
  assert (st['con_dist'][0] < 1e-3).all()


def _reset(s, n, seed, drop):
  rng = np.random.default_rng(seed)
  q = (np.asarray(s.init_q)[None] + rng.uniform(-0.1, 0.1, (n, s.nq))).astype(np.float32)
  q[:, 1] -= drop * rng.uniform(0.5, 1.0, n).astype(np.float32)      # rootz: towards the floor
  qd = (0.1 * rng.standard_normal((n, s.nv))).astype(np.float32)
  return q, qd


@pytest.mark.parametrize('model,variant,drop', [('hopper', 9, 0.05), ('walker2d', 5, 0.1), ('halfcheetah', 5, 0.45)])
def test_kernel_source_matches_oracle_on_capsule_models(model, variant, drop):
  _build()
  s = envs_assets.load(model)
  assert native.plan(s)['variant'] == variant
  n = 8
  q, qd = _reset(s, n, 0, drop)
  sim, o = Sim(s), O.Oracle(s)
  a, b = sim.init(q, qd), o.init(q, qd)
  for f in O.STATE_FIELDS:   # float32 rounding apart: the kernel source fuses its multiply-adds, the oracle does not
    np.testing.assert_allclose(a[f], b[f], rtol=1e-4, atol=2e-5 * max(1.0, float(np.abs(b[f]).max()) if b[f].size else 1.0), err_msg=f)
  rng = np.random.default_rng(1)
  active, checked = 0, 0
  for k in range(10):
    act = rng.uniform(-1, 1, (n, s.nu)).astype(np.float32)
    st_in = {f: b[f].copy() for f in O.STATE_FIELDS}
    a = sim.step(st_in, act, 1, diag=True)
    prev = b['stats'].copy()
    o.step(b, act, 1)
    np.testing.assert_allclose(a['con_dist'], b['con_dist'], rtol=1e-5, atol=1e-6)
    assert np.array_equal(a['con_dist'] < 0, b['con_dist'] < 0) or np.abs(b['con_dist']).min() < 1e-6
    same = ((b['stats'] - prev)[:, :2] == a['stats'][:, :2]).all(1)
    e = np.zeros(n)
    for f in ('q', 'qd', 'x_pos', 'x_rot', 'xd_ang', 'xd_vel'):
      ee = np.abs(a[f] - b[f]) / (1e-5 + 1e-4 * np.abs(b[f]))
      e = np.maximum(e, ee.reshape(n, -1).max(1))
    checked += int(same.sum())
    if same.any():
      assert np.median(e[same]) <= 0.2 and e[same].max() <= 5.0, (k, e, same)
    active += int((b['con_dist'] < 0).sum())
    o.step(b, act, 4)
  assert active > 0 and checked >= 0.6 * n * 10
  # forward vs reverse lane order: no intra-phase dependency on the capsule path either
  r = Sim(s, reverse=True)
  z = np.zeros((n, s.nu), np.float32)
  x, y = sim.step(sim.init(q, qd), z, 3), r.step(r.init(q, qd), z, 3)
  for f in O.STATE_FIELDS:
    assert np.array_equal(x[f], y[f]), f
