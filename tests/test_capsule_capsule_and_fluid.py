"""Capsule-capsule contacts between two moving links (Pusher) and fluid forces (Swimmer): the
oracle's restatement of mjx `capsule_capsule` and of `brax/fluid.py` against hand-derived answers, and the
kernel source (host lane emulator) against the oracle, forward and reverse lane order.

The reference's tests hold no numbers for either (contact_test.py covers plane-sphere and
sphere-sphere only; fluid is exercised through live-MuJoCo differential tests), so the answers below
are derived by hand from the published formulas; the jacobian / dynamics side of both is pinned by
tests/golden/ref_{pusher, swimmer}.npz (tests/test_reference_golden.py)."""
import numpy as np
import pytest

from brax_b200 import envs_assets, native
from brax_b200.io import mjcf
from oracle import oracle as O
from tests.simt.sim import Sim

# two free bodies, each a capsule of radius 0.1, half length 0.5: `a` along x, `b` along y
CROSSED = """
<mujoco model="crossed_capsules">
  <compiler angle="radian" inertiafromgeom="true"/>
  <option timestep="0.01"/>
  <default><geom friction="0.6 0.1 0.1"/></default>
  <worldbody>
    <body name="a" pos="0 0 1">
      <joint name="ra" type="free"/>
      <geom name="ca" fromto="-0.5 0 0 0.5 0 0" size="0.1" type="capsule" contype="1" conaffinity="0"/>
    </body>
    <body name="b" pos="0 0 1.15">
      <joint name="rb" type="free"/>
      <geom name="cb" fromto="0 -0.5 0 0 0.5 0" size="0.1" type="capsule" contype="0" conaffinity="1"/>
    </body>
  </worldbody>
</mujoco>
"""


def _build():
  import __graft_entry__ as g
  g.build()


def test_capsule_capsule_pair_table():
  s = mjcf.loads(CROSSED)
  cp = s.contact_pairs()
  assert list(cp.kind) == [2] and list(cp.link_a) == [0] and list(cp.link_b) == [1]
  np.testing.assert_allclose([cp.a_half[0], cp.a_radius[0], cp.half_len[0], cp.radius[0]], [0.5, 0.1, 0.5, 0.1])
  assert native.plan(s)['variant'] in (3, 5)       # the only variants that carry the two-body rows


@pytest.mark.parametrize('gap,shift', [(0.15, 0.0), (0.25, 0.2), (0.12, -0.3)])
def test_oracle_capsule_capsule_geometry(gap, shift):
  """Crossed capsules (axes x and y), `b` a height `gap` above `a` and moved by `shift` along x: the closest
  points are (shift, 0, z_a) and (shift, 0, z_a + gap); dist = gap - 2 r; the normal points from a to b."""
  _build()
  s = mjcf.loads(CROSSED)
  o = O.Oracle(s, np.float64)
  q = np.array([0, 0, 1.0, 1, 0, 0, 0, shift, 0, 1.0 + gap, 1, 0, 0, 0])
  dist, pos = o.contact(q)
  np.testing.assert_allclose(dist[0], gap - 0.2, atol=2e-6)
  # pos = pt_a + n (r_a + dist / 2): half way between the two surfaces
  np.testing.assert_allclose(pos[0], [shift, 0, 1.0 + 0.1 + 0.5 * (gap - 0.2)], atol=5e-6)


def test_oracle_capsule_capsule_end_to_end_clamp():
  """`b` moved beyond the end of `a` along x: the closest point on `a` is its end sphere centre (0.5, 0, z)."""
  _build()
  s = mjcf.loads(CROSSED)
  o = O.Oracle(s, np.float64)
  q = np.array([0, 0, 1.0, 1, 0, 0, 0, 0.62, 0, 1.09, 1, 0, 0, 0])
  dist, pos = o.contact(q)
  d = np.hypot(0.12, 0.09)
  np.testing.assert_allclose(dist[0], d - 0.2, atol=2e-6)
  n = np.array([0.12, 0, 0.09]) / d
  np.testing.assert_allclose(pos[0], np.array([0.5, 0, 1.0]) + n * (0.1 + 0.5 * (d - 0.2)), atol=5e-6)


def test_crossed_capsules_push_each_other_apart():
  """Two-body rows end to end: in free fall without gravity the penetrating pair separates along the normal with
  equal and opposite momentum (both bodies have the same mass)."""
  _build()
  s = mjcf.loads(CROSSED)
  s = s.replace(gravity=np.zeros(3, np.float32), solver_iterations=100, matrix_inv_iterations=0)
  o = O.Oracle(s, np.float64)
  st = o.init(np.array([[0, 0, 1.0, 1, 0, 0, 0, 0, 0, 1.15, 1, 0, 0, 0]]), np.zeros((1, 12)))
  for _ in range(20):
    o.step(st, np.zeros((1, 0)), 1)
  va, vb = st['qd'][0, 2], st['qd'][0, 8]
  assert va < -1e-3 and vb > 1e-3 and abs(va + vb) < 1e-6 * max(1.0, abs(vb)) + 1e-9, (va, vb)


def test_fluid_drag_of_a_translating_box():
  """brax/fluid.py on one free body moving along x without rotation or gravity: force = -3 pi d mu v
  - 0.5 rho b_y b_z |v| v with the box of the equivalent inertia, so qdd_x = force / mass after one init."""
  _build()
  xml = '''
<mujoco model="box">
  <compiler angle="radian" inertiafromgeom="true"/>
  <option timestep="0.01" density="1.2" viscosity="0.05" gravity="0 0 0"/>
  <worldbody>
    <body name="b" pos="0 0 1"><joint name="r" type="free"/><geom name="g" type="sphere" size="0.2" density="500"/></body>
  </worldbody>
</mujoco>'''
  s = mjcf.loads(xml)
  assert s.enable_fluid and native.plan(s)['variant'] in (3, 8)      # the only variants that carry the fluid forces
  o = O.Oracle(s, np.float64)
  v = 0.7
  st = o.init(np.array([[0, 0, 1.0, 1, 0, 0, 0]]), np.array([[v, 0, 0, 0, 0, 0]]))
  o.step(st, np.zeros((1, 0)), 1)
  mass = float(s.link.inertia.mass[0])
  inertia = np.diag(np.asarray(s.link.inertia.i[0], np.float64))
  box = np.sqrt(6.0 * np.array([inertia[1] + inertia[2] - inertia[0], inertia[0] + inertia[2] - inertia[1],
                                inertia[0] + inertia[1] - inertia[2]]) / mass)
  force = -3.0 * np.pi * box.mean() * 0.05 * v - 0.5 * 1.2 * box[1] * box[2] * abs(v) * v
  np.testing.assert_allclose(st['qf_smooth'][0, 0], force, rtol=1e-5)
  np.testing.assert_allclose(st['qf_smooth'][0, 1:], 0, atol=1e-9)


@pytest.mark.parametrize('model', ['pusher', 'swimmer'])
def test_kernel_source_matches_oracle(model):
  _build()
  s = envs_assets.load(model)
  n = 8
  rng = np.random.default_rng(0)
  q = (np.asarray(s.init_q)[None] + rng.uniform(-0.05, 0.05, (n, s.nq))).astype(np.float32)
  qd = (0.1 * rng.standard_normal((n, s.nv))).astype(np.float32)
  o = O.Oracle(s)
  if model == 'pusher':      # arm lowered onto the table, the object next to the wrist: both contact kinds active
    q[:, :7] = 0.0
    q[:, 1] = 0.40 + 0.05 * rng.uniform(size=n)
    w = o.init(q, qd)['x_pos'][:, 6]
    q[:, 7], q[:, 8] = w[:, 1] + 0.07, w[:, 0] - 0.40
  sim = Sim(s)
  a, b = sim.init(q, qd), o.init(q, qd)
  for f in O.STATE_FIELDS:
    loose = 10.0 if f.startswith('con_') else 1.0     # capsule-capsule tie-break in float32 (test_reference_golden.py)
    np.testing.assert_allclose(a[f], b[f], rtol=1e-5 * loose, atol=1e-6 * loose * max(1.0, float(np.abs(b[f]).max()) if b[f].size else 1.0), err_msg=f)
  inside, active = [], 0
  for k in range(8):
    act = rng.uniform(-1, 1, (n, s.nu)).astype(np.float32)
    st_in = {f: b[f].copy() for f in O.STATE_FIELDS}
    a = sim.step(st_in, act, 1, diag=True)
    o.step(b, act, 1)
    e = np.zeros(n)
    for f in ('q', 'qd', 'x_pos', 'x_rot', 'xd_ang', 'xd_vel'):
      e = np.maximum(e, (np.abs(a[f] - b[f]) / (1e-5 + 1e-4 * np.abs(b[f]))).reshape(n, -1).max(1))
    inside.append(e <= 1.0)
    active += int((b['con_dist'] < 0).sum()) if b['con_dist'].size else 0
  assert np.mean(inside) >= 0.85, np.mean(inside)
  assert model == 'swimmer' or active > 0
  r = Sim(s, reverse=True)      # reverse lane order: no intra-phase dependency on these paths either
  z = np.zeros((n, s.nu), np.float32)
  x, y = sim.step(sim.init(q, qd), z, 3), r.step(r.init(q, qd), z, 3)
  for f in O.STATE_FIELDS:
    assert np.array_equal(x[f], y[f]), f


def test_plan_for_a_model_with_fluid_forces_and_two_body_contacts():
  """Variant selection (bxg_model.h): a model needing BOTH optional code paths goes to the generic variant, which
  carries both, instead of being rejected; a forced variant lacking one of them is still an error."""
  import ctypes
  import numpy as np
  from brax_b200 import envs_assets, native
  pusher = envs_assets.load('pusher')                       # capsule-capsule contacts between moving links
  assert native.plan(pusher)['variant'] == 5
  both = pusher.replace(enable_fluid=True, viscosity=np.float32(0.1), density=np.float32(1.2))
  p = native.plan(both)
  assert p['variant'] == 3 and p['kernel_id'] == 3          # the generic any-size kernel: fluid + two-body rows
  swimmer = envs_assets.load('swimmer')                     # fluid only: the small 8-wide variant
  assert native.plan(swimmer)['variant'] == 8
  # the host emulator runs it (same selection code), finite results
  from tests.simt.sim import Sim
  sim = Sim(both)
  q = np.asarray(both.init_q, np.float32)[None].repeat(2, 0); qd = np.full((2, both.nv), 0.1, np.float32)
  st = sim.step(sim.init(q, qd), np.zeros((2, both.nu), np.float32), 2)
  assert all(np.isfinite(v).all() for v in st.values())
