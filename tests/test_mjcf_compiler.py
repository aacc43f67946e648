"""The MuJoCo-free model compiler (brax_b200/io/mjcf.py)."""
import os

import numpy as np
import pytest

from brax_b200.io import mjcf, model_json
from tests.conftest import ROOT

REF = '/root/reference/brax'

_PENDULUM = """
<mujoco model="pendulum">
  <compiler inertiafromgeom="true"/>
  <option gravity="0 0 -9.81" timestep="0.02"/>
  <worldbody>
    <body name="body1" pos="0 0 0">
      <joint axis="1 0 0" name="hinge1" pos="0 0 0" type="hinge"/>
      <geom name="sphere1" pos="0 0.5 0" size=".15" mass="1" type="sphere"/>
      <body name="fused" pos="0 0.25 0">
        <geom name="cap" fromto="0 0 0 0 0.25 0" size="0.05" type="capsule" density="100"/>
      </body>
      <body name="body2" pos="0 0.5 0">
        <joint axis="1 0 0" name="hinge2" pos="0 0 0" type="hinge" range="-30 30"/>
        <geom name="sphere2" pos="0 0.5 0" size=".15" mass="1" type="sphere"/>
      </body>
    </body>
  </worldbody>
  <actuator><motor joint="hinge2" gear="10" ctrlrange="-1 1"/></actuator>
</mujoco>
"""


def test_inline_model_and_body_fusing():
  s = mjcf.loads(_PENDULUM)
  assert s.link_types == '11' and s.link_parents == (-1, 0)
  assert s.link_names == ['body1', 'body2']          # the joint-less body was dissolved
  # the capsule of the dissolved body moved into body1's frame: it adds mass there
  r, h = 0.05, 0.25
  cap_mass = 100 * np.pi * (r * r * h + 4 / 3 * r ** 3)
  np.testing.assert_allclose(s.link.inertia.mass, [1 + cap_mass, 1], rtol=1e-6)
  com_y = (1 * 0.5 + cap_mass * (0.25 + 0.125)) / (1 + cap_mass)
  np.testing.assert_allclose(s.link.inertia.transform.pos[0], [0, com_y, 0], atol=1e-6)
  # sphere inertia 2/5 m r^2
  np.testing.assert_allclose(np.diag(s.link.inertia.i[1]), [0.4 * 0.15 ** 2] * 3, rtol=1e-6)
  # limits: only hinge2 is limited (autolimits), degrees -> radians
  lo, hi = s.dof.limit
  assert lo[0] == -np.inf and hi[0] == np.inf
  np.testing.assert_allclose([lo[1], hi[1]], np.deg2rad([-30, 30]), rtol=1e-6)
  assert s.nu == 1 and int(s.actuator.qd_id[0]) == 1 and float(s.actuator.gear[0]) == 10
  assert s.matrix_inv_iterations == 10 and s.solver_maxls == 20 and s.solver_iterations == 100


def test_invweight_matches_inverse_mass_matrix_diagonal():
  s = mjcf.loads(_PENDULUM)
  minv = np.linalg.inv(np.asarray(s.mass_mx0, np.float64))
  np.testing.assert_allclose(s.dof.invweight, np.diag(minv), rtol=1e-5)
  assert (np.asarray(s.link.invweight) > 0).all()


def test_unsupported_features_raise():
  with pytest.raises(NotImplementedError):
    mjcf.loads(_PENDULUM.replace('type="hinge"/>', 'type="ball"/>', 1))
  with pytest.raises(NotImplementedError):
    mjcf.loads(_PENDULUM.replace('timestep="0.02"', 'timestep="0.02" integrator="RK4"'))
  with pytest.raises(NotImplementedError):
    mjcf.loads(_PENDULUM.replace('name="hinge1"', 'name="hinge1" ref="0.1"'))


def test_json_round_trip(ant):
  s2 = model_json.loads(model_json.dumps(ant))
  from brax_b200.base import tree_leaves
  for a, b in zip(tree_leaves(ant), tree_leaves(s2)):
    assert np.array_equal(np.asarray(a), np.asarray(b))
  assert s2.link_types == ant.link_types and s2.link_parents == ant.link_parents


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference tree not mounted (GPU box)')
@pytest.mark.parametrize('name', ['ant', 'humanoid'])
def test_shipped_assets_match_a_fresh_compile(name):
  """brax_b200/assets/*.json are exactly what tools/gen_assets.py produces from
  the reference's MJCF (brax/envs/assets/*.xml)."""
  from brax_b200 import envs_assets
  from brax_b200.base import tree_leaves
  fresh = mjcf.load(f'{REF}/envs/assets/{name}.xml')
  shipped = envs_assets.load(name)
  for a, b in zip(tree_leaves(fresh), tree_leaves(shipped)):
    assert np.array_equal(np.asarray(a), np.asarray(b))


def test_plan_picks_expected_variants(ant, humanoid):
  from brax_b200 import native
  import __graft_entry__ as g
  g.build()
  pa, ph = native.plan(ant), native.plan(humanoid)
  assert pa['variant'] == 0 and pa['lanes_per_env'] == 16 and pa['nc'] == 24
  assert ph['variant'] == 1 and ph['lanes_per_env'] == 32 and ph['nc'] == 25
  for p in (pa, ph):
    assert p['smem_bytes_per_cta'] <= 227 * 1024 and p['envs_per_cta'] >= 8
