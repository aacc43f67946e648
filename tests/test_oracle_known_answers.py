"""Pins the oracle (oracle/bxg_oracle.c) and the model compiler against every
known answer the reference's own tests carry for this path and that can be
checked without a live MuJoCo."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.conftest import golden


def _rollout(sys, q, qd, act, dt, n, dtype=np.float32):
  sys = sys.tree_replace({'opt.timestep': np.float32(dt)})
  o = O.Oracle(sys, dtype)
  st = o.init(np.asarray(q, dtype)[None], np.asarray(qd, dtype)[None])
  act = np.asarray(act, dtype)[None]
  for _ in range(n):
    o.step(st, act, 1)
  return st['q'][0], st['qd'][0]


def test_sphere_plane_ant(ant):
  # reference brax/contact_test.py:29-44
  o = O.Oracle(ant)
  q = np.array([0, 0, 0.556008, 1, 0, 0, 0, 0, 1, 0, 1, 0, 1, 0, 1], np.float32)
  dist, pos = o.contact(q)
  cp = ant.contact_pairs()
  assert dist.shape[0] == 4
  np.testing.assert_array_almost_equal(pos[0], [0.61612, 0.61612, 0])
  np.testing.assert_array_almost_equal(cp.frame[0][0], [0, 0, 1])
  np.testing.assert_array_almost_equal(dist[0], 0)
  np.testing.assert_array_almost_equal(cp.friction[0], 1)
  assert (int(cp.link_a[0]), int(cp.link_b[0])) == (-1, 2)


def test_load_pendulum():
  # reference brax/io/mjcf_test.py:29-54
  sys = golden('triple_pendulum')
  np.testing.assert_array_almost_equal(sys.gravity, [0, 0, -9.81])
  assert sys.link_names == ['body1', 'body2', 'body3']
  np.testing.assert_array_almost_equal(sys.link.transform.pos, [[0, 0, 0], [0, 0.5, 0], [0, 0.5, 0]])
  np.testing.assert_array_almost_equal(sys.link.transform.rot, [[1, 0, 0, 0]] * 3)
  np.testing.assert_array_almost_equal(sys.link.inertia.i, np.tile(np.eye(3), (3, 1, 1)) * 0.009)
  np.testing.assert_array_almost_equal(sys.link.inertia.transform.pos, [[0, 0.5, 0]] * 3)
  np.testing.assert_array_almost_equal(sys.link.inertia.mass, [1, 1, 1])
  assert sys.link_types == '111'
  assert sys.link_parents == (-1, 0, 1)


def test_load_ant(ant):
  # reference brax/io/mjcf_test.py:56-67
  assert ant.link_names == ['torso', 'aux_1', '', 'aux_2', '', 'aux_3', '', 'aux_4', '']
  assert ant.link_types == 'f11111111'
  assert ant.link_parents == (-1, 0, 1, 0, 3, 0, 5, 0, 7)


def test_load_humanoid(humanoid):
  # reference brax/io/mjcf_test.py:69-89
  assert humanoid.link_names == [
      'torso', 'lwaist', 'pelvis', 'right_thigh', 'right_shin', 'left_thigh', 'left_shin',
      'right_upper_arm', 'right_lower_arm', 'left_upper_arm', 'left_lower_arm']
  assert humanoid.link_types == 'f2131312121'


def test_init_q(ant):
  # reference brax/kinematics_test.py:69-75
  np.testing.assert_almost_equal(ant.init_q, [0, 0, 0.55, 1, 0, 0, 0, 0, 1, 0, -1, 0, -1, 0, 1], 7)


def test_tree_levels(ant, humanoid):
  # reference brax/scan_test.py:49-72 (Ant level groups) / SURVEY appendix B
  def levels(sys):
    out = {}
    for i in range(sys.num_links()):
      out.setdefault(sys.link_depth(i), []).append(i)
    return [out[k] for k in sorted(out)]
  assert levels(ant) == [[0], [1, 3, 5, 7], [2, 4, 6, 8]]
  assert levels(humanoid) == [[0], [1, 7, 9], [2, 8, 10], [3, 5], [4, 6]]


def test_motor():
  # reference brax/actuator_test.py:50-64 (g_pipeline, dt 0.01, n 100, 5 dp)
  sys = golden('single_pendulum_motor')
  o = O.Oracle(sys)
  q, qd = np.zeros(1, np.float32), np.zeros(1, np.float32)
  act = np.array([1.0 / 150.0 * 0.5 * 9.81], np.float32)
  tau = o.to_tau(q, qd, act)
  np.testing.assert_array_almost_equal(tau, [0.5 * 9.81], 5)
  q2, qd2 = _rollout(sys, q, qd, act, 0.01, 100)
  np.testing.assert_array_almost_equal(q2, [0], decimal=5)
  np.testing.assert_array_almost_equal(qd2, [0], decimal=5)


def test_position():
  # reference brax/actuator_test.py:66-85
  sys = golden('single_pendulum_position')
  o = O.Oracle(sys)
  theta = np.pi / 2.0
  q, qd = np.array([theta], np.float32), np.zeros(1, np.float32)
  tau = o.to_tau(q, qd, np.array([theta], np.float32))
  np.testing.assert_array_almost_equal(tau, [0], 5)
  act = np.array([-(theta * 0.5**2) / (0.01**2 * 10.0) + theta], np.float32)
  q2, _ = _rollout(sys, q, qd, act, 0.01, 1)
  np.testing.assert_array_almost_equal(q2, [0], 1)


def test_velocity():
  # reference brax/actuator_test.py:87-103
  sys = golden('single_pendulum_velocity')
  o = O.Oracle(sys)
  theta = np.pi / 2.0
  q, qd = np.array([theta], np.float32), np.zeros(1, np.float32)
  tau = o.to_tau(q, qd, np.zeros(1, np.float32))
  np.testing.assert_array_almost_equal(tau, [0], 5)
  _, qd2 = _rollout(sys, q, qd, [1.0], 0.001, 200)
  np.testing.assert_array_almost_equal(qd2, [1], 3)


def test_force_limited():
  # reference brax/actuator_test.py:105-118: tau == frclimit * gear exactly
  sys = golden('single_pendulum_position_frclimit')
  o = O.Oracle(sys)
  q, qd = np.zeros(1, np.float32), np.zeros(1, np.float32)
  for act, frclimit in [(1000, 3.1), (-1000, -2.5)]:
    tau = o.to_tau(q, qd, np.array([act], np.float32))
    assert tau[0] == np.float32(frclimit) * np.float32(10)


def test_three_link_pendulum():
  # reference brax/actuator_test.py:120-144
  sys = golden('triple_pendulum_motor')
  theta = np.pi / 2.0
  q, qd = np.array([theta, 0, 0], np.float32), np.zeros(3, np.float32)
  q1, qd1 = _rollout(sys, q, qd, [0, 0, 0], 0.01, 1)
  np.testing.assert_array_almost_equal(q1, q, 2)
  np.testing.assert_array_almost_equal(qd1, np.zeros(3), 2)
  act = 1.0 / 150.0 * np.array([0, -10, 10])
  q2, qd2 = _rollout(sys, q, qd, act, 1e-3, 1)
  assert abs(q2[0] - q[0]) < 0.5e-2
  assert q2[1] < q[1] and q2[2] > q[2]
  assert qd2[1] < -0.2 and qd2[2] > -0.2


def test_inv_approximate():
  # reference brax/math_test.py:38-46: 100 iterations from zero reach inv(x)
  rng = np.random.default_rng(0)
  x = rng.standard_normal((4, 4)).astype(np.float32)
  x = (np.eye(4) * 0.001 + x @ x.T).astype(np.float32)
  got = O.inv_approximate(x, np.zeros((4, 4), np.float32), 100)
  np.testing.assert_array_almost_equal(got, np.linalg.inv(x.astype(np.float64)), decimal=4)
  got64 = O.inv_approximate(x, np.zeros((4, 4)), 100, np.float64)
  np.testing.assert_array_almost_equal(got64, np.linalg.inv(x.astype(np.float64)))
