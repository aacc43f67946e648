"""Physics invariants that pin the oracle where the reference has no stored
golden vectors (SURVEY.md section 8c "supplementary oracles")."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.conftest import golden


def _qmat(q):
  w, x, y, z = q / np.linalg.norm(q)
  return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                   [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                   [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def _qmul(u, v):
  return np.array([u[0] * v[0] - u[1] * v[1] - u[2] * v[2] - u[3] * v[3],
                   u[0] * v[1] + u[1] * v[0] + u[2] * v[3] - u[3] * v[2],
                   u[0] * v[2] - u[1] * v[3] + u[2] * v[0] + u[3] * v[1],
                   u[0] * v[3] + u[1] * v[2] - u[2] * v[1] + u[3] * v[0]])


@pytest.mark.parametrize('name', ['ant', 'humanoid'])
def test_mass_matrix_equals_jacobian_mass_matrix(name, ant, humanoid):
  """CRBA mass matrix (kinematics -> transform_com -> mass.matrix) equals
  sum_l m Jv^T Jv + Jw^T I Jw + armature with Jacobians taken by finite
  differences of the oracle's own forward kinematics: pins cdof, cinr, crb."""
  s = {'ant': ant, 'humanoid': humanoid}[name]
  o = O.Oracle(s, np.float64)
  rng = np.random.default_rng(1)
  q = s.init_q.astype(np.float64) + rng.uniform(-0.3, 0.3, s.nq)
  q[3:7] /= np.linalg.norm(q[3:7])
  L, nv = s.num_links(), s.nv
  ipos = s.link.inertia.transform.pos.astype(np.float64)
  irot = s.link.inertia.transform.rot.astype(np.float64)

  def poses(qq):
    st = o.init(qq[None], np.zeros((1, nv)))
    return st['x_pos'][0].copy(), st['x_rot'][0].copy(), st

  def move(qq, d, h):
    e = np.zeros(nv); e[d] = 1
    q2 = qq.copy()
    q2[0:3] += e[0:3] * h
    q2[3:7] = _qmul(qq[3:7], np.concatenate([[1.0], 0.5 * e[3:6] * h])); q2[3:7] /= np.linalg.norm(q2[3:7])
    q2[7:] += e[6:] * h
    return q2

  xp, xr, st = poses(q)
  coms = lambda p, r: np.array([p[l] + _qmat(r[l]) @ ipos[l] for l in range(L)])
  eps = 1e-6
  jv, jw = np.zeros((L, 3, nv)), np.zeros((L, 3, nv))
  for d in range(nv):
    p2, r2, _ = poses(move(q, d, eps)); p1, r1, _ = poses(move(q, d, -eps))
    jv[:, :, d] = (coms(p2, r2) - coms(p1, r1)) / (2 * eps)
    for l in range(L):
      w = (_qmat(r2[l]) - _qmat(r1[l])) / (2 * eps) @ _qmat(xr[l]).T
      jw[l, :, d] = [w[2, 1], w[0, 2], w[1, 0]]
  mn = np.diag(s.dof.armature.astype(np.float64))
  for l in range(L):
    r = _qmat(_qmul(xr[l], irot[l]))
    mn += float(s.link.inertia.mass[l]) * jv[l].T @ jv[l] + jw[l].T @ (r @ s.link.inertia.i[l].astype(np.float64) @ r.T) @ jw[l]
  m = st['mass_mx'][0]
  assert np.abs(m - m.T).max() == 0
  assert np.linalg.eigvalsh(m).min() > 0
  assert np.abs(m - mn).max() / np.abs(m).max() < 1e-6
  # init solves exactly (mass.py:103-104)
  assert np.abs(m @ st['mass_mx_inv'][0] - np.eye(nv)).max() < 1e-9


def test_free_fall_matches_closed_form(ant):
  """No contact, zero action: the root follows z0 - g t^2/2 under semi-implicit Euler
  (checks gravity in RNE, Minv and the free-joint integrator)."""
  o = O.Oracle(ant, np.float64)
  q = ant.init_q.astype(np.float64).copy(); q[2] = 50.0
  st = o.init(q[None], np.zeros((1, ant.nv)))
  n, dt = 100, float(ant.opt.timestep)
  com0 = st['root_com'][0, 0].copy()
  for _ in range(n):
    o.step(st, np.zeros((1, ant.nu)), 1)
  # semi-implicit Euler: z_n = z0 - g dt^2 n(n+1)/2 for the centre of mass
  expect = com0[2] - 9.81 * dt * dt * n * (n + 1) / 2
  assert abs(st['root_com'][0, 0, 2] - expect) < 1e-6
  assert np.abs(st['root_com'][0, 0, :2] - com0[:2]).max() < 1e-9


def test_pendulum_energy_drift_is_first_order():
  """Frictionless triple pendulum: energy drift shrinks with dt (pins the
  Coriolis/centrifugal terms of RNE: a sign error there breaks conservation)."""
  base = golden('triple_pendulum')
  drifts = []
  for dt in (2e-3, 1e-3):
    s = base.tree_replace({'opt.timestep': np.float32(dt)}).replace(matrix_inv_iterations=0)
    o = O.Oracle(s, np.float64)
    st = o.init(np.array([[0.8, -0.5, 0.3]]), np.zeros((1, 3)))

    def energy(st):
      m = st['mass_mx'][0]; qd = st['qd'][0]
      ke = 0.5 * qd @ m @ qd
      xi = st['cinr_pos'][0] / st['cinr_mass'][0][:, None] + st['root_com'][0]
      pe = 9.81 * float((st['cinr_mass'][0] * xi[:, 2]).sum())
      return ke + pe
    e0 = energy(st)
    worst = 0.0
    for _ in range(int(round(1.0 / dt))):
      o.step(st, np.zeros((1, 0)), 1)
      worst = max(worst, abs(energy(st) - e0))
    drifts.append(worst)
  # O(dt) integrator: halving dt halves the drift (a wrong bias force would leave
  # an O(1) drift that does not shrink)
  assert drifts[0] < 1.0
  assert 0.4 < drifts[1] / drifts[0] < 0.6


def test_static_contact_supports_weight(ant):
  """Ant at rest with a converged solver: contact + limit forces balance gravity,
  penetration stays sub-millimetre (pins contact rows, aref, PG and J^T x)."""
  s = ant.replace(solver_iterations=500)
  o = O.Oracle(s, np.float64)
  st = o.init(s.init_q.astype(np.float64)[None], np.zeros((1, s.nv)))
  for _ in range(500):
    o.step(st, np.zeros((1, s.nu)), 1)
  assert np.abs(st['qd'][0]).max() < 5e-2   # soft constraints creep slowly
  weight = 9.81 * float(s.link.inertia.mass.sum())
  assert abs(st['qf_constraint'][0, 2] - weight) / weight < 2e-2
  assert (st['con_dist'][0] < 0).all() and (st['con_dist'][0] > -1e-3).all()
  lo, hi = s.dof.limit
  qj = st['q'][0, 7:]
  assert (qj > lo[6:] - 5e-3).all() and (qj < hi[6:] + 5e-3).all()


def test_projected_gradient_properties():
  """jaxopt.ProjectedGradient restatement: x >= 0, converges to the NNLS optimum
  of 0.5||Ax+b||^2, stops early once the fixed-point error <= 1e-3."""
  rng = np.random.default_rng(3)
  for n in (4, 11, 24):
    g = rng.standard_normal((n, n))
    a = g @ g.T / n + 0.5 * np.eye(n)
    b = rng.standard_normal(n)
    x, stats = O.pg_solve(a, b, 2000, 15, np.float64)
    assert (x >= 0).all()
    grad = a.T @ (a @ x + b)
    # KKT of min f s.t. x>=0: grad >= 0 where x == 0, grad == 0 where x > 0
    assert np.linalg.norm(np.maximum(x - grad, 0) - x) <= 1e-3 + 1e-12
    assert stats[0] < 2000          # early exit happened
    x4, s4 = O.pg_solve(a, b, 4, 15, np.float32)
    assert s4[0] <= 4 and (x4 >= 0).all()
    f = lambda v: 0.5 * np.sum((a @ v + b) ** 2)
    assert f(x4.astype(np.float64)) <= f(np.zeros(n)) + 1e-9
  x0, s0 = O.pg_solve(np.eye(3), np.ones(3), 0, 15)
  assert (x0 == 0).all() and s0[0] == 0  # maxiter == 0 returns the initial point


def test_newton_schulz_accept_reject_and_cold_start():
  """math.inv_approximate semantics (math.py:278-305)."""
  rng = np.random.default_rng(0)
  g = rng.standard_normal((6, 6)); a = g @ g.T + np.eye(6)
  inv = np.linalg.inv(a)
  # warm start close to the answer converges in a few iterations
  x = O.inv_approximate(a, inv * 1.01, 5, np.float64)
  assert np.abs(x - inv).max() < 1e-10
  # ||I - A X0|| > 1 triggers the cold start 0.5 A^T / tr(A A^T); 0 iterations returns it
  x = O.inv_approximate(a, np.zeros((6, 6)), 0, np.float64)
  np.testing.assert_allclose(x, 0.5 * a.T / np.trace(a @ a.T), rtol=1e-12)
  # an exact inverse is left untouched (error cannot decrease below the start error of 1)
  x = O.inv_approximate(a, inv, 3, np.float64)
  assert np.abs(x - inv).max() < 1e-12
