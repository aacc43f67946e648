"""ctypes front-end for the C oracle (oracle/bxg_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(brax_b200/) must never import this module.

Parity status: see the header of bxg_oracle.c (pinned to 1e-9 against the reference's own
source run on NumPy, tests/test_reference_golden.py; "parity unpinned" only for jaxopt's
solver, the mjx colliders and MuJoCo's model compiler, which are not installable here).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Dict, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, '_build')
MAXL, MAXQ, MAXV, MAXU, MAXCON = 32, 96, 64, 64, 16
MAXC = 4 * MAXCON + MAXV

STATE_FIELDS = (
    'q', 'qd', 'x_pos', 'x_rot', 'xd_ang', 'xd_vel', 'root_com',
    'cinr_pos', 'cinr_rot', 'cinr_i', 'cinr_mass', 'cd_ang', 'cd_vel',
    'cdof_ang', 'cdof_vel', 'cdofd_ang', 'cdofd_vel',
    'mass_mx', 'mass_mx_inv', 'con_jac', 'con_diag', 'con_aref',
    'qf_smooth', 'qf_constraint', 'qdd')


def build(force: bool = False) -> None:
  """Compiles the oracle twice (float / double). gcc only, a few seconds."""
  os.makedirs(_BUILD, exist_ok=True)
  src = os.path.join(_HERE, 'bxg_oracle.c')
  for name, real in (('f32', 'float'), ('f64', 'double')):
    out = os.path.join(_BUILD, f'liboracle_{name}.so')
    if (not force and os.path.exists(out)
        and os.path.getmtime(out) >= os.path.getmtime(src)):
      continue
    cmd = ['gcc', '-O2', '-ffp-contract=off', '-fopenmp', '-fPIC', '-shared',
           f'-DORC_REAL={real}', src, '-o', out, '-lm']
    subprocess.run(cmd, check=True)


def _model_struct(ct):
  i32 = ctypes.c_int32
  return [
      ('num_links', i32), ('nq', i32), ('nv', i32), ('nu', i32), ('ncon', i32), ('nlim', i32),
      ('solver_iterations', i32), ('solver_maxls', i32), ('matrix_inv_iterations', i32), ('pad', i32),
      ('dt', ct), ('gravity', ct * 3),
      ('link_parent', i32 * MAXL), ('link_ndof', i32 * MAXL),
      ('link_q_adr', i32 * MAXL), ('link_qd_adr', i32 * MAXL),
      ('link_tf_pos', ct * (MAXL * 3)), ('link_tf_rot', ct * (MAXL * 4)), ('link_joint_pos', ct * (MAXL * 3)),
      ('inertia_pos', ct * (MAXL * 3)), ('inertia_rot', ct * (MAXL * 4)), ('inertia_i', ct * (MAXL * 9)),
      ('inertia_mass', ct * MAXL), ('link_invweight', ct * MAXL),
      ('dof_link', i32 * MAXV), ('dof_ang', ct * (MAXV * 3)), ('dof_vel', ct * (MAXV * 3)),
      ('dof_armature', ct * MAXV), ('dof_stiffness', ct * MAXV), ('dof_damping', ct * MAXV),
      ('dof_limit_lo', ct * MAXV), ('dof_limit_hi', ct * MAXV), ('dof_invweight', ct * MAXV),
      ('dof_solver_params', ct * (MAXV * 7)),
      ('act_q_id', i32 * MAXU), ('act_qd_id', i32 * MAXU),
      ('act_gain', ct * MAXU), ('act_gear', ct * MAXU), ('act_ctrl_lo', ct * MAXU), ('act_ctrl_hi', ct * MAXU),
      ('act_force_lo', ct * MAXU), ('act_force_hi', ct * MAXU), ('act_bias_q', ct * MAXU), ('act_bias_qd', ct * MAXU),
      ('con_link_a', i32 * MAXCON), ('con_link_b', i32 * MAXCON),
      ('con_plane_pos', ct * (MAXCON * 3)), ('con_frame', ct * (MAXCON * 9)), ('con_sphere_pos', ct * (MAXCON * 3)),
      ('con_radius', ct * MAXCON), ('con_friction', ct * MAXCON), ('con_solref', ct * (MAXCON * 2)),
      ('con_solimp', ct * (MAXCON * 5)),
      ('con_kind', i32 * MAXCON), ('con_geom_quat', ct * (MAXCON * 4)), ('con_half_len', ct * MAXCON),
      ('enable_fluid', i32), ('pad2', i32), ('viscosity', ct), ('density', ct),
      ('con_a_pos', ct * (MAXCON * 3)), ('con_a_quat', ct * (MAXCON * 4)), ('con_a_half', ct * MAXCON), ('con_a_radius', ct * MAXCON),
  ]


class Oracle:
  """One precision of the oracle bound to one System."""

  def __init__(self, sys, dtype=np.float32, threads: Optional[int] = None):
    build()
    self.dtype = np.dtype(dtype)
    name = 'f32' if self.dtype == np.float32 else 'f64'
    self.lib = ctypes.CDLL(os.path.join(_BUILD, f'liboracle_{name}.so'))
    self.ct = ctypes.c_float if self.dtype == np.float32 else ctypes.c_double
    assert self.lib.orc_sizeof_real() == self.dtype.itemsize

    class Model(ctypes.Structure):
      _fields_ = _model_struct(self.ct)

    class State(ctypes.Structure):
      _fields_ = [(f, ctypes.c_void_p) for f in STATE_FIELDS] + [
          ('con_dist', ctypes.c_void_p), ('stats', ctypes.c_void_p)]

    assert ctypes.sizeof(Model) == self.lib.orc_sizeof_model(), (
        ctypes.sizeof(Model), self.lib.orc_sizeof_model())
    self._State = State
    self.sys = sys
    self.model = Model()
    self._fill(sys)
    # explicit thread count (never inherited from OMP_NUM_THREADS, which launchers set to 1)
    self.threads = int(self.lib.orc_set_threads(int(threads) if threads is not None else (os.cpu_count() or 1)))
    self.lib.orc_step.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_int,
                                  ctypes.c_void_p, ctypes.c_void_p]
    self.lib.orc_init.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p,
                                  ctypes.c_void_p, ctypes.c_void_p]

  # ---------------------------------------------------------------- model
  def _fill(self, sys):
    m = self.model
    L = sys.num_links()
    cp = sys.contact_pairs()
    ncon = len(cp.geom1)
    nonfree = int(sum(int(t) for t in sys.link_types if t != 'f'))
    m.num_links, m.nq, m.nv, m.nu = L, sys.nq, sys.nv, sys.nu
    m.ncon = ncon
    m.nlim = nonfree if sys.dof.limit is not None else 0
    m.solver_iterations = int(sys.solver_iterations)
    m.solver_maxls = int(sys.solver_maxls)
    m.matrix_inv_iterations = int(sys.matrix_inv_iterations)
    m.dt = float(sys.opt.timestep)
    m.enable_fluid = int(bool(sys.enable_fluid))
    m.viscosity, m.density = float(np.asarray(sys.viscosity)), float(np.asarray(sys.density))

    def put(name, arr):
      a = np.asarray(arr).reshape(-1)
      dst = getattr(m, name)
      assert len(a) <= len(dst), name
      for i, v in enumerate(a):
        dst[i] = v.item() if hasattr(v, 'item') else v

    put('gravity', sys.gravity)
    qa, da = 0, 0
    parents, ndofs, qadr, dadr, dof_link = [], [], [], [], []
    for i, t in enumerate(sys.link_types):
      parents.append(sys.link_parents[i])
      nd = 0 if t == 'f' else int(t)
      ndofs.append(nd); qadr.append(qa); dadr.append(da)
      w = 6 if t == 'f' else nd
      dof_link.extend([i] * w)
      qa += 7 if t == 'f' else nd
      da += w
    put('link_parent', np.array(parents, np.int32)); put('link_ndof', np.array(ndofs, np.int32))
    put('link_q_adr', np.array(qadr, np.int32)); put('link_qd_adr', np.array(dadr, np.int32))
    put('dof_link', np.array(dof_link, np.int32))
    put('link_tf_pos', sys.link.transform.pos); put('link_tf_rot', sys.link.transform.rot)
    put('link_joint_pos', sys.link.joint.pos)
    put('inertia_pos', sys.link.inertia.transform.pos); put('inertia_rot', sys.link.inertia.transform.rot)
    put('inertia_i', sys.link.inertia.i); put('inertia_mass', sys.link.inertia.mass)
    put('link_invweight', sys.link.invweight)
    put('dof_ang', sys.dof.motion.ang); put('dof_vel', sys.dof.motion.vel)
    put('dof_armature', sys.dof.armature); put('dof_stiffness', sys.dof.stiffness)
    put('dof_damping', sys.dof.damping); put('dof_invweight', sys.dof.invweight)
    put('dof_solver_params', sys.dof.solver_params)
    if sys.dof.limit is not None:
      put('dof_limit_lo', sys.dof.limit[0]); put('dof_limit_hi', sys.dof.limit[1])
    if sys.nu:
      a = sys.actuator
      put('act_q_id', a.q_id); put('act_qd_id', a.qd_id)
      put('act_gain', a.gain); put('act_gear', a.gear)
      put('act_ctrl_lo', a.ctrl_range[:, 0]); put('act_ctrl_hi', a.ctrl_range[:, 1])
      put('act_force_lo', a.force_range[:, 0]); put('act_force_hi', a.force_range[:, 1])
      put('act_bias_q', a.bias_q); put('act_bias_qd', a.bias_qd)
    if ncon:
      put('con_link_a', cp.link_a); put('con_link_b', cp.link_b)
      put('con_plane_pos', cp.plane_pos); put('con_frame', cp.frame)
      put('con_sphere_pos', cp.sphere_pos); put('con_radius', cp.radius)
      put('con_friction', cp.friction); put('con_solref', cp.solref); put('con_solimp', cp.solimp)
      put('con_kind', cp.kind); put('con_geom_quat', cp.geom_quat); put('con_half_len', cp.half_len)
      put('con_a_pos', cp.a_pos); put('con_a_quat', cp.a_quat); put('con_a_half', cp.a_half); put('con_a_radius', cp.a_radius)

  # ---------------------------------------------------------------- state
  def shapes(self) -> Dict[str, tuple]:
    m = self.model
    L, nq, nv = m.num_links, m.nq, m.nv
    nc = 4 * m.ncon + m.nlim
    return {
        'q': (nq,), 'qd': (nv,), 'x_pos': (L, 3), 'x_rot': (L, 4), 'xd_ang': (L, 3), 'xd_vel': (L, 3),
        'root_com': (L, 3), 'cinr_pos': (L, 3), 'cinr_rot': (L, 4), 'cinr_i': (L, 3, 3), 'cinr_mass': (L,),
        'cd_ang': (L, 3), 'cd_vel': (L, 3), 'cdof_ang': (nv, 3), 'cdof_vel': (nv, 3),
        'cdofd_ang': (nv, 3), 'cdofd_vel': (nv, 3), 'mass_mx': (nv, nv), 'mass_mx_inv': (nv, nv),
        'con_jac': (nc, nv), 'con_diag': (nc,), 'con_aref': (nc,),
        'qf_smooth': (nv,), 'qf_constraint': (nv,), 'qdd': (nv,)}

  def alloc(self, n: int) -> Dict[str, np.ndarray]:
    st = {k: np.zeros((n,) + s, self.dtype) for k, s in self.shapes().items()}
    st['con_dist'] = np.zeros((n, max(self.model.ncon, 1)), self.dtype)
    st['stats'] = np.zeros((n, 4), np.int32)
    return st

  def _cstate(self, st):
    cs = self._State()
    for f in STATE_FIELDS + ('con_dist', 'stats'):
      a = st[f]
      assert a.flags['C_CONTIGUOUS']
      assert a.dtype == (np.int32 if f == 'stats' else self.dtype), f
      setattr(cs, f, a.ctypes.data)
    return cs

  def init(self, q: np.ndarray, qd: np.ndarray) -> Dict[str, np.ndarray]:
    """pipeline.init over a batch. q [n,nq], qd [n,nv]."""
    q = np.ascontiguousarray(np.atleast_2d(q), self.dtype)
    qd = np.ascontiguousarray(np.atleast_2d(qd), self.dtype)
    n = q.shape[0]
    st = self.alloc(n)
    cs = self._cstate(st)
    rc = self.lib.orc_init(ctypes.byref(self.model), n, q.ctypes.data, qd.ctypes.data, ctypes.byref(cs))
    assert rc == 0
    return st

  def step(self, st: Dict[str, np.ndarray], act: np.ndarray, n_frames: int = 1):
    """n_frames x pipeline.step in place; returns st."""
    n = st['q'].shape[0]
    act = np.ascontiguousarray(np.atleast_2d(act), self.dtype)
    if self.model.nu:
      assert act.shape == (n, self.model.nu), act.shape
    cs = self._cstate(st)
    rc = self.lib.orc_step(ctypes.byref(self.model), n, n_frames, ctypes.byref(cs), act.ctypes.data)
    assert rc == 0
    return st

  # ------------------------------------------------------------ components
  def to_tau(self, q, qd, act):
    q = np.ascontiguousarray(q, self.dtype); qd = np.ascontiguousarray(qd, self.dtype)
    act = np.ascontiguousarray(act, self.dtype)
    tau = np.zeros(self.model.nv, self.dtype)
    self.lib.orc_to_tau(ctypes.byref(self.model), q.ctypes.data_as(ctypes.c_void_p),
                        qd.ctypes.data_as(ctypes.c_void_p), act.ctypes.data_as(ctypes.c_void_p),
                        tau.ctypes.data_as(ctypes.c_void_p))
    return tau

  def contact(self, q):
    q = np.ascontiguousarray(q, self.dtype)
    dist = np.zeros(self.model.ncon, self.dtype)
    pos = np.zeros((self.model.ncon, 3), self.dtype)
    self.lib.orc_contact(ctypes.byref(self.model), q.ctypes.data_as(ctypes.c_void_p),
                         dist.ctypes.data_as(ctypes.c_void_p), pos.ctypes.data_as(ctypes.c_void_p))
    return dist, pos


def inv_approximate(a, a_inv, num_iter, dtype=np.float32):
  build()
  dtype = np.dtype(dtype)
  lib = ctypes.CDLL(os.path.join(_BUILD, 'liboracle_%s.so' % ('f32' if dtype == np.float32 else 'f64')))
  a = np.ascontiguousarray(a, dtype); x = np.ascontiguousarray(a_inv, dtype).copy()
  lib.orc_inv_approximate(a.ctypes.data_as(ctypes.c_void_p), x.ctypes.data_as(ctypes.c_void_p),
                          a.shape[0], int(num_iter))
  return x


def pg_solve(A, b, maxiter, maxls, dtype=np.float32):
  build()
  dtype = np.dtype(dtype)
  lib = ctypes.CDLL(os.path.join(_BUILD, 'liboracle_%s.so' % ('f32' if dtype == np.float32 else 'f64')))
  A = np.ascontiguousarray(A, dtype); b = np.ascontiguousarray(b, dtype)
  x = np.zeros(b.shape[0], dtype); stats = np.zeros(2, np.int32)
  rc = lib.orc_pg_solve(A.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p),
                        b.shape[0], int(maxiter), int(maxls),
                        x.ctypes.data_as(ctypes.c_void_p), stats.ctypes.data_as(ctypes.c_void_p))
  assert rc == 0
  return x, stats
