"""ctypes binding of the C ABI in include/bxg.h (libbxg.so, built in-tree).

PyTorch is used only for device memory and streams: every State leaf is a
torch CUDA tensor whose `data_ptr()` is handed to the library.  There is no CPU
fallback -- if the shared library or a CUDA device is missing, calls raise.
"""
from __future__ import annotations

import ctypes
import os
import threading
from typing import Dict, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libbxg.so')

MINV_NEWTON_SCHULZ = 0
MINV_CHOLESKY = 1
STEP_DIAGNOSTICS = 1
STEP_LEAN = 2          # BXG_STEP_LEAN: q, qd, x, mass_mx_inv in; q, qd, x, xd, mass_mx_inv out
NUM_PHASES = 16
LEAN_FIELDS = ('q', 'qd', 'x_pos', 'x_rot', 'xd_ang', 'xd_vel', 'mass_mx_inv')

STATE_FIELDS = (
    'q', 'qd', 'x_pos', 'x_rot', 'xd_ang', 'xd_vel', 'root_com',
    'cinr_pos', 'cinr_rot', 'cinr_i', 'cinr_mass', 'cd_ang', 'cd_vel',
    'cdof_ang', 'cdof_vel', 'cdofd_ang', 'cdofd_vel',
    'mass_mx', 'mass_mx_inv', 'con_jac', 'con_diag', 'con_aref',
    'qf_smooth', 'qf_constraint', 'qdd')

_i32, _f32 = ctypes.c_int32, ctypes.c_float
_pi, _pf = ctypes.POINTER(_i32), ctypes.POINTER(_f32)


class ModelDesc(ctypes.Structure):
  """Mirror of BxgModelDesc (include/bxg.h)."""
  _fields_ = [
      ('abi_version', _i32), ('num_links', _i32), ('nq', _i32), ('nv', _i32), ('nu', _i32),
      ('ncon', _i32), ('has_limit', _i32),
      ('solver_iterations', _i32), ('solver_maxls', _i32), ('matrix_inv_iterations', _i32),
      ('minv_mode', _i32), ('dt', _f32), ('gravity', _f32 * 3),
      ('link_parent', _pi), ('link_ndof', _pi),
      ('link_tf_pos', _pf), ('link_tf_rot', _pf), ('link_joint_pos', _pf),
      ('inertia_pos', _pf), ('inertia_rot', _pf), ('inertia_i', _pf), ('inertia_mass', _pf),
      ('link_invweight', _pf),
      ('dof_ang', _pf), ('dof_vel', _pf), ('dof_armature', _pf), ('dof_stiffness', _pf),
      ('dof_damping', _pf), ('dof_limit_lo', _pf), ('dof_limit_hi', _pf), ('dof_invweight', _pf),
      ('dof_solver_params', _pf),
      ('act_q_id', _pi), ('act_qd_id', _pi), ('act_gain', _pf), ('act_gear', _pf),
      ('act_ctrl_lo', _pf), ('act_ctrl_hi', _pf), ('act_force_lo', _pf), ('act_force_hi', _pf),
      ('act_bias_q', _pf), ('act_bias_qd', _pf),
      ('con_link_a', _pi), ('con_link_b', _pi), ('con_plane_pos', _pf), ('con_frame', _pf),
      ('con_sphere_pos', _pf), ('con_radius', _pf), ('con_friction', _pf),
      ('con_solref', _pf), ('con_solimp', _pf),
      ('con_kind', _pi), ('con_geom_quat', _pf), ('con_half_len', _pf),
      ('enable_fluid', _i32), ('viscosity', _f32), ('density', _f32),
      ('con_a_pos', _pf), ('con_a_quat', _pf), ('con_a_half', _pf), ('con_a_radius', _pf),
  ]


class StateC(ctypes.Structure):
  _fields_ = [(f, ctypes.c_void_p) for f in STATE_FIELDS]


class DiagC(ctypes.Structure):
  _fields_ = [('con_dist', ctypes.c_void_p), ('stats', ctypes.c_void_p), ('phase_cycles', ctypes.c_void_p)]


ENV_ROOT_VELOCITY = 1
ENV_COM_VELOCITY = 2
ENV_PLANAR = 3
ENV_CARTPOLE = 4
ENV_DOUBLE_CARTPOLE = 5
ENV_REACHER = 6
ENV_SWIMMER = 7
ENV_STANDUP = 8
ENV_PUSHER = 9
ENV_NUM_METRICS = 10


class EnvSpecC(ctypes.Structure):
  """Mirror of BxgEnvSpec (include/bxg.h)."""
  _fields_ = [('kind', _i32), ('obs_skip', _i32), ('terminate_when_unhealthy', _i32), ('episode_length', _i32),
              ('forward_reward_weight', _f32), ('ctrl_cost_weight', _f32), ('healthy_reward', _f32),
              ('healthy_z_min', _f32), ('healthy_z_max', _f32), ('env_dt', _f32),
              ('healthy_angle_min', _f32), ('healthy_angle_max', _f32),
              ('healthy_state_min', _f32), ('healthy_state_max', _f32),
              ('tip_link', _i32), ('target_link', _i32), ('tip_pos', _f32 * 3), ('object_link', _i32)]


class EnvIOC(ctypes.Structure):
  """Mirror of BxgEnvIO (include/bxg.h)."""
  _fields_ = [('obs', ctypes.c_void_p), ('reward', ctypes.c_void_p), ('done', ctypes.c_void_p),
              ('metrics', ctypes.c_void_p), ('steps', ctypes.c_void_p), ('truncation', ctypes.c_void_p),
              ('first_state', ctypes.POINTER(StateC)), ('first_obs', ctypes.c_void_p), ('flags', _i32)]


def make_desc(sys, minv_mode: int = MINV_NEWTON_SCHULZ, cp=None, cache: Optional[dict] = None) -> Tuple[ModelDesc, list]:
  """System -> BxgModelDesc.  Returns (desc, keepalive arrays).  cp: `sys.contact_pairs()` if the caller has it.
  cache: conversions keyed by the identity of the source leaf, shared between the descs of a batched System (whose
  unmapped leaves are the same objects in every env's System): each shared leaf is converted once."""
  keep = []

  def conv(a, dtype, ptr_t, key):
    k = (id(a) if key is None else key, dtype)
    if cache is not None and k in cache:
      return cache[k][1]
    arr = np.ascontiguousarray(np.asarray(a, dtype=dtype).reshape(-1))
    if arr.size == 0:
      arr = np.zeros(1, dtype)
    p = arr.ctypes.data_as(ptr_t)
    if cache is not None:
      cache[k] = (a, p, arr)    # (the source object stays alive, so its id is not reused)
    keep.append(arr)
    return p

  def fp(a, key=None):
    return conv(a, np.float32, _pf, key)

  def ip(a, key=None):
    return conv(a, np.int32, _pi, key)

  cp = sys.contact_pairs() if cp is None else cp
  d = ModelDesc()
  d.abi_version = 4
  d.num_links, d.nq, d.nv, d.nu = sys.num_links(), sys.nq, sys.nv, sys.nu
  d.ncon = len(cp.geom1)
  d.has_limit = 0 if sys.dof.limit is None else 1
  d.solver_iterations = int(sys.solver_iterations)
  d.solver_maxls = int(sys.solver_maxls)
  d.matrix_inv_iterations = int(sys.matrix_inv_iterations)
  d.minv_mode = int(minv_mode)
  d.dt = float(sys.opt.timestep)
  for i in range(3):
    d.gravity[i] = float(sys.gravity[i])
  d.link_parent = ip(sys.link_parents, key=('parents', sys.link_parents))
  d.link_ndof = ip([0 if t == 'f' else int(t) for t in sys.link_types], key=('ndof', sys.link_types))
  d.link_tf_pos = fp(sys.link.transform.pos); d.link_tf_rot = fp(sys.link.transform.rot)
  d.link_joint_pos = fp(sys.link.joint.pos)
  d.inertia_pos = fp(sys.link.inertia.transform.pos); d.inertia_rot = fp(sys.link.inertia.transform.rot)
  d.inertia_i = fp(sys.link.inertia.i); d.inertia_mass = fp(sys.link.inertia.mass)
  d.link_invweight = fp(sys.link.invweight)
  d.dof_ang = fp(sys.dof.motion.ang); d.dof_vel = fp(sys.dof.motion.vel)
  d.dof_armature = fp(sys.dof.armature); d.dof_stiffness = fp(sys.dof.stiffness)
  d.dof_damping = fp(sys.dof.damping)
  if sys.dof.limit is not None:
    d.dof_limit_lo = fp(sys.dof.limit[0]); d.dof_limit_hi = fp(sys.dof.limit[1])
  d.dof_invweight = fp(sys.dof.invweight); d.dof_solver_params = fp(sys.dof.solver_params)
  a = sys.actuator
  d.act_q_id = ip(a.q_id); d.act_qd_id = ip(a.qd_id)
  d.act_gain = fp(a.gain); d.act_gear = fp(a.gear)
  cr = np.asarray(a.ctrl_range, np.float32).reshape(-1, 2)
  fr = np.asarray(a.force_range, np.float32).reshape(-1, 2)
  d.act_ctrl_lo = fp(cr[:, 0], key=('clo', id(a.ctrl_range))); d.act_ctrl_hi = fp(cr[:, 1], key=('chi', id(a.ctrl_range)))
  d.act_force_lo = fp(fr[:, 0], key=('flo', id(a.force_range))); d.act_force_hi = fp(fr[:, 1], key=('fhi', id(a.force_range)))
  d.act_bias_q = fp(a.bias_q); d.act_bias_qd = fp(a.bias_qd)
  d.con_link_a = ip(cp.link_a); d.con_link_b = ip(cp.link_b)
  d.con_plane_pos = fp(cp.plane_pos); d.con_frame = fp(cp.frame)
  d.con_sphere_pos = fp(cp.sphere_pos); d.con_radius = fp(cp.radius)
  d.con_friction = fp(cp.friction); d.con_solref = fp(cp.solref); d.con_solimp = fp(cp.solimp)
  d.con_kind = ip(cp.kind); d.con_geom_quat = fp(cp.geom_quat); d.con_half_len = fp(cp.half_len)
  d.enable_fluid = int(bool(sys.enable_fluid))
  d.viscosity, d.density = float(np.asarray(sys.viscosity)), float(np.asarray(sys.density))
  d.con_a_pos = fp(cp.a_pos); d.con_a_quat = fp(cp.a_quat); d.con_a_half = fp(cp.a_half); d.con_a_radius = fp(cp.a_radius)
  return d, keep


def num_constraints(sys) -> int:
  ncon = len(sys.contact_pairs().geom1)
  nlim = 0
  if sys.dof.limit is not None:
    nlim = sum(int(t) for t in sys.link_types if t != 'f')
  return 4 * ncon + nlim


def state_shapes(sys) -> Dict[str, tuple]:
  L, nq, nv, nc = sys.num_links(), sys.nq, sys.nv, num_constraints(sys)
  return {
      'q': (nq,), 'qd': (nv,), 'x_pos': (L, 3), 'x_rot': (L, 4), 'xd_ang': (L, 3), 'xd_vel': (L, 3),
      'root_com': (L, 3), 'cinr_pos': (L, 3), 'cinr_rot': (L, 4), 'cinr_i': (L, 3, 3), 'cinr_mass': (L,),
      'cd_ang': (L, 3), 'cd_vel': (L, 3), 'cdof_ang': (nv, 3), 'cdof_vel': (nv, 3),
      'cdofd_ang': (nv, 3), 'cdofd_vel': (nv, 3), 'mass_mx': (nv, nv), 'mass_mx_inv': (nv, nv),
      'con_jac': (nc, nv), 'con_diag': (nc,), 'con_aref': (nc,),
      'qf_smooth': (nv,), 'qf_constraint': (nv,), 'qdd': (nv,)}


_lib = None
_lib_lock = threading.Lock()


def lib() -> ctypes.CDLL:
  """Loads libbxg.so.  Raises (loudly) if the extension has not been built."""
  global _lib
  with _lib_lock:
    if _lib is None:
      path = os.environ.get('BXG_LIB', LIB_PATH)   # tuning: an alternative build of the same library
      if not os.path.exists(path):
        raise RuntimeError(
            f'{path} is missing: build it with `python -c "import '
            '__graft_entry__ as g; g.build()"`. There is no CPU fallback.')
      l = ctypes.CDLL(path)
      l.bxg_last_error.restype = ctypes.c_char_p
      l.bxg_launch_count.restype = ctypes.c_int64
      l.bxg_model_create.argtypes = [ctypes.POINTER(ModelDesc), ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
      l.bxg_model_destroy.argtypes = [ctypes.c_void_p]
      l.bxg_model_num_constraints.argtypes = [ctypes.c_void_p]
      l.bxg_model_kernel_id.argtypes = [ctypes.c_void_p]
      l.bxg_model_create_batched.argtypes = [ctypes.POINTER(ModelDesc), ctypes.c_int64, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
      l.bxg_model_num_models.argtypes = [ctypes.c_void_p]
      l.bxg_model_num_models.restype = ctypes.c_int64
      l.bxg_plan.argtypes = [ctypes.POINTER(ModelDesc), ctypes.POINTER(ctypes.c_int32)]
      l.bxg_launch_shape.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_int32)]
      l.bxg_init.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                             ctypes.POINTER(StateC), ctypes.c_void_p]
      l.bxg_step.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.POINTER(StateC),
                             ctypes.c_void_p, ctypes.POINTER(StateC), ctypes.c_int32,
                             ctypes.POINTER(DiagC), ctypes.c_void_p]
      l.bxg_env_obs_size.argtypes = [ctypes.c_void_p, ctypes.POINTER(EnvSpecC)]
      l.bxg_env_reset.argtypes = [ctypes.c_void_p, ctypes.POINTER(EnvSpecC), ctypes.c_int64, ctypes.c_void_p,
                                  ctypes.c_void_p, ctypes.POINTER(StateC), ctypes.c_void_p, ctypes.c_void_p]
      l.bxg_env_step.argtypes = [ctypes.c_void_p, ctypes.POINTER(EnvSpecC), ctypes.c_int64, ctypes.c_int32,
                                 ctypes.POINTER(StateC), ctypes.c_void_p, ctypes.POINTER(StateC),
                                 ctypes.POINTER(EnvIOC), ctypes.c_void_p]
      _f, _vp = ctypes.c_float, ctypes.c_void_p
      l.bxg_gae.argtypes = [_vp, _vp, _vp, _vp, _vp, ctypes.c_int32, ctypes.c_int64, _f, _f, _vp, _vp, _vp]
      l.bxg_policy_act.argtypes = [_vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_int64, ctypes.c_int32,
                                   ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, _f, _vp, _vp, _vp, _vp]
      l.bxg_ppo_head.argtypes = [_vp] * 8 + [ctypes.c_int32, ctypes.c_int64, ctypes.c_int32, _f, _f, _f, _f, _f, ctypes.c_int32, _f] + [_vp] * 7
      if l.bxg_abi_version() != 4:
        raise RuntimeError('libbxg.so ABI version mismatch')
      _lib = l
    return _lib


def plan(sys, minv_mode: int = MINV_NEWTON_SCHULZ) -> dict:
  """Host-only: kernel variant and shared-memory footprint for a System."""
  desc, keep = make_desc(sys, minv_mode)
  info = (ctypes.c_int32 * 8)()
  _check(lib().bxg_plan(ctypes.byref(desc), info), 'bxg_plan')
  names = ('variant', 'lanes_per_env', 'model_words', 'env_words', 'envs_per_cta', 'smem_bytes_per_cta', 'nc', 'kernel_id')
  return dict(zip(names, list(info)[:8]))


def launch_count() -> int:
  return int(lib().bxg_launch_count())


def _check(rc: int, what: str):
  if rc != 0:
    msg = lib().bxg_last_error().decode('utf-8', 'replace')
    raise RuntimeError(f'{what} failed (code {rc}): {msg}')


class NativeModel:
  """A System uploaded to one CUDA device (BxgModel handle)."""

  def __init__(self, sys, device: int, minv_mode: int = MINV_NEWTON_SCHULZ):
    import torch
    if not torch.cuda.is_available():
      raise RuntimeError('no CUDA device: brax_b200 has no CPU fallback')
    self.sys = sys
    self.device = int(device)
    self.minv_mode = minv_mode
    self.shapes = state_shapes(sys)
    self.nc = num_constraints(sys)
    self.ncon = len(sys.contact_pairs().geom1)
    desc, keep = make_desc(sys, minv_mode)
    h = ctypes.c_void_p()
    _check(lib().bxg_model_create(ctypes.byref(desc), self.device, ctypes.byref(h)), 'bxg_model_create')
    self._h = h
    self.kernel_id = int(lib().bxg_model_kernel_id(h))
    del keep

  def __del__(self):
    h, self._h = getattr(self, '_h', None), None
    if h is not None and _lib is not None:
      _lib.bxg_model_destroy(h)

  def launch_shape(self, n: int) -> dict:
    """The launch a batch of n envs actually gets (grid, threads, shared memory, envs per CTA)."""
    info = (ctypes.c_int32 * 4)()
    _check(lib().bxg_launch_shape(self._h, n, info), 'bxg_launch_shape')
    return dict(zip(('grid', 'threads_per_cta', 'smem_bytes_per_cta', 'envs_per_cta'), list(info)))

  # -- buffers ---------------------------------------------------------------
  def alloc(self, n: int, lean: bool = False) -> Dict[str, 'torch.Tensor']:
    import torch
    dev = torch.device('cuda', self.device)
    return {k: torch.empty((n,) + s, dtype=torch.float32, device=dev) for k, s in self.shapes.items() if not lean or k in LEAN_FIELDS}

  def _cstate(self, bufs, lean: bool = False) -> StateC:
    cs = StateC()
    for f in (LEAN_FIELDS if lean else STATE_FIELDS):
      t = bufs[f]
      assert t.is_cuda and t.is_contiguous() and t.dtype.is_floating_point and t.element_size() == 4, f
      # the model's constants live on self.device and the kernel is launched there: a leaf on another GPU is a bug
      assert t.device.index == self.device, f'{f} is on cuda:{t.device.index}, the model on cuda:{self.device}'
      setattr(cs, f, t.data_ptr())
    return cs

  # -- calls -----------------------------------------------------------------
  def init(self, q, qd, out: Optional[dict] = None) -> dict:
    import torch
    n = q.shape[0]
    assert q.shape == (n, self.sys.nq) and qd.shape == (n, self.sys.nv), (q.shape, qd.shape)
    q = q.contiguous().float(); qd = qd.contiguous().float()
    out = self.alloc(n) if out is None else out
    cs = self._cstate(out)
    stream = torch.cuda.current_stream(q.device).cuda_stream
    _check(lib().bxg_init(self._h, n, q.data_ptr(), qd.data_ptr(), ctypes.byref(cs), stream), 'bxg_init')
    return out

  def step(self, bufs: dict, act, n_frames: int = 1, out: Optional[dict] = None,
           diag: Optional[dict] = None, lean: bool = False) -> dict:
    """lean: BXG_STEP_LEAN (include/bxg.h): `bufs` / `out` need only native.LEAN_FIELDS."""
    import torch
    n = bufs['q'].shape[0]
    if self.sys.nu:
      assert act is not None and act.shape == (n, self.sys.nu), None if act is None else act.shape
      act = act.contiguous().float()
      act_ptr = act.data_ptr()
    else:
      act_ptr = None
    out = self.alloc(n, lean) if out is None else out
    cin, cout = self._cstate(bufs, lean), self._cstate(out, lean)
    flags, dg = (STEP_LEAN if lean else 0), None
    if diag is not None:
      flags |= STEP_DIAGNOSTICS
      dg = DiagC()
      dg.con_dist = diag['con_dist'].data_ptr() if self.ncon else None
      dg.stats = diag['stats'].data_ptr()
      dg.phase_cycles = diag['phase_cycles'].data_ptr() if 'phase_cycles' in diag else None
    stream = torch.cuda.current_stream(bufs['q'].device).cuda_stream
    _check(lib().bxg_step(self._h, n, int(n_frames), ctypes.byref(cin), act_ptr, ctypes.byref(cout),
                          flags, ctypes.byref(dg) if dg is not None else None, stream), 'bxg_step')
    return out

  # -- fused env calls ---------------------------------------------------------
  def env_obs_size(self, spec: EnvSpecC) -> int:
    return int(lib().bxg_env_obs_size(self._h, ctypes.byref(spec)))

  def env_reset(self, spec: EnvSpecC, q, qd):
    """pipeline.init + observation of the fresh state.  Returns (state bufs, obs)."""
    import torch
    n = q.shape[0]
    q = q.contiguous().float(); qd = qd.contiguous().float()
    out = self.alloc(n)
    obs = torch.empty((n, self.env_obs_size(spec)), dtype=torch.float32, device=q.device)
    cs = self._cstate(out)
    stream = torch.cuda.current_stream(q.device).cuda_stream
    _check(lib().bxg_env_reset(self._h, ctypes.byref(spec), n, q.data_ptr(), qd.data_ptr(), ctypes.byref(cs),
                               obs.data_ptr(), stream), 'bxg_env_reset')
    return out, obs

  def env_step(self, spec: EnvSpecC, bufs: dict, action, n_frames: int, io: dict, out: Optional[dict] = None,
               first: Optional[dict] = None, first_obs=None, lean: bool = False) -> dict:
    """AutoReset(Episode(env)).step over the batch.  `io` holds the per-env arrays
    obs, reward, done, metrics (+ optional steps, truncation), updated in place."""
    import torch
    n = bufs['q'].shape[0]
    action = action.contiguous().float()
    assert action.shape == (n, self.sys.nu), action.shape
    out = self.alloc(n, lean) if out is None else out
    cin, cout = self._cstate(bufs, lean), self._cstate(out, lean)
    eio = EnvIOC()
    eio.flags = STEP_LEAN if lean else 0
    for k in ('obs', 'reward', 'done', 'metrics'):
      t = io[k]
      assert t.is_cuda and t.is_contiguous() and t.dtype == torch.float32, k
      setattr(eio, k, t.data_ptr())
    eio.steps = io['steps'].data_ptr() if io.get('steps') is not None else None
    eio.truncation = io['truncation'].data_ptr() if io.get('truncation') is not None else None
    cfirst = None
    if first is not None:
      cfirst = self._cstate(first, lean)
      eio.first_state = ctypes.pointer(cfirst)
      eio.first_obs = first_obs.data_ptr()
    stream = torch.cuda.current_stream(action.device).cuda_stream
    _check(lib().bxg_env_step(self._h, ctypes.byref(spec), n, int(n_frames), ctypes.byref(cin), action.data_ptr(),
                              ctypes.byref(cout), ctypes.byref(eio), stream), 'bxg_env_step')
    return out

  def alloc_diag(self, n: int) -> dict:
    import torch
    dev = torch.device('cuda', self.device)
    return {'con_dist': torch.zeros((n, max(self.ncon, 1)), dtype=torch.float32, device=dev),
            'stats': torch.zeros((n, 4), dtype=torch.int32, device=dev),
            'phase_cycles': torch.zeros(NUM_PHASES, dtype=torch.int64, device=dev)}


class BatchedNativeModel(NativeModel):
  """One System per env (bxg_model_create_batched): what the reference's DomainRandomizationVmapWrapper reaches with
  `jax.vmap` over a System with batched leaves (envs/wrappers/training.py:223-260).  `systems[e]` is env e's System;
  all share one topology.  Every call on this model takes exactly len(systems) envs."""

  def __init__(self, systems, device: int):
    import torch
    if not torch.cuda.is_available():
      raise RuntimeError('no CUDA device: brax_b200 has no CPU fallback')
    systems = list(systems)
    assert systems, 'no Systems'
    sys = systems[0]
    self.sys = sys
    self.systems = systems
    self.device = int(device)
    self.minv_mode = MINV_NEWTON_SCHULZ
    self.shapes = state_shapes(sys)
    self.nc = num_constraints(sys)
    self.ncon = len(sys.contact_pairs().geom1)
    descs = (ModelDesc * len(systems))()
    keep = []
    geom = ('geom_pos', 'geom_quat', 'geom_size', 'geom_friction', 'geom_solref', 'geom_solimp')
    cp0 = sys.contact_pairs()
    cache = {}
    for e, s in enumerate(systems):
      # the contact pairs derive from the geom leaves only: shared (the same arrays) unless those are randomised
      same_geoms = all(getattr(s, f) is getattr(sys, f) for f in geom)
      d, k = make_desc(s, MINV_NEWTON_SCHULZ, cp0 if same_geoms else None, cache)
      ctypes.memmove(ctypes.byref(descs, e * ctypes.sizeof(ModelDesc)), ctypes.byref(d), ctypes.sizeof(ModelDesc))
      keep.append(k)
    h = ctypes.c_void_p()
    _check(lib().bxg_model_create_batched(descs, len(systems), self.device, ctypes.byref(h)), 'bxg_model_create_batched')
    self._h = h
    self.kernel_id = int(lib().bxg_model_kernel_id(h))
    self.num_models = int(lib().bxg_model_num_models(h))
    del keep


def model_for(sys, device: int, minv_mode: int = MINV_NEWTON_SCHULZ) -> NativeModel:
  """Per-(System object, device, mode) cache of uploaded models.  The cache lives
  on the System object itself, so it is released with it and a `sys.replace(...)`
  copy (different constants) never sees a stale upload."""
  cache = getattr(sys, '_bxg_models', None)
  if cache is None:
    cache = {}
    object.__setattr__(sys, '_bxg_models', cache)   # System is a frozen dataclass
  key = (int(device), int(minv_mode))
  m = cache.get(key)
  if m is None:
    m = NativeModel(sys, device, minv_mode)
    cache[key] = m
  return m
