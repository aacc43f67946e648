"""TEST-ONLY front-end of tests/simt/bxg_sim.cpp (host emulation of the kernel's
lane-group execution; see the header of that file).  Not a CPU fallback."""
import ctypes
import os
import subprocess

import numpy as np

from brax_b200 import native

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, 'libbxg_sim.so')
_SO64 = os.path.join(_HERE, 'libbxg_sim_f64.so')   # the same source with BXG_REAL = double (kernel logic vs the float64 goldens)
_SRC = os.path.join(_HERE, 'bxg_sim.cpp')
_CSRC = os.path.join(os.path.dirname(os.path.dirname(_HERE)), 'brax_b200', 'csrc')


def build(force=False):
  deps = [_SRC] + [os.path.join(_CSRC, f) for f in ('bxg_core.cuh', 'bxg_model.h')]
  # explicit fmaf() calls become one instruction where the CPU has FMA (same result as libm's software fmaf)
  try:
    hw_fma = ' fma ' in open('/proc/cpuinfo').read()
  except OSError:
    hw_fma = False
  jobs = []
  for so, extra in ((_SO, []), (_SO64, ['-DBXG_REAL=double', '-DBXG_SIM_F64'])):
    if not force and os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(d) for d in deps):
      continue
    jobs.append(subprocess.Popen(['g++', '-O2', '-ffp-contract=off', '-fPIC', '-shared', '-std=c++17', '-Wno-unknown-pragmas'] + (['-mfma'] if hw_fma else []) + extra + [_SRC, '-o', so]))
  for j in jobs:
    if j.wait() != 0:
      raise RuntimeError('g++ failed for the host emulator')


class Sim:
  def __init__(self, sys, variant=-1, reverse=False, minv_mode=native.MINV_NEWTON_SCHULZ, generic=False, dtype=np.float32):
    """variant: -1 auto, 0 = (G16, nv<=16, nc<=24), 1 = (G32, nv<=24, nc<=28),
    2 = (G32, nv<=32, nc<=32), 3 = generic kernel, 4 = (G16, nv<=24, nc<=28), 5 = (G32, nv<=16, nc<=64)."""
    build()
    self.dtype = np.dtype(dtype)
    self.lib = ctypes.CDLL(_SO if self.dtype == np.float32 else _SO64)
    assert self.lib.sim_sizeof_real() == self.dtype.itemsize
    self.sys = sys
    self.desc, self._keep = native.make_desc(sys, minv_mode)
    self.G, self.reverse = int(variant), int(reverse) | (2 if generic else 0)
    self.shapes = native.state_shapes(sys)
    self.ncon = len(sys.contact_pairs().geom1)

  def alloc(self, n):
    return {k: np.zeros((n,) + s, self.dtype) for k, s in self.shapes.items()}

  @staticmethod
  def _cstate(b, dtype=np.float32):
    cs = native.StateC()
    for f in native.STATE_FIELDS:
      assert b[f].dtype == dtype and b[f].flags['C_CONTIGUOUS'], f
      setattr(cs, f, b[f].ctypes.data)
    return cs

  def init(self, q, qd):
    q = np.ascontiguousarray(np.atleast_2d(q), self.dtype); qd = np.ascontiguousarray(np.atleast_2d(qd), self.dtype)
    n = q.shape[0]
    out = self.alloc(n)
    cs = self._cstate(out, self.dtype)
    rc = self.lib.sim_init(ctypes.byref(self.desc), self.G, self.reverse, ctypes.c_int64(n),
                           q.ctypes.data_as(ctypes.c_void_p), qd.ctypes.data_as(ctypes.c_void_p), ctypes.byref(cs))
    assert rc == 0, rc
    return out

  def step(self, st, act, n_frames=1, diag=False, lean=False):
    n = st['q'].shape[0]
    act = np.ascontiguousarray(np.atleast_2d(act), self.dtype)
    if lean:   # the other leaves are neither read nor written: poison them so that a stray read shows
      st = {k: (np.ascontiguousarray(st[k], self.dtype) if k in native.LEAN_FIELDS else np.full((n,) + self.shapes[k], np.nan, self.dtype)) for k in native.STATE_FIELDS}
    else:
      st = {k: np.ascontiguousarray(st[k], self.dtype) for k in native.STATE_FIELDS}
    out = self.alloc(n)
    if lean:
      for k in out:
        out[k][...] = np.nan
    cin, cout = self._cstate(st, self.dtype), self._cstate(out, self.dtype)
    dg = native.DiagC()
    con_dist = np.zeros((n, max(self.ncon, 1)), np.float32); stats = np.zeros((n, 4), np.int32)
    dg.con_dist = con_dist.ctypes.data; dg.stats = stats.ctypes.data
    rc = self.lib.sim_step(ctypes.byref(self.desc), self.G, self.reverse, ctypes.c_int64(n), int(n_frames),
                           ctypes.byref(cin), act.ctypes.data_as(ctypes.c_void_p), ctypes.byref(cout),
                           (1 if diag else 0) | (native.STEP_LEAN if lean else 0), ctypes.byref(dg))
    assert rc == 0, rc
    if diag:
      out['con_dist'] = con_dist; out['stats'] = stats
    return out


class SimEnv:
  """Host-emulated fused env step (bxg_env_reset / bxg_env_step semantics)."""

  def __init__(self, env, variant=-1, reverse=False):
    """env: a brax_b200.envs FusedEnv-like object exposing .sys, .spec, .n_frames (no CUDA is touched)."""
    self.sim = Sim(env.sys, variant=variant, reverse=reverse)
    self.spec, self.n_frames = env.spec, env.n_frames
    self.sys = env.sys
    L, nq, nv = env.sys.num_links(), env.sys.nq, env.sys.nv
    base = nq - env.spec.obs_skip + nv
    self.obs_size = base + 16 * L + nv if env.spec.kind in (native.ENV_COM_VELOCITY, native.ENV_STANDUP) else base
    if env.spec.kind == native.ENV_DOUBLE_CARTPOLE:
      self.obs_size = 1 + 2 * (nq - 1) + nv
    if env.spec.kind == native.ENV_PUSHER:
      self.obs_size = 2 * env.sys.nu + 9
    if env.spec.kind == native.ENV_REACHER:
      self.obs_size = 4 + (nq - 2) + 2 + 3

  def reset(self, q, qd):
    q = np.ascontiguousarray(q, np.float32); qd = np.ascontiguousarray(qd, np.float32)
    n = q.shape[0]
    out = self.sim.alloc(n)
    obs = np.zeros((n, self.obs_size), np.float32)
    cs = Sim._cstate(out)
    rc = self.sim.lib.sim_env_reset(ctypes.byref(self.sim.desc), self.sim.G, self.sim.reverse, ctypes.byref(self.spec),
                                    ctypes.c_int64(n), q.ctypes.data_as(ctypes.c_void_p), qd.ctypes.data_as(ctypes.c_void_p),
                                    ctypes.byref(cs), obs.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0, rc
    return out, obs

  def step(self, st, action, done, steps, first=None, first_obs=None):
    n = st['q'].shape[0]
    action = np.ascontiguousarray(action, np.float32)
    out = self.sim.alloc(n)
    io = {'obs': np.zeros((n, self.obs_size), np.float32), 'reward': np.zeros(n, np.float32),
          'done': np.ascontiguousarray(done, np.float32).copy(), 'metrics': np.zeros((n, native.ENV_NUM_METRICS), np.float32),
          'steps': np.ascontiguousarray(steps, np.float32).copy(), 'truncation': np.zeros(n, np.float32)}
    eio = native.EnvIOC()
    for k, v in io.items():
      setattr(eio, k, v.ctypes.data)
    keep = None
    if first is not None:
      keep = Sim._cstate(first)
      eio.first_state = ctypes.pointer(keep)
      first_obs = np.ascontiguousarray(first_obs, np.float32)
      eio.first_obs = first_obs.ctypes.data
    cin, cout = Sim._cstate(st), Sim._cstate(out)
    rc = self.sim.lib.sim_env_step(ctypes.byref(self.sim.desc), self.sim.G, self.sim.reverse, ctypes.byref(self.spec),
                                   ctypes.c_int64(n), int(self.n_frames), ctypes.byref(cin),
                                   action.ctypes.data_as(ctypes.c_void_p), ctypes.byref(cout), ctypes.byref(eio))
    assert rc == 0, rc
    return out, io
