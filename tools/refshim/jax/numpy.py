"""jax.numpy on NumPy: same names, immutable-update `.at[]`, JAX's default dtypes
(float -> DEFAULT_FLOAT, int -> int32).  DEFAULT_FLOAT is float64 unless REFSHIM_F32=1."""
import os as _os

import numpy as _np

DEFAULT_FLOAT = _np.float32 if _os.environ.get('REFSHIM_F32') == '1' else _np.float64
pi, inf, newaxis = _np.pi, _np.inf, None
float32, float64, int32, bool_ = _np.float32, _np.float64, _np.int32, _np.bool_


class _AtIndex:
  def __init__(self, arr, idx):
    self.arr, self.idx = arr, idx

  def set(self, v):
    out = _np.array(self.arr, copy=True); out[self.idx] = v; return _wrap(out)

  def add(self, v):
    out = _np.array(self.arr, copy=True); _np.add.at(out, self.idx, v); return _wrap(out)

  def multiply(self, v):
    out = _np.array(self.arr, copy=True); _np.multiply.at(out, self.idx, v); return _wrap(out)


class _At:
  def __init__(self, arr):
    self.arr = arr

  def __getitem__(self, idx):
    return _AtIndex(self.arr, idx)


class Arr(_np.ndarray):
  """ndarray with jax's functional update syntax."""

  @property
  def at(self):
    return _At(self)

  def take(self, indices, axis=None, mode=None, **kw):
    return _wrap(_np.take(_np.asarray(self), indices, axis=axis, mode=mode or 'raise'))

  # jax arrays are immutable: augmented assignment rebinds (and may promote), never writes in place
  def __iadd__(self, o): return self + o
  def __isub__(self, o): return self - o
  def __imul__(self, o): return self * o
  def __itruediv__(self, o): return self / o


ndarray = Arr


def _fix(a):
  if a.dtype == _np.float64 and DEFAULT_FLOAT is _np.float32:
    a = a.astype(_np.float32)
  elif a.dtype == _np.int64:
    a = a.astype(_np.int32)
  return a


def _wrap(x):
  if isinstance(x, _np.ndarray):
    return _fix(x).view(Arr)
  if isinstance(x, _np.generic):
    return _fix(_np.asarray(x)).view(Arr)
  if isinstance(x, tuple):
    return tuple(_wrap(v) for v in x)
  if isinstance(x, list):
    return [_wrap(v) for v in x]
  return x


def array(x, dtype=None):
  a = _np.asarray(x, dtype=dtype)
  if dtype is None and a.dtype.kind == 'f':
    a = a.astype(DEFAULT_FLOAT)
  return _wrap(_np.array(a, copy=True))


asarray = array


def zeros(shape, dtype=None):
  return _wrap(_np.zeros(shape, dtype or DEFAULT_FLOAT))


def ones(shape, dtype=None):
  return _wrap(_np.ones(shape, dtype or DEFAULT_FLOAT))


def eye(n, dtype=None):
  return _wrap(_np.eye(n, dtype=dtype or DEFAULT_FLOAT))


def full(shape, v, dtype=None):
  return _wrap(_np.full(shape, v, dtype or (DEFAULT_FLOAT if isinstance(v, float) else None)))


def take(x, indices, axis=None, mode=None, **kw):
  return _wrap(_np.take(_np.asarray(x), indices, axis=axis, mode=mode or 'raise'))


WHERE_HOOK = None   # tools/gen_reference_golden_big.py: observes the reference's branch selections (it never alters them)


def where(c, a=None, b=None):
  if a is None:
    return _wrap(_np.where(c))
  if WHERE_HOOK is not None:
    import sys as _sys
    WHERE_HOOK(_sys._getframe(1).f_code.co_name, c)
  return _wrap(_np.where(c, a, b))


class _Linalg:
  def __getattr__(self, name):
    f = getattr(_np.linalg, name)
    return lambda *a, **k: _wrap(f(*a, **k))


linalg = _Linalg()


def __getattr__(name):   # everything else: the NumPy function of the same name, outputs re-wrapped
  f = getattr(_np, name)
  if callable(f) and not isinstance(f, type):
    def g(*a, **k):
      if isinstance(k.get('axis'), range):   # jax accepts any sequence of axes
        k['axis'] = tuple(k['axis'])
      return _wrap(f(*a, **k))
    g.__name__ = name
    return g
  return f
