"""Deterministic stand-in for jax.random: keys are uint32[2]; streams come from NumPy's
PCG64 (NOT threefry: the golden files record the drawn q / qd, nothing depends on the bits)."""
import numpy as _np

from jax import numpy as jp


def PRNGKey(seed):
  return _np.array([0, seed], _np.uint32).view(jp.Arr)


def _gen(key):
  k = _np.asarray(key, _np.uint64).reshape(-1)
  return _np.random.default_rng(int(k[0]) * (1 << 32) + int(k[1]))


def split(key, num=2):
  g = _gen(key)
  return g.integers(0, 1 << 32, size=(num, 2), dtype=_np.uint32).view(jp.Arr)


def uniform(key, shape=(), dtype=None, minval=0.0, maxval=1.0):
  return jp.array(_gen(key).uniform(minval, maxval, size=shape))


def normal(key, shape=(), dtype=None):
  return jp.array(_gen(key).standard_normal(size=shape))
