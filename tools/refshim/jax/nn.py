"""jax.nn entry points used by brax.training.distribution."""
import numpy as _np

from jax import numpy as jp


def softplus(x):
  return jp.array(_np.logaddexp(_np.asarray(x), 0.0))


def swish(x):
  x = _np.asarray(x)
  return jp.array(x / (1.0 + _np.exp(-x)))


silu = swish
