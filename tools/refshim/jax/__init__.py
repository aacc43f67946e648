"""Minimal NumPy-backed stand-in for the JAX entry points the reference path uses (see ../README.md)."""
import functools

import numpy as _np

from jax import numpy  # noqa: F401  (jax.numpy)
from jax import lax, nn, ops, random, scipy, tree, tree_util, typing  # noqa: F401
from jax.tree_util import tree_map as _tree_map, tree_leaves as _tree_leaves

Array = _np.ndarray


class _Config:   # the stand-in computes in float64 (tools/refshim/README.md)
  jax_enable_x64 = True


config = _Config()


class custom_jvp:   # forward evaluation only
  def __init__(self, fun):
    self.fun = fun
    functools.update_wrapper(self, fun)

  def __call__(self, *a, **k):
    return self.fun(*a, **k)

  def defjvp(self, f):
    return f


def jit(f=None, **kw):
  return f if f is not None else (lambda g: g)


def _axes_for(in_axes, n):
  if isinstance(in_axes, (tuple, list)):
    assert len(in_axes) == n, (in_axes, n)
    return list(in_axes)
  return [in_axes] * n


def vmap(fun, in_axes=0, out_axes=0):
  """vmap as a Python loop over the mapped axis, outputs stacked along out_axes."""
  @functools.wraps(fun)
  def mapped(*args):
    axes = _axes_for(in_axes, len(args))
    size = None
    for a, ax in zip(args, axes):
      if ax is None:
        continue
      assert isinstance(ax, int), 'pytree in_axes are not supported by this stand-in'
      for leaf in _tree_leaves(a):
        s = _np.shape(leaf)[ax]
        assert size is None or size == s, 'vmap: inconsistent sizes'
        size = s
    assert size is not None and size > 0, 'vmap over an empty or unmapped batch'
    outs = []
    for i in range(size):
      sl = [a if ax is None else _tree_map(lambda x, ax=ax: numpy.take(x, i, axis=ax), a) for a, ax in zip(args, axes)]
      outs.append(fun(*sl))
    return _tree_map(lambda *xs: numpy.stack(xs, axis=out_axes), *outs)
  return mapped
