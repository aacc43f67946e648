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


def _prefix_apply(a, ax, leaf_fn):
  """Walks `a` along the in_axes prefix tree `ax` (None: unmapped subtree, int: every leaf below mapped on that axis)."""
  import dataclasses
  from jax import tree_util as _tu
  if ax is None:
    return a
  if isinstance(ax, int):
    return _tree_map(lambda x: leaf_fn(x, ax), a)
  if isinstance(a, (tuple, list)):
    out = [_prefix_apply(t, ax[i], leaf_fn) for i, t in enumerate(a)]
    return type(a)(*out) if hasattr(a, '_fields') else type(a)(out)
  if isinstance(a, dict):
    return {k: _prefix_apply(v, ax[k], leaf_fn) for k, v in a.items()}
  if type(a) in _tu._registered:
    return dataclasses.replace(a, **{fl.name: _prefix_apply(getattr(a, fl.name), getattr(ax, fl.name), leaf_fn)
                                     for fl in _tu._node_fields(a)})
  raise TypeError(f'in_axes prefix {ax!r} does not match argument {type(a)}')


def vmap(fun, in_axes=0, out_axes=0):
  """vmap as a Python loop over the mapped axis, outputs stacked along out_axes.  in_axes entries may be ints, None or
  pytree prefixes of the argument (a System with 0 at its domain-randomised leaves: envs/wrappers/training.py:250-260)."""
  @functools.wraps(fun)
  def mapped(*args):
    axes = _axes_for(in_axes, len(args))
    sizes = set()
    for a, ax in zip(args, axes):
      _prefix_apply(a, ax, lambda x, k: sizes.add(_np.shape(x)[k]) or x)
    assert len(sizes) == 1, f'vmap: inconsistent or missing mapped sizes {sizes}'
    size = sizes.pop()
    assert size > 0, 'vmap over an empty batch'
    outs = []
    for i in range(size):
      sl = [_prefix_apply(a, ax, lambda x, k: numpy.take(x, i, axis=k)) for a, ax in zip(args, axes)]
      outs.append(fun(*sl))
    return _tree_map(lambda *xs: numpy.stack(xs, axis=out_axes), *outs)
  return mapped
