"""lax.scan as a Python loop."""
from jax import numpy as jp
from jax.tree_util import tree_map, tree_leaves


def scan(f, init, xs=None, length=None, reverse=False):
  # reverse: the loop runs from the last element to the first; ys keep the order of xs (jax.lax.scan semantics)
  if xs is not None and not tree_leaves(xs):
    xs = None                      # scan over () with an explicit length
  n = length if xs is None else jp.shape(tree_leaves(xs)[0])[0]
  carry, ys = init, []
  for i in (range(n - 1, -1, -1) if reverse else range(n)):
    x = None if xs is None else tree_map(lambda a: a[i], xs)
    carry, y = f(carry, x)
    ys.append(y)
  if reverse:
    ys = ys[::-1]
  if ys and ys[0] is not None:
    ys = tree_map(lambda *v: jp.stack(v), *ys)
  else:
    ys = None
  return carry, ys


def cond(pred, t, f, *ops):
  return t(*ops) if pred else f(*ops)


def stop_gradient(x):
  return x


def psum(x, axis_name=None):
  raise NotImplementedError('no device axis in the stand-in')
