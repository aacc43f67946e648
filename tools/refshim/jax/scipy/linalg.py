import scipy.linalg as _sl
from jax import numpy as jp


def solve(a, b, assume_a='gen', **kw):
  return jp._wrap(_sl.solve(a, b, assume_a=assume_a))
