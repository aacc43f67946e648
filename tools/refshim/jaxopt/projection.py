from jax import numpy as jp


def projection_non_negative(x, hyperparams=None):
  return jp.maximum(x, 0)
