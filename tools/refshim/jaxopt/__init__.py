"""Stand-in for jaxopt.ProjectedGradient (jaxopt is not installable here).

NOT the reference: a restatement of jaxopt/_src/proximal_gradient.py with its defaults
(stepsize=0 -> backtracking line search, tol=1e-3, acceleration=True (FISTA),
decrease_factor=0.5), the same statement oracle/bxg_oracle.c makes.  The objective VALUE is
evaluated by calling the reference's own closure; its gradient is formed from the `a`, `b`
the closure captures (objective = sum(0.5 (a x + b)^2): grad = a^T (a x + b))."""
import collections

import numpy as np

from jax import numpy as jp
from jaxopt import projection  # noqa: F401

OptStep = collections.namedtuple('OptStep', 'params state')


class ProjectedGradient:
  def __init__(self, fun, projection, maxiter=500, maxls=15, tol=1e-3, stepsize=0.0, decrease_factor=0.5,
               acceleration=True, implicit_diff=True, **kw):
    assert stepsize == 0.0 and acceleration and decrease_factor == 0.5
    self.fun, self.projection, self.maxiter, self.maxls, self.tol = fun, projection, maxiter, maxls, tol
    free = dict(zip(fun.__code__.co_freevars, (c.cell_contents for c in fun.__closure__)))
    self.a, self.b = free['a'], free['b']
    self.stats = [0, 0]

  def _grad(self, x):
    return self.a.T @ (self.a @ x + self.b)

  def run(self, init_params, *args, **kwargs):
    dt = np.asarray(init_params).dtype.type
    eps = np.finfo(dt).eps
    x = jp.array(init_params); y = jp.array(init_params)
    t, stepsize, error, it = dt(1), dt(1), np.inf, 0
    while it < self.maxiter and (it == 0 or error > self.tol):
      fy, g = self.fun(y), self._grad(y)
      s = stepsize
      xn = self.projection(y - s * g)
      ls = 0
      while True:
        diff = xn - y
        sqdist, vd = jp.sum(diff ** 2), jp.sum(diff * g)
        fn = self.fun(xn)
        self.stats[1] += 1
        if not (s * (fn - fy) > s * vd + dt(0.5) * sqdist + eps) or ls >= self.maxls:
          break
        s = s * dt(0.5)
        xn = self.projection(y - s * g)
        ls += 1
      stepsize = dt(1) if s <= 1e-6 else s / dt(0.5)
      tn = dt(0.5) * (dt(1) + np.sqrt(dt(1) + dt(4) * t * t))
      y = xn + ((t - dt(1)) / tn) * (xn - x)
      gn = self._grad(xn)
      error = float(np.sqrt(np.sum((self.projection(xn - gn) - xn) ** 2)))
      x, t = xn, tn
      it += 1
      self.stats[0] += 1
    return OptStep(x, None)
