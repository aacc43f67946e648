"""Reference-sized golden rollouts from the REFERENCE'S OWN SOURCE, with its branch decisions.

Same mechanism as tools/gen_reference_golden.py (brax.generalized.pipeline imported unmodified
from /root/reference, run on NumPy float64 through tools/refshim/), at 32 envs x 50 env-steps
(250 physics substeps per env) for Ant, Humanoid and Humanoid-falls (BASELINE configs 1, 2, 4
reset distributions, brax_b200/workloads.py).  Per physics substep of the reference run it records

  stats[t, f, e] = (projected-gradient iterations, line-search trials,
                    Newton-Schulz accepted candidates, Newton-Schulz cold start)
  con_active[t, f, e, c] = contact c active (constraint.py:174: dist < 0)

observed from the reference's own `jp.where` selections in `math.inv_approximate`
(math.py:297,302) and from the calls jaxopt's stand-in makes to the reference's objective closure
(constraint.py:222-228), so that branch agreement of the oracle, the emulator and the CUDA path is
asserted against the REFERENCE RUN, not against our own oracle.  Stored:

  q0, qd0, act[T, E, nu]                 inputs (float32 values)
  q[T, E, nq], qd[T, E, nv]              state after every env-step (float64)
  ck_steps[K], ck_minv[K, 2, E, nv, nv]  mass_mx_inv before and after env-steps ck_steps (float64): with q, qd they
                                         determine the whole State, so every checkpoint is an exact one-env-step map
  stats[T, F, E, 4] (int16), con_active[T, F, E, ncon] (bool)

  python tools/gen_reference_golden_big.py [ant,humanoid,humanoid_falls]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import gen_reference_golden as G   # noqa: E402  (sets up the stand-ins and imports the reference)

import jax.numpy as jp             # noqa: E402  (the stand-in)
import jaxopt                      # noqa: E402  (the stand-in)
from brax_b200 import workloads    # noqa: E402

E, T = 32, 50
CK = (0, 10, 20, 30, 40)


class Recorder:
  """Observes the reference run: never changes a value."""

  def __init__(self):
    self.reset()
    jp.WHERE_HOOK = self.on_where
    orig_run = jaxopt.ProjectedGradient.run

    def run(pg, *a, **k):
      before = list(pg.stats)
      out = orig_run(pg, *a, **k)
      self.pg_iters += pg.stats[0] - before[0]; self.pg_trials += pg.stats[1] - before[1]
      return out
    jaxopt.ProjectedGradient.run = run

  def reset(self):
    self.pg_iters = self.pg_trials = self.ns_accepts = self.ns_cold = 0

  def on_where(self, fn, cond):
    if fn == 'body_fn':            # math.py:297  a_inv_next = where(err_next < err, a_inv_next, a_inv)
      self.ns_accepts += int(bool(np.asarray(cond)))
    elif fn == 'inv_approximate':  # math.py:302  cold start where(safe_norm(r0) > 1, ...)
      self.ns_cold += int(bool(np.asarray(cond)))

  def take(self):
    v = (self.pg_iters, self.pg_trials, self.ns_accepts, self.ns_cold)
    self.reset()
    return v


def main():
  only = set(sys.argv[1].split(',')) if len(sys.argv) > 1 else None
  rec = Recorder()
  for name in ('ant', 'humanoid', 'humanoid_falls'):
    if only is not None and name not in only:
      continue
    s, q0, qd0 = workloads.reset(name, 0, E, seed=11, device='cpu')
    nf = workloads.N_FRAMES[name]
    q0 = q0.numpy().astype(np.float64); qd0 = qd0.numpy().astype(np.float64)
    act = np.stack([workloads.action(name, 0, E, seed=11, step=t, device='cpu').numpy() for t in range(T)]).astype(np.float64)
    G.mjx.PAIRS = s.contact_pairs() if s.geom_bodyid is not None and len(s.contact_pairs().geom1) else None
    rs = G.reference_system(s)
    ncon = len(s.contact_pairs().geom1)
    q = np.zeros((T, E, s.nq)); qd = np.zeros((T, E, s.nv))
    stats = np.zeros((T, nf, E, 4), np.int16); con_active = np.zeros((T, nf, E, ncon), bool)
    ck_minv = np.zeros((len(CK), 2, E, s.nv, s.nv))
    t0 = time.time()
    for e in range(E):
      rec.reset()
      st = G.ref_pipeline.init(rs, jp.array(q0[e]), jp.array(qd0[e]))
      rec.reset()   # (init's exact inverse is not a Newton-Schulz run)
      for t in range(T):
        if t in CK:
          ck_minv[CK.index(t), 0, e] = np.asarray(st.mass_mx_inv)
        for f in range(nf):
          st = G.ref_pipeline.step(rs, st, jp.array(act[t, e]))
          stats[t, f, e] = rec.take()
          con_active[t, f, e] = np.asarray(st.con_diag)[0:4 * ncon:4] != 0
        q[t, e] = np.asarray(st.q); qd[t, e] = np.asarray(st.qd)
        if t in CK:
          ck_minv[CK.index(t), 1, e] = np.asarray(st.mass_mx_inv)
      print(f'{name}: env {e + 1}/{E} done, {time.time() - t0:.0f} s', flush=True)
    path = os.path.join(ROOT, 'tests', 'golden', f'ref_big_{name}.npz')
    np.savez_compressed(path, q0=q0, qd0=qd0, act=act, q=q, qd=qd, stats=stats, con_active=con_active,
                        ck_steps=np.array(CK), ck_minv=ck_minv)
    sm = stats.reshape(-1, 4).astype(np.int64)
    print(f'{name}: {E} envs x {T} env-steps; per substep mean pg iters {sm[:, 0].mean():.2f}, trials {sm[:, 1].mean():.2f}, '
          f'N-S accepts {sm[:, 2].mean():.2f}, cold starts {sm[:, 3].mean():.3f}; active contacts {con_active.mean():.3f}; '
          f'finite {np.isfinite(q).all()}; wrote {path} ({os.path.getsize(path) // 1024} KB)')


if __name__ == '__main__':
  main()
