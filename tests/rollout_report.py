"""1000-step free-running rollouts (BASELINE configs[0]/[1] shape): CUDA path vs the
float32 oracle vs the float64 oracle on identical inputs.  Trajectories of a
contact-rich articulated system decorrelate, so besides the
per-env agreement over time this records ENSEMBLE statistics, which must agree
between implementations however long the rollout is.
  python tests/rollout_report.py [ant|humanoid] [n_env] [n_steps] -> JSON on stdout"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # repo root
import torch  # noqa: E402
from brax_b200 import workloads  # noqa: E402
from brax_b200.generalized import pipeline  # noqa: E402
from oracle import oracle as O  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else 'ant'
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
dev = torch.device('cuda', 0)
sys_, q, qd = workloads.reset(model, 0, n, 0, dev)
nf = workloads.N_FRAMES[model]
o32, o64 = O.Oracle(sys_, np.float32), O.Oracle(sys_, np.float64)
r32 = o32.init(q.cpu().numpy(), qd.cpu().numpy())
r64 = o64.init(q.cpu().numpy().astype(np.float64), qd.cpu().numpy().astype(np.float64))
st = pipeline.init(sys_, q, qd)
marks = sorted(set([1, 2, 5, 10, 20, 50, 100, 200, 500, 1000]) & set(range(1, steps + 1)))
rows = []


def stats(qv, qdv, dist):
  return {'root_z_mean': float(qv[:, 2].mean()), 'root_z_std': float(qv[:, 2].std()),
          'root_xy_dist_mean': float(np.linalg.norm(qv[:, :2], axis=1).mean()),
          'joint_speed_rms': float(np.sqrt((qdv[:, 6:] ** 2).mean())), 'contacts_active_frac': float((dist < 0).mean())}


def within(a, b):
  e = np.abs(a - b) / (1e-5 + 1e-4 * np.abs(b))
  return float((e.max(1) <= 1).mean())


for k in range(1, steps + 1):
  act = workloads.action(model, 0, n, 0, k, dev)
  st = pipeline.step(sys_, st, act, debug=(k in marks), n_frames=nf)
  a = act.cpu().numpy()
  o32.step(r32, a, nf); o64.step(r64, a.astype(np.float64), nf)
  if k in marks:
    gq, gqd = st.q.cpu().numpy(), st.qd.cpu().numpy()
    gd = st.contact['con_dist'].cpu().numpy()
    # the reference algorithm diverges for a few envs on long random-action rollouts (no termination):
    # count them per implementation, compare the rest
    fin = np.isfinite(gq).all(1) & np.isfinite(r32['q']).all(1) & np.isfinite(r64['q']).all(1)
    fin &= (np.abs(gq[:, 2]) < 5) & (np.abs(r32['q'][:, 2]) < 5) & (np.abs(r64['q'][:, 2]) < 5)
    row_nonfinite = {'gpu': int((~np.isfinite(gq).all(1)).sum()), 'oracle_f32': int((~np.isfinite(r32['q']).all(1)).sum()),
                     'oracle_f64': int((~np.isfinite(r64['q']).all(1)).sum()), 'envs_compared': int(fin.sum())}
    gq, gqd, gd = gq[fin], gqd[fin], gd[fin]
    a32 = {k2: r32[k2][fin] for k2 in ('q', 'qd', 'con_dist')}
    a64 = {k2: r64[k2][fin] for k2 in ('q', 'qd', 'con_dist')}
    rows.append({'env_step': k, 'diverged_envs': row_nonfinite,
                 'frac_envs_q_within_tol_gpu_vs_f32': within(gq, a32['q']),
                 'frac_envs_q_within_tol_f64_vs_f32': within(a64['q'], a32['q']),
                 'median_max_abs_q_gap_gpu_vs_f32': float(np.median(np.abs(gq - a32['q']).max(1))),
                 'median_max_abs_q_gap_f64_vs_f32': float(np.median(np.abs(a64['q'] - a32['q']).max(1))),
                 'ensemble_gpu': stats(gq, gqd, gd),
                 'ensemble_oracle_f32': stats(a32['q'], a32['qd'], a32['con_dist']),
                 'ensemble_oracle_f64': stats(a64['q'], a64['qd'], a64['con_dist'])})
print(json.dumps({'model': model, 'envs': n, 'env_steps': steps, 'n_frames': nf, 'tolerance': 'rtol 1e-4 / atol 1e-5 on q',
                  'rows': rows}, indent=1))
