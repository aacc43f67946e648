"""Share of the step's time per phase, measured in place with clock64() by a tuning build of the library:
  python tools/build_alt.py timers -DBXG_PHASE_TIMERS
  BXG_LIB=brax_b200/libbxg_timers.so python tools/phase_timers.py [ant|humanoid|humanoid_falls] [n_env]
Per-warp cycles between phase boundaries, summed over warps and env passes (include/bxg.h BxgDiag.phase_cycles)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from brax_b200 import native, workloads  # noqa: E402

NAMES = ['load', 'dynamics (tau, RNE)', 'constraint.force', 'integrate', 'kinematics', 'transform_com', 'mass.matrix',
         'matrix_inv (Newton-Schulz)', 'constraint.jacobian', 'env prologue / epilogue', 'store', 'lean entry', 'CTA barrier after constraint.force', 'constraint.force: active set, A, b', 'constraint.force: FISTA iterations (without the line searches)', 'constraint.force: line-search trials']


def main():
  model = sys.argv[1] if len(sys.argv) > 1 else 'ant'
  n = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 30 * 8
  dev = torch.device('cuda', 0)
  s, q, qd = workloads.reset(model, 0, n, 0, dev)
  nm = native.NativeModel(s, 0)
  st = nm.init(q, qd)
  nf = workloads.N_FRAMES[model]
  for k in range(3):
    st = nm.step(st, workloads.action(model, 0, n, 0, k, dev), nf)
  diag = nm.alloc_diag(n)
  for k in range(3, 6):
    st = nm.step(st, workloads.action(model, 0, n, 0, k, dev), nf, diag=diag)
  torch.cuda.synchronize()
  cyc = diag['phase_cycles'].cpu().numpy().astype(float)
  tot = cyc.sum()
  ls_ = nm.launch_shape(n)
  warp_substeps = ls_['grid'] * (ls_['threads_per_cta'] // 32) * -(-n // (ls_['grid'] * ls_['envs_per_cta'])) * 3 * nf
  stats = diag['stats'].cpu().numpy().astype(float)      # counters accumulated over the three timed launches
  out = {'model': model, 'n_env': n, 'launch': ls_, 'share_of_warp_cycles': {NAMES[i]: round(cyc[i] / tot, 4) for i in range(len(NAMES)) if cyc[i] > 0},
         'cycles_per_warp_substep': {NAMES[i]: round(cyc[i] / warp_substeps) for i in range(len(NAMES)) if cyc[i] > 0},
         'cycles_per_warp_substep_total': round(tot / warp_substeps),
         'per_env_substep': {'pg_iterations': round(stats[:, 0].mean() / (3 * nf), 2), 'line_search_trials': round(stats[:, 1].mean() / (3 * nf), 2),
                             'newton_schulz_accepts': round(stats[:, 2].mean() / (3 * nf), 2)}}
  print(json.dumps(out, indent=1))


if __name__ == '__main__':
  main()
