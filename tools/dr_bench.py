"""Throughput of the per-env-model kernels (domain randomisation) next to the shared-model kernels.

  python tools/dr_bench.py [n_envs]
"""
import json
import sys
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brax_b200 import base, native, workloads   # noqa: E402


def main():
  out = {}
  for name, n in (('ant', int(sys.argv[1]) if len(sys.argv) > 1 else 32768), ('humanoid', 8192)):
    s, q, qd = workloads.reset(name, 0, n, 0, 'cuda')
    rng = np.random.default_rng(0)
    scale = rng.uniform(0.8, 1.25, (n, s.num_links())).astype(np.float32)
    leaves = {'link.inertia.mass': np.asarray(s.link.inertia.mass, np.float32)[None] * scale,
              'link.inertia.i': np.asarray(s.link.inertia.i, np.float32)[None] * scale[:, :, None, None]}
    in_axes = base.tree_map(lambda x: None, s).tree_replace({k: 0 for k in leaves})
    import time
    t0 = time.time()
    systems = base.unbatch(s.tree_replace(leaves), in_axes)
    nm_b = native.BatchedNativeModel(systems, 0)
    t_create = time.time() - t0
    nm = native.NativeModel(s, 0)
    nf = workloads.N_FRAMES[name]
    acts = [workloads.action(name, 0, n, 0, k, 'cuda') for k in range(4)]
    res = {}
    for tag, m in (('shared', nm), ('per_env', nm_b)):
      a = m.init(q, qd); b = m.alloc(n)
      for k in range(3):
        m.step(a, acts[k % 4], nf, out=b); a, b = b, a
      torch.cuda.synchronize()
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record()
      for k in range(10):
        m.step(a, acts[k % 4], nf, out=b); a, b = b, a
      e1.record(); torch.cuda.synchronize()
      res[tag] = n * 10 / (e0.elapsed_time(e1) * 1e-3)
    out[name] = {'n_envs': n, 'shared_model_env_steps_per_s': res['shared'], 'per_env_models_env_steps_per_s': res['per_env'],
                 'model_create_s': t_create, 'kernel_ids': [nm.kernel_id, nm_b.kernel_id]}
    print(name, json.dumps(out[name]))
  os.makedirs('gpurun_out', exist_ok=True)
  with open('gpurun_out/r02_dr_bench.json', 'w') as f:
    json.dump(out, f, indent=1)


if __name__ == '__main__':
  main()
