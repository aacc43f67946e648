"""Throughput of the pipeline step for any compiled asset (env-steps/s, n_frames as the env uses):
  python tools/model_bench.py halfcheetah 65536 [n_frames]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from brax_b200 import envs_assets, native

name = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
nf = int(sys.argv[3]) if len(sys.argv) > 3 else {'hopper': 4, 'walker2d': 4, 'swimmer': 4, 'reacher': 2, 'inverted_pendulum': 2, 'inverted_double_pendulum': 2}.get(name, 5)
dev = torch.device('cuda', 0)
s = envs_assets.load(name)
g = torch.Generator(device='cpu').manual_seed(0)
q = torch.as_tensor(np.asarray(s.init_q, np.float32))[None] + (torch.rand((n, s.nq), generator=g) * 0.02 - 0.01)
qd = torch.rand((n, s.nv), generator=g) * 0.02 - 0.01
nm = native.model_for(s, 0)
a, b = nm.init(q.to(dev).contiguous(), qd.to(dev).contiguous()), nm.alloc(n)
act = (torch.rand((n, s.nu), generator=g) * 2 - 1).to(dev).contiguous()
for _ in range(3):
  nm.step(a, act, nf, out=b); a, b = b, a
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
torch.cuda.synchronize(); e0.record()
for _ in range(reps):
  nm.step(a, act, nf, out=b); a, b = b, a
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(json.dumps({'model': name, 'envs': n, 'n_frames': nf, 'plan': native.plan(s), 'ms_per_env_step': ms, 'env_steps_per_s': n / (ms * 1e-3),
                  'finite': bool(torch.isfinite(a['q']).all())}))
