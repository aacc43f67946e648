"""Throughput vs batch size (env-steps/s) for the pipeline step: small RL-sized batches up to 1 M envs.
  python tools/batch_sweep.py [ant|humanoid]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from brax_b200 import native, workloads

model = sys.argv[1] if len(sys.argv) > 1 else 'ant'
dev = torch.device('cuda', 0)
rows = []
for n in (64, 256, 1024, 2048, 4096, 8192, 16384, 65536, 262144):
  sys_, q, qd = workloads.reset(model, 0, n, 0, dev)
  nm = native.model_for(sys_, 0)
  a, b = nm.init(q, qd), nm.alloc(n)
  act = workloads.action(model, 0, n, 0, 0, dev)
  for _ in range(3):
    nm.step(a, act, 5, out=b); a, b = b, a
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  reps = 20 if n <= 16384 else 5
  torch.cuda.synchronize(); e0.record()
  for _ in range(reps):
    nm.step(a, act, 5, out=b); a, b = b, a
  e1.record(); torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / reps
  rows.append({'envs': n, 'ms_per_env_step': ms, 'env_steps_per_s': n / (ms * 1e-3)})
print(json.dumps({'model': model, 'rows': rows}))
