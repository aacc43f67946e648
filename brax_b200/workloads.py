"""Synthetic workloads named by BASELINE.json `configs`.

reset distributions follow the reference envs:
  ant       q = init_q + U(-0.1, 0.1), qd = 0.1 N(0,1)     (envs/ant.py:205-215)
  humanoid  q = init_q + U(-0.01,0.01), qd = U(-0.01,0.01) (envs/humanoid.py:227-239)
actions U(-1,1); Humanoid rescales them to the ctrl range (envs/humanoid.py:260-262).
"""
from __future__ import annotations

import numpy as np
import torch

from brax_b200 import envs_assets, sharding

N_FRAMES = {'ant': 5, 'humanoid': 5, 'humanoid_falls': 5}
# algorithmic bytes per env-step: read q,qd,act; write q,qd,x,xd (fp32), SURVEY 8d
ALGO_BYTES = {'ant': 732, 'humanoid': 1016, 'humanoid_falls': 1016}


def reset(model: str, env_begin: int, n_env: int, seed: int, device):
  sys = envs_assets.load('humanoid' if model == 'humanoid_falls' else model)
  init_q = torch.as_tensor(np.asarray(sys.init_q, np.float32), device=device)
  if model == 'ant':
    q = init_q[None] + sharding.uniform(env_begin, n_env, sys.nq, seed, 1, -0.1, 0.1, device)
    qd = 0.1 * sharding.normal(env_begin, n_env, sys.nv, seed, 2, device)
  elif model == 'humanoid':
    q = init_q[None] + sharding.uniform(env_begin, n_env, sys.nq, seed, 1, -0.01, 0.01, device)
    qd = sharding.uniform(env_begin, n_env, sys.nv, seed, 2, -0.01, 0.01, device)
  elif model == 'humanoid_falls':
    # BASELINE configs[3] "contact-heavy (randomised falls)", SURVEY 8d config 4: root height
    # U(0.5, 1.4), uniformly random root orientation, joints uniform within their ranges, qd ~ N(0,1)
    sys = envs_assets.load('humanoid')
    init_q = torch.as_tensor(np.asarray(sys.init_q, np.float32), device=device)
    q = init_q[None].repeat(n_env, 1)
    q[:, 2] = sharding.uniform(env_begin, n_env, 1, seed, 1, 0.5, 1.4, device)[:, 0]
    quat = sharding.normal(env_begin, n_env, 4, seed, 3, device)
    q[:, 3:7] = quat / quat.norm(dim=1, keepdim=True)
    lo = torch.as_tensor(np.asarray(sys.dof.limit[0][6:], np.float32), device=device)
    hi = torch.as_tensor(np.asarray(sys.dof.limit[1][6:], np.float32), device=device)
    u = sharding.uniform(env_begin, n_env, sys.nq - 7, seed, 4, 0.0, 1.0, device)
    q[:, 7:] = lo[None] + (hi - lo)[None] * u
    qd = sharding.normal(env_begin, n_env, sys.nv, seed, 2, device)
  else:
    raise ValueError(model)
  return sys, q.contiguous(), qd.contiguous()


def action(model: str, env_begin: int, n_env: int, seed: int, step: int, device):
  sys = envs_assets.load('humanoid' if model == 'humanoid_falls' else model)
  a = sharding.uniform(env_begin, n_env, sys.nu, seed, 1000 + step, -1.0, 1.0, device)
  if model.startswith('humanoid'):
    lo = torch.as_tensor(np.asarray(sys.actuator.ctrl_range[:, 0], np.float32), device=device)
    hi = torch.as_tensor(np.asarray(sys.actuator.ctrl_range[:, 1], np.float32), device=device)
    a = (a + 1) * (hi - lo) * 0.5 + lo
  return a.contiguous()
