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

N_FRAMES = {'ant': 5, 'humanoid': 5}
# algorithmic bytes per env-step: read q,qd,act; write q,qd,x,xd (fp32), SURVEY 8d
ALGO_BYTES = {'ant': 732, 'humanoid': 1016}


def reset(model: str, env_begin: int, n_env: int, seed: int, device):
  sys = envs_assets.load(model)
  init_q = torch.as_tensor(np.asarray(sys.init_q, np.float32), device=device)
  if model == 'ant':
    q = init_q[None] + sharding.uniform(env_begin, n_env, sys.nq, seed, 1, -0.1, 0.1, device)
    qd = 0.1 * sharding.normal(env_begin, n_env, sys.nv, seed, 2, device)
  elif model == 'humanoid':
    q = init_q[None] + sharding.uniform(env_begin, n_env, sys.nq, seed, 1, -0.01, 0.01, device)
    qd = sharding.uniform(env_begin, n_env, sys.nv, seed, 2, -0.01, 0.01, device)
  else:
    raise ValueError(model)
  return sys, q.contiguous(), qd.contiguous()


def action(model: str, env_begin: int, n_env: int, seed: int, step: int, device):
  sys = envs_assets.load(model)
  a = sharding.uniform(env_begin, n_env, sys.nu, seed, 1000 + step, -1.0, 1.0, device)
  if model == 'humanoid':
    lo = torch.as_tensor(np.asarray(sys.actuator.ctrl_range[:, 0], np.float32), device=device)
    hi = torch.as_tensor(np.asarray(sys.actuator.ctrl_range[:, 1], np.float32), device=device)
    a = (a + 1) * (hi - lo) * 0.5 + lo
  return a.contiguous()
