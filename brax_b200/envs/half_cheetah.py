"""HalfCheetah (reference `brax/envs/half_cheetah.py`, backend='generalized').

Physics: plane-capsule contacts, generic kernel variant
(49 constraint rows).  The env arithmetic is the root-velocity kind Ant uses, with no
health term and no termination: reward = w * x_velocity - ctrl_cost (half_cheetah.py:178-199)."""
import numpy as np
import torch

from brax_b200 import envs_assets, native, sharding
from brax_b200.envs.base import FusedEnv

METRICS = ('x_position', 'x_velocity', 'reward_ctrl', 'reward_run')
# slots of the root-velocity metrics row written by the kernel (bxg_core.cuh env_epilogue)
_SLOTS = {'reward_run': 0, 'reward_ctrl': 2, 'x_position': 4, 'x_velocity': 7}


class Halfcheetah(FusedEnv):
  """Constructor arguments as reference envs/half_cheetah.py:124-133."""

  def __init__(self, forward_reward_weight=1.0, ctrl_cost_weight=0.1, reset_noise_scale=0.1,
               exclude_current_positions_from_observation=True, backend='generalized', n_frames=5, **kwargs):
    if backend != 'generalized':
      raise ValueError('brax_b200 implements the generalized backend only')
    spec = native.EnvSpecC()
    spec.kind = native.ENV_ROOT_VELOCITY
    spec.obs_skip = 1 if exclude_current_positions_from_observation else 0
    spec.terminate_when_unhealthy = 0
    spec.forward_reward_weight = forward_reward_weight
    spec.ctrl_cost_weight = ctrl_cost_weight
    spec.healthy_reward = 0.0
    spec.healthy_z_min, spec.healthy_z_max = -3.0e38, 3.0e38
    self._reset_noise_scale = reset_noise_scale
    super().__init__(envs_assets.load('halfcheetah'), spec, METRICS, n_frames, metric_slots=_SLOTS, **kwargs)

  def _reset_q_qd(self, env_begin, n, seed, device):
    # q = init_q + U(-s, s); qd = s * N(0, 1)   (half_cheetah.py:157-163)
    s = self._reset_noise_scale
    init_q = torch.as_tensor(np.asarray(self.sys.init_q, np.float32), device=device)
    q = init_q[None] + sharding.uniform(env_begin, n, self.sys.nq, seed, 1, -s, s, device)
    qd = s * sharding.normal(env_begin, n, self.sys.nv, seed, 2, device)
    return q.contiguous(), qd.contiguous()
