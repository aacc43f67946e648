"""Humanoid (reference `brax/envs/humanoid.py`, backend='generalized')."""
import numpy as np
import torch

from brax_b200 import envs_assets, native, sharding
from brax_b200.envs.base import FusedEnv

METRICS = ('forward_reward', 'reward_linvel', 'reward_quadctrl', 'reward_alive', 'x_position', 'y_position',
           'distance_from_origin', 'x_velocity', 'y_velocity')


class Humanoid(FusedEnv):
  """Constructor arguments as reference envs/humanoid.py:180-191."""

  def __init__(self, forward_reward_weight=1.25, ctrl_cost_weight=0.1, healthy_reward=5.0,
               terminate_when_unhealthy=True, healthy_z_range=(1.0, 2.0), reset_noise_scale=1e-2,
               exclude_current_positions_from_observation=True, backend='generalized', n_frames=5, **kwargs):
    if backend != 'generalized':
      raise ValueError('brax_b200 implements the generalized backend only')
    spec = native.EnvSpecC()
    spec.kind = native.ENV_COM_VELOCITY
    spec.obs_skip = 2 if exclude_current_positions_from_observation else 0
    spec.terminate_when_unhealthy = int(bool(terminate_when_unhealthy))
    spec.forward_reward_weight = forward_reward_weight
    spec.ctrl_cost_weight = ctrl_cost_weight
    spec.healthy_reward = healthy_reward
    spec.healthy_z_min, spec.healthy_z_max = healthy_z_range
    self._reset_noise_scale = reset_noise_scale
    super().__init__(envs_assets.load('humanoid'), spec, METRICS, n_frames, **kwargs)

  def _reset_q_qd(self, env_begin, n, seed, device):
    # q = init_q + U(-s, s); qd = U(-s, s)   (humanoid.py:231-237)
    s = self._reset_noise_scale
    init_q = torch.as_tensor(np.asarray(self.sys.init_q, np.float32), device=device)
    q = init_q[None] + sharding.uniform(env_begin, n, self.sys.nq, seed, 1, -s, s, device)
    qd = sharding.uniform(env_begin, n, self.sys.nv, seed, 2, -s, s, device)
    return q.contiguous(), qd.contiguous()
