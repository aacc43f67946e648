"""Ant (reference `brax/envs/ant.py`, backend='generalized')."""
import numpy as np
import torch

from brax_b200 import envs_assets, native, sharding
from brax_b200.envs.base import FusedEnv

METRICS = ('reward_forward', 'reward_survive', 'reward_ctrl', 'reward_contact', 'x_position', 'y_position',
           'distance_from_origin', 'x_velocity', 'y_velocity', 'forward_reward')


class Ant(FusedEnv):
  """Constructor arguments as reference envs/ant.py:147-160."""

  def __init__(self, ctrl_cost_weight=0.5, use_contact_forces=False, contact_cost_weight=5e-4,
               healthy_reward=1.0, terminate_when_unhealthy=True, healthy_z_range=(0.2, 1.0),
               contact_force_range=(-1.0, 1.0), reset_noise_scale=0.1,
               exclude_current_positions_from_observation=True, backend='generalized', n_frames=5, **kwargs):
    if backend != 'generalized':
      raise ValueError('brax_b200 implements the generalized backend only')
    if use_contact_forces:
      raise NotImplementedError('use_contact_forces not implemented.')   # as the reference, ant.py:202-203
    spec = native.EnvSpecC()
    spec.kind = native.ENV_ROOT_VELOCITY
    spec.obs_skip = 2 if exclude_current_positions_from_observation else 0
    spec.terminate_when_unhealthy = int(bool(terminate_when_unhealthy))
    spec.forward_reward_weight = 1.0
    spec.ctrl_cost_weight = ctrl_cost_weight
    spec.healthy_reward = healthy_reward
    spec.healthy_z_min, spec.healthy_z_max = healthy_z_range
    self._reset_noise_scale = reset_noise_scale
    super().__init__(envs_assets.load('ant'), spec, METRICS, n_frames, **kwargs)

  def _reset_q_qd(self, env_begin, n, seed, device):
    # q = init_q + U(-s, s); qd = s * N(0, 1)   (ant.py:209-213)
    s = self._reset_noise_scale
    init_q = torch.as_tensor(np.asarray(self.sys.init_q, np.float32), device=device)
    q = init_q[None] + sharding.uniform(env_begin, n, self.sys.nq, seed, 1, -s, s, device)
    qd = s * sharding.normal(env_begin, n, self.sys.nv, seed, 2, device)
    return q.contiguous(), qd.contiguous()
