"""Hopper and Walker2d (reference `brax/envs/hopper.py`, `brax/envs/walker2d.py`, backend='generalized').

Planar models with plane-capsule contacts.  Env arithmetic: the planar kind
(forward velocity of link 0; healthy = strict ranges on z, the root angle and, for Hopper, the state
vector; obs = q with q[1] replaced by the root's world z, qd clipped to +-10)."""
import numpy as np
import torch

from brax_b200 import envs_assets, native, sharding
from brax_b200.envs.base import FusedEnv

METRICS = ('reward_forward', 'reward_ctrl', 'reward_healthy', 'x_position', 'x_velocity')
_SLOTS = {'reward_forward': 0, 'reward_healthy': 1, 'reward_ctrl': 2, 'x_position': 4, 'x_velocity': 7}
_INF = 3.0e38


class _Planar(FusedEnv):
  def __init__(self, asset, forward_reward_weight, ctrl_cost_weight, healthy_reward, terminate_when_unhealthy,
               healthy_state_range, healthy_z_range, healthy_angle_range, reset_noise_scale,
               exclude_current_positions_from_observation, backend, n_frames, **kwargs):
    if backend != 'generalized':
      raise ValueError('brax_b200 implements the generalized backend only')
    clampf = lambda v: float(max(-_INF, min(_INF, v)))   # noqa: E731  (keeps the spec's floats finite)
    spec = native.EnvSpecC()
    spec.kind = native.ENV_PLANAR
    spec.obs_skip = 1 if exclude_current_positions_from_observation else 0
    spec.terminate_when_unhealthy = int(bool(terminate_when_unhealthy))
    spec.forward_reward_weight = forward_reward_weight
    spec.ctrl_cost_weight = ctrl_cost_weight
    spec.healthy_reward = healthy_reward
    spec.healthy_z_min, spec.healthy_z_max = clampf(healthy_z_range[0]), clampf(healthy_z_range[1])
    spec.healthy_angle_min, spec.healthy_angle_max = clampf(healthy_angle_range[0]), clampf(healthy_angle_range[1])
    spec.healthy_state_min, spec.healthy_state_max = clampf(healthy_state_range[0]), clampf(healthy_state_range[1])
    self._reset_noise_scale = reset_noise_scale
    super().__init__(envs_assets.load(asset), spec, METRICS, n_frames, metric_slots=_SLOTS, **kwargs)

  def _reset_q_qd(self, env_begin, n, seed, device):
    # q = init_q + U(-s, s); qd = U(-s, s)   (hopper.py:194-202, walker2d.py:181-189)
    s = self._reset_noise_scale
    init_q = torch.as_tensor(np.asarray(self.sys.init_q, np.float32), device=device)
    q = init_q[None] + sharding.uniform(env_begin, n, self.sys.nq, seed, 1, -s, s, device)
    qd = sharding.uniform(env_begin, n, self.sys.nv, seed, 2, -s, s, device)
    return q.contiguous(), qd.contiguous()


class Hopper(_Planar):
  """Constructor arguments as reference envs/hopper.py:146-158."""

  def __init__(self, forward_reward_weight=1.0, ctrl_cost_weight=1e-3, healthy_reward=1.0, terminate_when_unhealthy=True,
               healthy_state_range=(-100.0, 100.0), healthy_z_range=(0.7, float('inf')), healthy_angle_range=(-0.2, 0.2),
               reset_noise_scale=5e-3, exclude_current_positions_from_observation=True, backend='generalized',
               n_frames=4, **kwargs):
    super().__init__('hopper', forward_reward_weight, ctrl_cost_weight, healthy_reward, terminate_when_unhealthy,
                     healthy_state_range, healthy_z_range, healthy_angle_range, reset_noise_scale,
                     exclude_current_positions_from_observation, backend, n_frames, **kwargs)


class Walker2d(_Planar):
  """Constructor arguments as reference envs/walker2d.py:129-140 (no state range)."""

  def __init__(self, forward_reward_weight=1.0, ctrl_cost_weight=1e-3, healthy_reward=1.0, terminate_when_unhealthy=True,
               healthy_z_range=(0.8, 2.0), healthy_angle_range=(-1.0, 1.0), reset_noise_scale=5e-3,
               exclude_current_positions_from_observation=True, backend='generalized', n_frames=4, **kwargs):
    super().__init__('walker2d', forward_reward_weight, ctrl_cost_weight, healthy_reward, terminate_when_unhealthy,
                     (-float('inf'), float('inf')), healthy_z_range, healthy_angle_range, reset_noise_scale,
                     exclude_current_positions_from_observation, backend, n_frames, **kwargs)
