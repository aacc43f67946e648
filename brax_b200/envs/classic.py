"""The small classic-control envs (reference `brax/envs/{inverted_pendulum, inverted_double_pendulum,
reacher, swimmer}.py`, backend='generalized'), HumanoidStandup and Pusher.

No contacts; slide joints (the carts), a 2-dof link (Reacher's target) and, for Swimmer, the fluid
forces of `brax/fluid.py` (compiled into the generic kernel variant and the small 4-lane one Swimmer runs on).  Each env has its own kind in
the kernel's env epilogue (include/bxg.h BXG_ENV_CARTPOLE ... BXG_ENV_SWIMMER)."""
import math

import numpy as np
import torch

from brax_b200 import envs_assets, native, sharding
from brax_b200.envs.base import FusedEnv, State

_INF = 3.0e38


def _spec(kind):
  spec = native.EnvSpecC()
  spec.kind = kind
  spec.obs_skip = 0
  spec.terminate_when_unhealthy = 0
  spec.forward_reward_weight = 1.0
  spec.ctrl_cost_weight = 0.0
  spec.healthy_reward = 0.0
  spec.healthy_z_min, spec.healthy_z_max = -_INF, _INF
  spec.healthy_angle_min, spec.healthy_angle_max = -_INF, _INF
  spec.healthy_state_min, spec.healthy_state_max = -_INF, _INF
  return spec


def _check_backend(backend):
  if backend != 'generalized':
    raise ValueError('brax_b200 implements the generalized backend only')


class InvertedPendulum(FusedEnv):
  """Reference envs/inverted_pendulum.py:99-154: reward 1 per step, done when |pole angle| > 0.2;
  the action is rescaled from [-1, 1] to the actuator's ctrl range."""

  def __init__(self, backend='generalized', n_frames=2, **kwargs):
    _check_backend(backend)
    spec = _spec(native.ENV_CARTPOLE)
    spec.healthy_angle_max = 0.2
    super().__init__(envs_assets.load('inverted_pendulum'), spec, (), n_frames, **kwargs)

  def _reset_q_qd(self, env_begin, n, seed, device):
    # q = init_q + U(-0.01, 0.01); qd = U(-0.01, 0.01)   (inverted_pendulum.py:116-122)
    init_q = torch.as_tensor(np.asarray(self.sys.init_q, np.float32), device=device)
    q = init_q[None] + sharding.uniform(env_begin, n, self.sys.nq, seed, 1, -0.01, 0.01, device)
    qd = sharding.uniform(env_begin, n, self.sys.nv, seed, 2, -0.01, 0.01, device)
    return q.contiguous(), qd.contiguous()

  # info['time_out'] = done lets PPO bootstrap on time-outs (inverted_pendulum.py:128,146)
  def reset(self, rng) -> State:
    st = super().reset(rng)
    st.info['time_out'] = st.done.clone()
    return st

  def step(self, state: State, action: torch.Tensor) -> State:
    st = super().step(state, action)
    st.info['time_out'] = st.done
    return st


class InvertedDoublePendulum(FusedEnv):
  """Reference envs/inverted_double_pendulum.py:129-195: alive bonus 10 minus distance and velocity
  penalties of the tip (0.6 m up the second pole); done when the tip is at or below z = 1."""

  def __init__(self, backend='generalized', n_frames=2, **kwargs):
    _check_backend(backend)
    spec = _spec(native.ENV_DOUBLE_CARTPOLE)
    spec.healthy_reward = 10.0     # alive_bonus
    spec.healthy_z_min = 1.0       # done = y <= 1
    spec.tip_link = 2
    spec.tip_pos[0], spec.tip_pos[1], spec.tip_pos[2] = 0.0, 0.0, 0.6
    super().__init__(envs_assets.load('inverted_double_pendulum'), spec, (), n_frames, **kwargs)

  def _reset_q_qd(self, env_begin, n, seed, device):
    # q = init_q + U(-0.01, 0.01); qd = 0.01 * N(0, 1)   (inverted_double_pendulum.py:148-152)
    init_q = torch.as_tensor(np.asarray(self.sys.init_q, np.float32), device=device)
    q = init_q[None] + sharding.uniform(env_begin, n, self.sys.nq, seed, 1, -0.01, 0.01, device)
    qd = 0.01 * sharding.normal(env_begin, n, self.sys.nv, seed, 2, device)
    return q.contiguous(), qd.contiguous()


class Reacher(FusedEnv):
  """Reference envs/reacher.py:157-248: reward = -|tip - target| - sum(action^2); the target is a
  2-dof link whose q is drawn at reset and never actuated."""

  def __init__(self, backend='generalized', n_frames=2, **kwargs):
    _check_backend(backend)
    spec = _spec(native.ENV_REACHER)
    spec.tip_link, spec.target_link = 1, 2
    spec.tip_pos[0], spec.tip_pos[1], spec.tip_pos[2] = 0.11, 0.0, 0.0
    super().__init__(envs_assets.load('reacher'), spec, ('reward_dist', 'reward_ctrl'), n_frames, **kwargs)

  def _reset_q_qd(self, env_begin, n, seed, device):
    # q = init_q + U(-0.1, 0.1); qd = U(-0.005, 0.005); q[2:] = target, qd[2:] = 0   (reacher.py:177-188);
    # target = dist * (cos, sin)(ang), dist = 0.2 U, ang = 2 pi U   (reacher.py:241-248)
    init_q = torch.as_tensor(np.asarray(self.sys.init_q, np.float32), device=device)
    q = init_q[None] + sharding.uniform(env_begin, n, self.sys.nq, seed, 1, -0.1, 0.1, device)
    qd = sharding.uniform(env_begin, n, self.sys.nv, seed, 2, -0.005, 0.005, device)
    dist = 0.2 * sharding.uniform(env_begin, n, 1, seed, 3, 0.0, 1.0, device)
    ang = (math.pi * 2.0) * sharding.uniform(env_begin, n, 1, seed, 4, 0.0, 1.0, device)
    q[:, 2:] = torch.cat([dist * torch.cos(ang), dist * torch.sin(ang)], 1)
    qd[:, 2:] = 0.0
    return q.contiguous(), qd.contiguous()


_SWIMMER_METRICS = ('reward_fwd', 'reward_ctrl', 'x_position', 'y_position', 'distance_from_origin',
                    'x_velocity', 'y_velocity', 'forward_reward')
_SWIMMER_SLOTS = {'reward_fwd': 0, 'reward_ctrl': 2, 'x_position': 4, 'y_position': 5, 'distance_from_origin': 6,
                  'x_velocity': 7, 'y_velocity': 8, 'forward_reward': 9}


class Swimmer(FusedEnv):
  """Constructor arguments as reference envs/swimmer.py:110-136.  The model moves through a
  viscous, dense medium: `sys.enable_fluid` (swimmer.xml option density / viscosity)."""

  def __init__(self, forward_reward_weight=1.0, ctrl_cost_weight=1e-4, reset_noise_scale=0.1,
               exclude_current_positions_from_observation=True, backend='generalized', n_frames=4, **kwargs):
    _check_backend(backend)
    spec = _spec(native.ENV_SWIMMER)
    spec.obs_skip = 2 if exclude_current_positions_from_observation else 0
    spec.forward_reward_weight = forward_reward_weight
    spec.ctrl_cost_weight = ctrl_cost_weight
    self._reset_noise_scale = reset_noise_scale
    super().__init__(envs_assets.load('swimmer'), spec, _SWIMMER_METRICS, n_frames, metric_slots=_SWIMMER_SLOTS, **kwargs)

  def _reset_q_qd(self, env_begin, n, seed, device):
    # qpos = init_q + U(-s, s); qvel = U(-s, s)   (swimmer.py:140-142, 196-198)
    s = self._reset_noise_scale
    init_q = torch.as_tensor(np.asarray(self.sys.init_q, np.float32), device=device)
    q = init_q[None] + sharding.uniform(env_begin, n, self.sys.nq, seed, 1, -s, s, device)
    qd = sharding.uniform(env_begin, n, self.sys.nv, seed, 2, -s, s, device)
    return q.contiguous(), qd.contiguous()


class HumanoidStandup(FusedEnv):
  """Reference envs/humanoidstandup.py:180-274: the humanoid starts lying down; reward = torso z / dt + 1
  - 0.01 * sum(action^2), never done.  15 floor contacts (torso, thighs, feet spheres and the limbs'
  capsule ends) make 77 constraint rows: the generic kernel variant."""

  def __init__(self, backend='generalized', n_frames=5, **kwargs):
    _check_backend(backend)
    spec = _spec(native.ENV_STANDUP)
    spec.obs_skip = 2
    spec.ctrl_cost_weight = 0.01
    spec.healthy_reward = 1.0
    super().__init__(envs_assets.load('humanoidstandup'), spec, ('reward_linup', 'reward_quadctrl'), n_frames, **kwargs)

  def _reset_q_qd(self, env_begin, n, seed, device):
    # qpos = init_q + U(-0.01, 0.01); qvel = U(-0.01, 0.01)   (humanoidstandup.py:202-209)
    init_q = torch.as_tensor(np.asarray(self.sys.init_q, np.float32), device=device)
    q = init_q[None] + sharding.uniform(env_begin, n, self.sys.nq, seed, 1, -0.01, 0.01, device)
    qd = sharding.uniform(env_begin, n, self.sys.nv, seed, 2, -0.01, 0.01, device)
    return q.contiguous(), qd.contiguous()


class Pusher(FusedEnv):
  """Reference envs/pusher.py:143-237: a 7-dof arm pushes a capsule towards a goal on a table.  The
  wrist capsules collide with the table (plane-capsule) and with the object (capsule-capsule, both
  links moving).  Reward (from the state BEFORE the step): -|object - goal| - 0.1 * sum(action^2)
  - 0.5 * |object - tips_arm|."""

  def __init__(self, backend='generalized', n_frames=5, **kwargs):
    _check_backend(backend)
    sys = envs_assets.load('pusher')
    names = list(sys.link_names)
    spec = _spec(native.ENV_PUSHER)
    # tips_arm is fused into r_wrist_roll_link; the reference uses the parent r_wrist_flex_link (pusher.py:158-160)
    spec.tip_link, spec.object_link, spec.target_link = names.index('r_wrist_flex_link'), names.index('object'), names.index('goal')
    super().__init__(sys, spec, ('reward_dist', 'reward_ctrl', 'reward_near'), n_frames, **kwargs)

  def _reset_q_qd(self, env_begin, n, seed, device):
    # cylinder_pos = [U(-0.3, -1e-6), U(-0.2, 0.2)], pushed out to >= 0.17 from the goal at the origin;
    # q[-4:] = [cylinder_pos, goal_pos]; qd = U(-0.005, 0.005), qd[-4:] = 0   (pusher.py:166-188)
    q = torch.as_tensor(np.asarray(self.sys.init_q, np.float32), device=device)[None].repeat(n, 1)
    cyl = torch.cat([sharding.uniform(env_begin, n, 1, seed, 1, -0.3, -1e-6, device),
                     sharding.uniform(env_begin, n, 1, seed, 2, -0.2, 0.2, device)], 1)
    norm = torch.linalg.norm(cyl, dim=1, keepdim=True)
    cyl = cyl * torch.where(norm < 0.17, 0.17 / norm, torch.ones_like(norm))
    q[:, -4:-2] = cyl
    q[:, -2:] = 0.0
    qd = sharding.uniform(env_begin, n, self.sys.nv, seed, 3, -0.005, 0.005, device)
    qd[:, -4:] = 0.0
    return q.contiguous(), qd.contiguous()
