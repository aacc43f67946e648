"""NumPy restatement of the reference envs around the physics oracle.

TEST INFRASTRUCTURE ONLY (see oracle/bxg_oracle.c header).  Pinned against the
reference's own envs + training wrappers run on NumPy (tests/golden/ref_env_*.npz,
tests/test_reference_golden.py).  Follows, statement by statement, float32:
  brax/envs/ant.py:233-279            Ant.step / _get_obs
  brax/envs/humanoid.py:256-354       Humanoid.step / _get_obs / _com
  brax/envs/half_cheetah.py:178-212   Halfcheetah.step / _get_obs
  brax/envs/hopper.py:219-276         Hopper.step / _get_obs
  brax/envs/walker2d.py:200-273       Walker2d.step / _get_obs
  brax/envs/inverted_pendulum.py:131-154         InvertedPendulum.step / _get_obs
  brax/envs/inverted_double_pendulum.py:161-195  InvertedDoublePendulum.step / _get_obs
  brax/envs/reacher.py:199-239        Reacher.step / _get_obs
  brax/envs/swimmer.py:157-194        Swimmer.step / _get_obs
  brax/envs/humanoidstandup.py:220-274  HumanoidStandup.step / _get_obs
  brax/envs/pusher.py:195-237         Pusher.step / _get_obs
  brax/envs/wrappers/training.py:98-158  EpisodeWrapper.step, AutoResetWrapper.step
  brax/actuator.py:23-57              to_tau (for Humanoid's qfrc_actuator)
"""
from __future__ import annotations

import numpy as np

from oracle import oracle as O

f32 = np.float32


def _rotate(v, q):
  s, u = q[..., 0:1], q[..., 1:]
  r = f32(2) * (np.sum(u * v, -1, keepdims=True) * u) + (s * s - np.sum(u * u, -1, keepdims=True)) * v
  return r + f32(2) * s * np.cross(u, v)


class EnvOracle:
  def __init__(self, sys, kind, *, forward_reward_weight=1.0, ctrl_cost_weight=0.0, healthy_reward=0.0,
               terminate_when_unhealthy=True, healthy_z_range=(-np.inf, np.inf), exclude_current_positions=True,
               n_frames=5, episode_length=None, auto_reset=False,
               healthy_angle_range=(-np.inf, np.inf), healthy_state_range=(-np.inf, np.inf)):
    self.sys, self.kind = sys, kind
    self.o = O.Oracle(sys, np.float32)
    self.w_fwd, self.w_ctrl, self.h_rew = f32(forward_reward_weight), f32(ctrl_cost_weight), f32(healthy_reward)
    self.term = terminate_when_unhealthy
    self.zmin, self.zmax = f32(healthy_z_range[0]), f32(healthy_z_range[1])
    self.planar = kind in ('hopper', 'walker2d')
    self.skip = (1 if kind == 'halfcheetah' or self.planar else 2) if exclude_current_positions else 0
    self.classic = kind in ('inverted_pendulum', 'inverted_double_pendulum', 'reacher', 'swimmer')
    if self.classic and kind != 'swimmer':
      self.skip = 0
    self.amin, self.amax = f32(healthy_angle_range[0]), f32(healthy_angle_range[1])
    self.smin, self.smax = f32(healthy_state_range[0]), f32(healthy_state_range[1])
    self.n_frames = n_frames
    self.dt = f32(sys.opt.timestep) * f32(n_frames)
    self.episode_length, self.auto_reset = episode_length, auto_reset
    self.mass = np.asarray(sys.link.inertia.mass, f32)
    self.ipos = np.asarray(sys.link.inertia.transform.pos, f32)

  # -- helpers ---------------------------------------------------------------
  def _com(self, st):
    x_i = st['x_pos'] + _rotate(self.ipos[None], st['x_rot'])
    mass_sum = f32(0)
    for m in self.mass:
      mass_sum = f32(mass_sum + m)
    acc = np.zeros((st['x_pos'].shape[0], 3), f32)
    for l in range(len(self.mass)):
      acc = (acc + self.mass[l] * x_i[:, l]).astype(f32)
    return (acc / mass_sum).astype(f32), mass_sum, x_i

  def _to_tau(self, act, q, qd):
    a = self.sys.actuator
    n = q.shape[0]
    tau = np.zeros((n, self.sys.nv), f32)
    if self.sys.nu == 0:
      return tau
    qv, qdv = q[:, np.asarray(a.q_id)], qd[:, np.asarray(a.qd_id)]
    act = np.clip(act, a.ctrl_range[:, 0], a.ctrl_range[:, 1]).astype(f32)
    bias = (a.gear * (qv * a.bias_q + qdv * a.bias_qd)).astype(f32)
    force = (a.gain * act + bias).astype(f32)
    force = np.clip(force, a.force_range[:, 0], a.force_range[:, 1]).astype(f32)
    force = (force * a.gear).astype(f32)
    np.add.at(tau, (slice(None), np.asarray(a.qd_id)), force)
    return tau

  def _tip(self, st, link, p):
    return (st['x_pos'][:, link] + _rotate(np.asarray(p, f32)[None], st['x_rot'][:, link])).astype(f32)

  def _pusher_links(self):
    names = list(self.sys.link_names)
    return names.index('r_wrist_flex_link'), names.index('object'), names.index('goal')

  @staticmethod
  def _safe_norm(v):
    zero = np.all(np.abs(v) <= 1e-8, axis=1)
    return np.where(zero, f32(0), np.sqrt(np.sum(np.where(zero[:, None], f32(1), v) ** 2, -1))).astype(f32)

  def obs(self, st, action):
    if self.kind == 'pusher':   # [q[:7], qd[:7], x_i.pos[tips_arm], x_i.pos[object], x_i.pos[goal]]
      tip, obj, goal = self._pusher_links()
      x_i = (st['x_pos'] + _rotate(self.ipos[None], st['x_rot'])).astype(f32)
      return np.concatenate([st['q'][:, :7], st['qd'][:, :7], x_i[:, tip], x_i[:, obj], x_i[:, goal]], 1).astype(f32)
    if self.kind == 'inverted_double_pendulum':   # [q[:1], sin(q[1:]), cos(q[1:]), clip(qd, -10, 10)]
      q = st['q']
      return np.concatenate([q[:, :1], np.sin(q[:, 1:]), np.cos(q[:, 1:]), np.clip(st['qd'], f32(-10), f32(10))], 1).astype(f32)
    if self.kind == 'reacher':   # [cos(theta), sin(theta), q[2:], tip_vel[:2], tip_pos - target_pos]
      theta = st['q'][:, :2]
      tip = self._tip(st, 1, (0.11, 0, 0))
      p = np.asarray((0.11, 0, 0), f32)[None]
      tip_vel = (st['xd_vel'][:, 1] - np.cross(p, st['xd_ang'][:, 1])).astype(f32)
      return np.concatenate([np.cos(theta), np.sin(theta), st['q'][:, 2:], tip_vel[:, :2], tip - st['x_pos'][:, 2]], 1).astype(f32)
    if self.planar:   # position = q.at[1].set(x.pos[0, 2]); velocity = clip(qd, -10, 10)
      pos = st['q'].copy(); pos[:, 1] = st['x_pos'][:, 0, 2]
      return np.concatenate([pos[:, self.skip:], np.clip(st['qd'], f32(-10), f32(10))], 1).astype(f32)
    q, qd = st['q'][:, self.skip:], st['qd']
    if self.kind in ('ant', 'halfcheetah', 'inverted_pendulum', 'swimmer'):
      return np.concatenate([q, qd], 1).astype(f32)
    n, L = q.shape[0], len(self.mass)
    com, mass_sum, x_i = self._com(st)
    com_inertia = np.concatenate([st['cinr_i'].reshape(n, L, 9), np.broadcast_to(self.mass[None, :, None], (n, L, 1))], 2)
    off = (x_i - st['x_pos']).astype(f32)
    vel = (st['xd_vel'] - np.cross(off, st['xd_ang'])).astype(f32)
    com_vel = (self.mass[None, :, None] * vel / mass_sum).astype(f32)
    com_velocity = np.concatenate([com_vel, st['xd_ang']], 2)
    tau = self._to_tau(action, st['q'], st['qd'])
    return np.concatenate([q, qd, com_inertia.reshape(n, -1), com_velocity.reshape(n, -1), tau], 1).astype(f32)

  # -- env API ---------------------------------------------------------------
  def reset(self, q, qd):
    st = self.o.init(q, qd)
    n = st['q'].shape[0]
    ob = self.obs(st, np.zeros((n, self.sys.nu), f32))
    env = {'ps': st, 'obs': ob, 'reward': np.zeros(n, f32), 'done': np.zeros(n, f32), 'metrics': {},
           'steps': np.zeros(n, f32), 'truncation': np.zeros(n, f32)}
    if self.auto_reset:
      env['first_ps'] = {k: v.copy() for k, v in st.items()}
      env['first_obs'] = ob.copy()
    return env

  def _env_step(self, ps0, action):
    """Ant.step / Humanoid.step on pipeline states (no wrappers)."""
    action = np.asarray(action, f32)
    if self.classic:
      return self._classic_step(ps0, action)
    if self.kind == 'pusher':
      lo, hi = self.sys.actuator.ctrl_range[:, 0], self.sys.actuator.ctrl_range[:, 1]
      action = ((action + f32(1)) * (hi - lo) * f32(0.5) + lo).astype(f32)
      ps = {k: v.copy() for k, v in ps0.items()}
      self.o.step(ps, action, self.n_frames)
      tip, obj, goal = self._pusher_links()
      x_i = (ps0['x_pos'] + _rotate(self.ipos[None], ps0['x_rot'])).astype(f32)   # the PRE-step state (pusher.py:204-207)
      near = -self._safe_norm(x_i[:, obj] - x_i[:, tip])
      dist = -self._safe_norm(x_i[:, obj] - x_i[:, goal])
      sq = np.zeros(action.shape[0], f32)
      for a in range(action.shape[1]):
        sq = (sq + action[:, a] * action[:, a]).astype(f32)
      reward = ((dist + f32(0.1) * -sq).astype(f32) + f32(0.5) * near).astype(f32)
      return ps, self.obs(ps, action), reward, np.zeros_like(reward), {'reward_dist': dist, 'reward_ctrl': -sq, 'reward_near': near}
    if self.kind == 'humanoidstandup':
      lo, hi = self.sys.actuator.ctrl_range[:, 0], self.sys.actuator.ctrl_range[:, 1]
      action = ((action + f32(1)) * (hi - lo) * f32(0.5) + lo).astype(f32)
      ps = {k: v.copy() for k, v in ps0.items()}
      self.o.step(ps, action, self.n_frames)
      uph = ((ps['x_pos'][:, 0, 2] - f32(0)) / self.dt).astype(f32)
      sq = np.zeros(action.shape[0], f32)
      for a in range(action.shape[1]):
        sq = (sq + action[:, a] * action[:, a]).astype(f32)
      quad = (f32(0.01) * sq).astype(f32)
      reward = ((uph + f32(1)).astype(f32) - quad).astype(f32)
      return ps, self.obs(ps, action), reward, np.zeros_like(reward), {'reward_linup': uph, 'reward_quadctrl': -quad}
    if self.kind == 'humanoid':
      lo, hi = self.sys.actuator.ctrl_range[:, 0], self.sys.actuator.ctrl_range[:, 1]
      action = ((action + f32(1)) * (hi - lo) * f32(0.5) + lo).astype(f32)
    ps = {k: v.copy() for k, v in ps0.items()}
    if self.kind == 'humanoid':
      before, _, _ = self._com(ps0)
    else:
      before = ps0['x_pos'][:, 0].copy()
    self.o.step(ps, action, self.n_frames)
    after = self._com(ps)[0] if self.kind == 'humanoid' else ps['x_pos'][:, 0]
    velocity = ((after - before) / self.dt).astype(f32)
    forward = velocity[:, 0] if self.kind == 'ant' else (self.w_fwd * velocity[:, 0]).astype(f32)
    z = ps['x_pos'][:, 0, 2]
    healthy = np.where(z < self.zmin, f32(0), f32(1)).astype(f32)
    healthy = np.where(z > self.zmax, f32(0), healthy).astype(f32)
    if self.planar:   # strict ranges on z, angle = q[2] and the state vector [q[2:], qd]
      angle = ps['q'][:, 2]
      sv = np.concatenate([ps['q'][:, 2:], ps['qd']], 1)
      ok = (self.zmin < z) & (z < self.zmax) & (self.amin < angle) & (angle < self.amax)
      ok &= np.all((self.smin < sv) & (sv < self.smax), axis=1)
      healthy = ok.astype(f32)
    h_rew = np.full_like(healthy, self.h_rew) if self.term else (self.h_rew * healthy).astype(f32)
    sq = np.zeros(action.shape[0], f32)
    for a in range(action.shape[1]):
      sq = (sq + action[:, a] * action[:, a]).astype(f32)
    ctrl = (self.w_ctrl * sq).astype(f32)
    reward = ((forward + h_rew).astype(f32) - ctrl).astype(f32)
    done = (f32(1) - healthy).astype(f32) if self.term else np.zeros_like(healthy)
    dist = np.sqrt(np.sum(after.astype(f32) ** 2, -1)).astype(f32)
    if self.kind == 'halfcheetah':   # reward = forward_reward - ctrl_cost; done is never set (half_cheetah.py:189-199)
      reward = (forward - ctrl).astype(f32)
      m = {'x_position': after[:, 0], 'x_velocity': velocity[:, 0], 'reward_ctrl': -ctrl, 'reward_run': forward}
    elif self.planar:
      m = {'reward_forward': forward, 'reward_ctrl': -ctrl, 'reward_healthy': h_rew, 'x_position': after[:, 0], 'x_velocity': velocity[:, 0]}
    elif self.kind == 'ant':
      m = {'reward_forward': forward, 'reward_survive': h_rew, 'reward_ctrl': -ctrl, 'reward_contact': np.zeros_like(ctrl),
           'x_position': after[:, 0], 'y_position': after[:, 1], 'distance_from_origin': dist,
           'x_velocity': velocity[:, 0], 'y_velocity': velocity[:, 1], 'forward_reward': forward}
    else:
      m = {'forward_reward': forward, 'reward_linvel': forward, 'reward_quadctrl': -ctrl, 'reward_alive': h_rew,
           'x_position': after[:, 0], 'y_position': after[:, 1], 'distance_from_origin': dist,
           'x_velocity': velocity[:, 0], 'y_velocity': velocity[:, 1]}
    return ps, self.obs(ps, action), reward, done, m

  def _classic_step(self, ps0, action):
    """InvertedPendulum / InvertedDoublePendulum / Reacher / Swimmer .step on pipeline states."""
    if self.kind == 'inverted_pendulum':
      lo, hi = self.sys.actuator.ctrl_range[:, 0], self.sys.actuator.ctrl_range[:, 1]
      action = ((action + f32(1)) * (hi - lo) * f32(0.5) + lo).astype(f32)
    ps = {k: v.copy() for k, v in ps0.items()}
    self.o.step(ps, action, self.n_frames)
    n = action.shape[0]
    ob = self.obs(ps, action)
    sq = np.zeros(n, f32)
    for a in range(action.shape[1]):
      sq = (sq + action[:, a] * action[:, a]).astype(f32)
    m = {}
    if self.kind == 'inverted_pendulum':
      reward = np.ones(n, f32)
      done = np.where(np.abs(ob[:, 1]) > f32(0.2), f32(1), f32(0)).astype(f32)
    elif self.kind == 'inverted_double_pendulum':
      tip = self._tip(ps, 2, (0, 0, 0.6))
      x, y = tip[:, 0], tip[:, 2]
      v1, v2 = ps['qd'][:, 1], ps['qd'][:, 2]
      dist_penalty = (f32(0.01) * x ** 2 + (y - f32(2)) ** 2).astype(f32)
      vel_penalty = (f32(1e-3) * v1 ** 2 + f32(5e-3) * v2 ** 2).astype(f32)
      done = np.where(y <= 1, f32(1), f32(0)).astype(f32)
      reward = (((f32(1) - done) * f32(10) - dist_penalty) - vel_penalty).astype(f32)
    elif self.kind == 'reacher':
      v = ob[:, -3:]
      zero = np.all(np.abs(v) <= 1e-8, axis=1)           # math.safe_norm
      norm = np.where(zero, f32(0), np.sqrt(np.sum(np.where(zero[:, None], f32(1), v) ** 2, -1))).astype(f32)
      m = {'reward_dist': -norm, 'reward_ctrl': -sq}
      reward = (m['reward_dist'] + m['reward_ctrl']).astype(f32)
      done = np.zeros(n, f32)
    else:   # swimmer
      xy = ps['q'][:, :2]
      vel = ((xy - ps0['q'][:, :2]) / self.dt).astype(f32)
      forward = (self.w_fwd * vel[:, 0]).astype(f32)
      ctrl = (self.w_ctrl * sq).astype(f32)
      reward = (forward - ctrl).astype(f32)
      done = np.zeros(n, f32)
      m = {'reward_fwd': forward, 'reward_ctrl': -ctrl, 'x_position': xy[:, 0], 'y_position': xy[:, 1],
           'distance_from_origin': np.sqrt(xy[:, 0] ** 2 + xy[:, 1] ** 2).astype(f32), 'x_velocity': vel[:, 0], 'y_velocity': vel[:, 1],
           'forward_reward': np.zeros(n, f32)}   # swimmer.py never updates it after reset
    return ps, ob, reward, done, m

  def step(self, env, action):
    """AutoReset(Episode(env)).step with action_repeat = 1."""
    steps = np.where(env['done'] != 0, f32(0), env['steps']).astype(f32) if self.episode_length else env['steps']
    ps, ob, reward, done, m = self._env_step(env['ps'], action)
    trunc = np.zeros_like(done)
    if self.episode_length:
      steps = (steps + f32(1)).astype(f32)
      over = steps >= self.episode_length
      trunc = np.where(over, f32(1) - done, f32(0)).astype(f32)
      done = np.where(over, f32(1), done).astype(f32)
    out = dict(env)
    if self.auto_reset:
      d = done != 0
      ps = {k: (np.where(d.reshape((-1,) + (1,) * (v.ndim - 1)), env['first_ps'][k], v) if k in env['first_ps'] and v.dtype != np.int32 else v)
            for k, v in ps.items()}
      ob = np.where(d[:, None], env['first_obs'], ob)
    out.update(ps=ps, obs=ob, reward=reward, done=done, metrics=m, steps=steps, truncation=trunc)
    return out
