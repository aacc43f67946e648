"""Environment interface (reference `brax/envs/base.py:33-165`), natively batched.

`State` has the reference fields.  `FusedEnv.step` is one CUDA launch: the
n_frames physics substeps, the env's obs / reward / done / metrics and -- when
wrapped -- the EpisodeWrapper / AutoResetWrapper arithmetic
(`brax/envs/wrappers/training.py:75-158`) all happen while the env's state is in
shared memory.
"""
from __future__ import annotations

import dataclasses
from typing import Any, Dict, Optional

import numpy as np
import torch

from brax_b200 import base, native, sharding
from brax_b200.generalized.base import State as PipelineState


@dataclasses.dataclass(frozen=True)
class State(base.Base):
  """Environment state for training and inference (reference envs/base.py:33-43)."""
  pipeline_state: Optional[PipelineState]
  obs: Any
  reward: Any
  done: Any
  metrics: Dict[str, Any] = dataclasses.field(default_factory=dict)
  info: Dict[str, Any] = dataclasses.field(default_factory=dict)


class FusedEnv:
  """A brax-style env whose step runs fused on the GPU.

  Subclasses set `sys`, `spec` (native.EnvSpecC), `metric_names`, `n_frames`
  and implement `_reset_q_qd(env_begin, n, seed, device)`.
  """

  backend = 'generalized'

  def __init__(self, sys: base.System, spec: native.EnvSpecC, metric_names, n_frames: int,
               episode_length: Optional[int] = None, auto_reset: bool = False,
               batch_size: Optional[int] = None, device=None, env_id_offset: int = 0, metric_slots=None,
               action_repeat: int = 1, lean: bool = False):
    self.sys = sys
    self.systems = None   # one System per env (DomainRandomizationVmapWrapper, envs/wrappers/training.py), else None
    # lean: the pipeline state carries q, qd, x, xd and mass_mx_inv only (BXG_STEP_LEAN): the step recomputes the
    # rest in shared memory, bit-identically; a quarter of the State in HBM, what a trainer's rollout wants
    self.lean = bool(lean)
    self.spec = spec
    self.metric_names = tuple(metric_names)
    # slot of each metric in the kernel's per-env metrics row (default: in order)
    self._metric_slots = dict(metric_slots) if metric_slots else {k: i for i, k in enumerate(self.metric_names)}
    self._n_frames = int(n_frames)
    self.episode_length = episode_length
    self.auto_reset = auto_reset
    self.action_repeat = int(action_repeat)   # EpisodeWrapper's scan over env.step (wrappers/training.py:99-104)
    self.batch_size = batch_size
    self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    if self.device.type == 'cuda' and self.device.index is None:   # 'cuda' means the CURRENT device, not device 0
      self.device = torch.device('cuda', torch.cuda.current_device())
    self.env_id_offset = int(env_id_offset)   # global id of env 0 (sharding across ranks)
    spec.episode_length = int(episode_length) if episode_length else 0
    spec.env_dt = float(np.float32(sys.opt.timestep) * np.float32(self._n_frames))

  # -- reference properties ---------------------------------------------------
  @property
  def dt(self) -> float:
    return float(self.spec.env_dt)

  @property
  def n_frames(self) -> int:
    return self._n_frames

  @property
  def action_size(self) -> int:
    return self.sys.act_size()

  @property
  def observation_size(self) -> int:
    return self._model().env_obs_size(self.spec)

  @property
  def unwrapped(self) -> 'FusedEnv':
    return self

  def _model(self) -> native.NativeModel:
    if self.systems is not None:   # domain randomisation: env e steps with systems[e]
      m = self._batched_models.get(self.device.index)
      if m is None:
        m = self._batched_models[self.device.index] = native.BatchedNativeModel(self.systems, self.device.index)
      return m
    return native.model_for(self.sys, self.device.index)

  # -- API --------------------------------------------------------------------
  def reset(self, rng) -> State:
    """rng: an int seed.  Env e draws its noise from (seed, global env id)."""
    n = self.batch_size or 1
    seed = int(rng)
    q, qd = self._reset_q_qd(self.env_id_offset, n, seed, self.device)
    bufs, obs = self._model().env_reset(self.spec, q, qd)
    if self.lean:
      bufs = {k: v for k, v in bufs.items() if k in native.LEAN_FIELDS}
    zeros = lambda: torch.zeros(n, dtype=torch.float32, device=self.device)
    metrics = {k: zeros() for k in self.metric_names}
    info: Dict[str, Any] = {}
    if self.episode_length:
      info['steps'] = zeros(); info['truncation'] = zeros()
      # EpisodeWrapper.reset (wrappers/training.py:83-96)
      info['episode_done'] = zeros()
      info['episode_metrics'] = {k: zeros() for k in ('sum_reward', 'length') + self.metric_names}
    if self.auto_reset:
      info['first_pipeline_state'] = PipelineState.from_flat(bufs)
      info['first_obs'] = obs
    return State(PipelineState.from_flat({k: v.clone() for k, v in bufs.items()}) if self.auto_reset
                 else PipelineState.from_flat(bufs), obs.clone() if self.auto_reset else obs, zeros(), zeros(), metrics, info)

  def step(self, state: State, action: torch.Tensor) -> State:
    n = state.obs.shape[0]
    model = self._model()
    io = {
        'obs': torch.empty_like(state.obs), 'reward': torch.empty(n, dtype=torch.float32, device=self.device),
        'done': state.done.clone(),   # previous done in, new done out
        'metrics': torch.empty((n, native.ENV_NUM_METRICS), dtype=torch.float32, device=self.device),
        'steps': state.info['steps'].clone() if 'steps' in state.info else None,
        'truncation': torch.empty(n, dtype=torch.float32, device=self.device) if 'steps' in state.info else None,
    }
    first = first_obs = None
    if 'first_pipeline_state' in state.info:
      first = {k: v.contiguous() for k, v in state.info['first_pipeline_state'].to_flat().items() if v is not None}
      first_obs = state.info['first_obs'].contiguous()
    bufs = {k: v.contiguous() for k, v in state.pipeline_state.to_flat().items() if v is not None}
    reward_sum = None
    if self.action_repeat > 1:
      # EpisodeWrapper.step (wrappers/training.py:98-112): the inner env steps `action_repeat` times with one action,
      # rewards are summed, `steps` advances by action_repeat and the episode / auto-reset logic runs once, on the
      # last inner step.  The first action_repeat - 1 steps are bare env steps (no episode bookkeeping, no reset).
      bare = type(self.spec).from_buffer_copy(self.spec)
      bare.episode_length = 0
      scratch = {'obs': io['obs'], 'reward': torch.empty_like(io['reward']), 'done': torch.zeros_like(io['done']), 'metrics': io['metrics']}
      reward_sum = torch.zeros_like(io['reward'])
      for _ in range(self.action_repeat - 1):
        bufs = model.env_step(bare, bufs, action, self._n_frames, scratch, lean=self.lean)
        reward_sum += scratch['reward']
      if io['steps'] is not None:   # the kernel adds 1 after its own reset-on-previous-done: hand it that reset already applied
        io['steps'] = torch.where(io['done'] != 0, torch.zeros_like(io['steps']), io['steps']) + float(self.action_repeat - 1)
        io['done'] = torch.zeros_like(io['done'])
    out = model.env_step(self.spec, bufs, action, self._n_frames, io, first=first, first_obs=first_obs, lean=self.lean)
    if reward_sum is not None:
      io['reward'] = reward_sum + io['reward']
    metrics = {k: io['metrics'][:, self._metric_slots[k]] for k in self.metric_names}
    info = dict(state.info)
    if io['steps'] is not None:
      info['steps'] = io['steps']; info['truncation'] = io['truncation']
    if 'episode_metrics' in state.info:
      # EpisodeWrapper.step (wrappers/training.py:114-127): episode sums restart after an episode that ended
      keep = 1.0 - state.info['episode_done']
      em = state.info['episode_metrics']
      new = {'sum_reward': em['sum_reward'] * keep + io['reward'], 'length': em['length'] * keep + float(self.action_repeat)}
      for k in self.metric_names:
        new[k] = (em[k] * keep + metrics[k]) if k != 'reward' else em[k]
      info['episode_metrics'] = new
      info['episode_done'] = io['done']
    return State(PipelineState.from_flat(out), io['obs'], io['reward'], io['done'], metrics, info)

  def _reset_q_qd(self, env_begin, n, seed, device):
    raise NotImplementedError
