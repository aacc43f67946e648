"""Physics pipeline for the generalized coordinates engine -- B200 backend.

Drop-in for reference `brax/generalized/pipeline.py`: same two functions, same
argument meaning, same `State` leaves.  Differences that follow from replacing a
traced JAX function by a natively batched CUDA kernel:

  * `q`, `qd`, `act` and every State leaf carry a leading env axis [n, ...]
    (unbatched inputs are accepted and get n = 1).  The reference reaches the
    same shapes through `jax.vmap(env.step)` (envs/wrappers/training.py:66-72).
  * `step(..., n_frames=k)` runs k substeps with one action inside one launch,
    which is what `PipelineEnv.pipeline_step` does with `lax.scan`
    (envs/base.py:128-137).
  * No VJP: `jax.grad` through `step` (APG, pipeline_test.py:50-79) is out of
    scope (SURVEY.md section 8b "What breaks").
"""
from __future__ import annotations

from typing import Optional

import torch

from brax_b200 import native
from brax_b200.base import System
from brax_b200.generalized.base import State


def _device_index(t: torch.Tensor) -> int:
  if not t.is_cuda:
    raise RuntimeError(
        'brax_b200.generalized.pipeline needs CUDA tensors: there is no CPU fallback')
  return t.device.index if t.device.index is not None else torch.cuda.current_device()


def _validate(sys: System) -> None:
  """The subset of `mjcf.validate_model` that applies (io/mjcf.py:236-314)."""
  if sys.enable_fluid and sys.num_links() > 32:
    raise NotImplementedError('fluid forces run on the generic kernel variant: at most 32 links')


def init(
    sys: System,
    q: torch.Tensor,
    qd: torch.Tensor,
    unused_act: Optional[torch.Tensor] = None,
    unused_ctrl: Optional[torch.Tensor] = None,
    debug: bool = False,
    minv_mode: int = native.MINV_NEWTON_SCHULZ,
) -> State:
  """Initializes physics state (reference pipeline.py:32-61).

  Args:
    sys: a brax_b200 System
    q: (n, q_size) or (q_size,) joint position vector(s), CUDA
    qd: (n, qd_size) or (qd_size,) joint velocity vector(s), CUDA
    debug: if True, adds contact distances to the state for debugging
  """
  _validate(sys)
  squeeze = q.dim() == 1
  if squeeze:
    q, qd = q[None], qd[None]
  model = native.model_for(sys, _device_index(q), minv_mode)
  bufs = model.init(q, qd)
  contact = None
  if debug:
    contact = _contact_debug(model, bufs)
  st = State.from_flat(bufs, contact)
  return _squeeze(st) if squeeze else st


def step(
    sys: System,
    state: State,
    act: Optional[torch.Tensor],
    debug: bool = False,
    n_frames: int = 1,
    minv_mode: int = native.MINV_NEWTON_SCHULZ,
) -> State:
  """Performs `n_frames` physics steps with one action (reference pipeline.py:64-94).

  Args:
    sys: a brax_b200 System
    state: physics state prior to step
    act: (n, act_size) or (act_size,) actuator input; None iff act_size == 0
    debug: if True, adds contact distances to the state for debugging
  """
  squeeze = state.q.dim() == 1
  if squeeze:
    state = _unsqueeze(state)
    act = None if act is None else act[None]
  model = native.model_for(sys, _device_index(state.q), minv_mode)
  bufs = {k: v.contiguous() for k, v in state.to_flat().items()}
  diag = model.alloc_diag(bufs['q'].shape[0]) if debug else None
  out = model.step(bufs, act, n_frames=n_frames, diag=diag)
  st = State.from_flat(out, diag if debug else None)
  return _squeeze(st) if squeeze else st


def _contact_debug(model, bufs):
  """debug=True after init: plane-sphere penetration distances of the fresh state
  (brax/contact.py:28-67 + mjx plane-sphere).  Debug-only host-side torch ops; the
  step kernel produces the same quantity itself (BXG_STEP_DIAGNOSTICS)."""
  import numpy as np
  n = bufs['q'].shape[0]
  diag = model.alloc_diag(n)
  cp = model.sys.contact_pairs()
  if len(cp.geom1) == 0:
    return diag
  dev = bufs['q'].device
  lb = torch.as_tensor(np.asarray(cp.link_b, np.int64), device=dev)
  pos, rot = bufs['x_pos'][:, lb], bufs['x_rot'][:, lb]          # [n, ncon, 3|4]
  v = torch.as_tensor(np.asarray(cp.sphere_pos, np.float32), device=dev)[None].expand_as(pos)
  s_, u = rot[..., :1], rot[..., 1:]
  r = 2 * ((u * v).sum(-1, keepdim=True) * u) + (s_ * s_ - (u * u).sum(-1, keepdim=True)) * v + 2 * s_ * torch.cross(u, v, dim=-1)
  sp = pos + r
  nrm = torch.as_tensor(np.asarray(cp.plane_normal, np.float32), device=dev)[None]
  pp = torch.as_tensor(np.asarray(cp.plane_pos, np.float32), device=dev)[None]
  rad = torch.as_tensor(np.asarray(cp.radius, np.float32), device=dev)[None]
  diag['con_dist'] = (((sp - pp) * nrm).sum(-1) - rad).contiguous()
  return diag


def _squeeze(st: State) -> State:
  from brax_b200.base import tree_map
  return tree_map(lambda x: x[0] if isinstance(x, torch.Tensor) else x, st)


def _unsqueeze(st: State) -> State:
  from brax_b200.base import tree_map
  return tree_map(lambda x: x[None] if isinstance(x, torch.Tensor) else x, st)
