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

from brax_b200 import contact as contact_lib
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
    debug: if True, adds the contacts of the fresh state (`brax_b200.contact.Contact`) for debugging
  """
  _validate(sys)
  squeeze = q.dim() == 1
  if squeeze:
    q, qd = q[None], qd[None]
  model = native.model_for(sys, _device_index(q), minv_mode)
  bufs = model.init(q, qd)
  st = State.from_flat(bufs, None)
  if debug:   # reference pipeline.py:58-60: contact.get(sys, x)
    st = st.replace(contact=contact_lib.get(sys, st.x))
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
    debug: if True, adds the new state's contacts and the step's solver counters (`brax_b200.contact.Contact`)
  """
  squeeze = state.q.dim() == 1
  if squeeze:
    state = _unsqueeze(state)
    act = None if act is None else act[None]
  model = native.model_for(sys, _device_index(state.q), minv_mode)
  bufs = {k: v.contiguous() for k, v in state.to_flat().items()}
  diag = model.alloc_diag(bufs['q'].shape[0]) if debug else None
  out = model.step(bufs, act, n_frames=n_frames, diag=diag)
  st = State.from_flat(out, None)
  if debug:   # reference pipeline.py:91-92; the distances are the ones the kernel's colliders produced for this state
    st = st.replace(contact=contact_lib.get(sys, st.x, kernel_dist=diag['con_dist'] if model.ncon else None, solver_stats=diag['stats']))
  return _squeeze(st) if squeeze else st


def _squeeze(st: State) -> State:
  from brax_b200.base import tree_map
  return tree_map(lambda x: x[0] if isinstance(x, torch.Tensor) else x, st)


def _unsqueeze(st: State) -> State:
  from brax_b200.base import tree_map
  return tree_map(lambda x: x[None] if isinstance(x, torch.Tensor) else x, st)
