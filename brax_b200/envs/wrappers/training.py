"""`brax.envs.wrappers.training.wrap` for fused envs (reference wrappers/training.py:28-57).

The reference stacks VmapWrapper / EpisodeWrapper / AutoResetWrapper objects around an env.
Here the env is natively batched and the Episode / AutoReset arithmetic runs inside the step
kernel, so `wrap` returns a copy of the env with those switches on."""
import copy
from typing import Callable, Optional, Tuple


from brax_b200 import base
from brax_b200.envs.base import FusedEnv


def wrap(env: FusedEnv, episode_length: int = 1000, action_repeat: int = 1, randomization_fn=None,
         batch_size: Optional[int] = None) -> FusedEnv:
  """Episode bookkeeping + auto-reset, as training.wrap applies them (reference :28-57)."""
  out = copy.copy(env) if randomization_fn is None else DomainRandomizationVmapWrapper(env, randomization_fn)
  out.spec = type(env.spec).from_buffer_copy(env.spec)     # the spec carries episode_length
  out.episode_length = int(episode_length)
  out.spec.episode_length = int(episode_length)
  out.auto_reset = True
  out.action_repeat = int(action_repeat)
  if batch_size is not None:
    if out.systems is not None and batch_size != len(out.systems):
      raise ValueError(f'batch_size {batch_size} != {len(out.systems)} randomised Systems')
    out.batch_size = batch_size
  return out


def DomainRandomizationVmapWrapper(env: FusedEnv,
                                   randomization_fn: Callable[[base.System], Tuple[base.System, base.System]]) -> FusedEnv:
  """Wrapper for domain randomization (reference wrappers/training.py:223-260).

  `randomization_fn(sys) -> (sys_v, in_axes)` as in the reference: `sys_v` is the env's System with a leading env axis on
  the randomised leaves, `in_axes` a System-shaped tree holding 0 at those leaves and None elsewhere.  The reference
  vmaps `env.step` over (`sys_v`, state, action); here env e of the batch reads the constants of its own model
  (`bxg_model_create_batched`, include/bxg.h) inside the same fused launch.  The randomised leaves must not change the
  topology, the time step, gravity or the solver iteration counts, nor `init_q` (the reset noise is added to the nominal
  `init_q`).  Returns a copy of the env whose batch size is the number of Systems."""
  sys_v, in_axes = randomization_fn(env.sys)
  if in_axes is not None and getattr(in_axes, 'init_q', None) is not None:
    raise NotImplementedError('randomising init_q is not supported: reset adds its noise to the nominal init_q')
  systems = base.unbatch(sys_v, in_axes)
  out = copy.copy(env)
  out.systems = [s.cast() if hasattr(s, 'cast') else s for s in systems]
  out.batch_size = len(systems)
  out._batched_models = {}
  return out
