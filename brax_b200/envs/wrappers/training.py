"""`brax.envs.wrappers.training.wrap` for fused envs (reference wrappers/training.py:28-57).

The reference stacks VmapWrapper / EpisodeWrapper / AutoResetWrapper objects around an env.
Here the env is natively batched and the Episode / AutoReset arithmetic runs inside the step
kernel, so `wrap` returns a copy of the env with those switches on."""
import copy
from typing import Optional

from brax_b200.envs.base import FusedEnv


def wrap(env: FusedEnv, episode_length: int = 1000, action_repeat: int = 1, randomization_fn=None,
         batch_size: Optional[int] = None) -> FusedEnv:
  """Episode bookkeeping + auto-reset, as training.wrap applies them (reference :28-57)."""
  if randomization_fn is not None:
    raise NotImplementedError('domain randomisation needs a per-env System, which the kernel does not take')
  out = copy.copy(env)
  out.spec = type(env.spec).from_buffer_copy(env.spec)     # the spec carries episode_length
  out.episode_length = int(episode_length)
  out.spec.episode_length = int(episode_length)
  out.auto_reset = True
  out.action_repeat = int(action_repeat)
  if batch_size is not None:
    out.batch_size = batch_size
  return out
