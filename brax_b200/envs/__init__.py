"""Environments in scope (reference `brax/envs/__init__.py:35-107`): ant, humanoid, halfcheetah, hopper, walker2d,
inverted_pendulum, inverted_double_pendulum, reacher, swimmer, humanoidstandup, pusher (`fast` is a toy without physics and is not mirrored)."""
from typing import Optional

from brax_b200.envs.ant import Ant
from brax_b200.envs.base import FusedEnv, State
from brax_b200.envs.classic import HumanoidStandup, InvertedDoublePendulum, InvertedPendulum, Pusher, Reacher, Swimmer
from brax_b200.envs.half_cheetah import Halfcheetah
from brax_b200.envs.hopper import Hopper, Walker2d
from brax_b200.envs.humanoid import Humanoid

_envs = {'ant': Ant, 'humanoid': Humanoid, 'halfcheetah': Halfcheetah, 'hopper': Hopper, 'walker2d': Walker2d,
         'inverted_pendulum': InvertedPendulum, 'inverted_double_pendulum': InvertedDoublePendulum,
         'reacher': Reacher, 'swimmer': Swimmer, 'humanoidstandup': HumanoidStandup, 'pusher': Pusher}


def get_environment(env_name: str, **kwargs) -> FusedEnv:
  """Returns an environment from the registry (reference envs/__init__.py:51-63)."""
  return _envs[env_name](**kwargs)


def create(env_name: str, episode_length: int = 1000, action_repeat: int = 1, auto_reset: bool = True,
           batch_size: Optional[int] = None, **kwargs) -> FusedEnv:
  """reference envs/__init__.py:76-107.  The Episode / Vmap / AutoReset wrappers
  are not separate objects here: their arithmetic is fused into the step kernel."""
  return _envs[env_name](episode_length=episode_length, auto_reset=auto_reset, batch_size=batch_size,
                         action_repeat=action_repeat, **kwargs)
