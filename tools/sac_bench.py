"""End-to-end SAC on Ant with the fused B200 env step (acting + replay + learner), env-steps/sec.
Hyper-parameters of the reference's SAC example for Ant-class envs (notebooks/training.ipynb): 128 envs,
batch 512, grad_updates_per_step 32, discounting 0.97, reward_scaling 30, lr 6e-4, min_replay_size 8192.
  python tools/sac_bench.py [env_steps] [env]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brax_b200.training import sac  # noqa: E402

if __name__ == '__main__':
  steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
  env = sys.argv[2] if len(sys.argv) > 2 else 'ant'
  log = []
  net, m = sac.train(env, num_timesteps=steps, episode_length=1000, num_envs=128, batch_size=512, grad_updates_per_step=32,
                     discounting=0.97, reward_scaling=30.0, learning_rate=6e-4, min_replay_size=8192, max_replay_size=1048576,
                     normalize_observations=True, progress_every=50,
                     progress_fn=lambda t, mm: log.append((t, round(mm['episode_reward'], 2), round(mm['sps']))))
  print(json.dumps({'metric': 'SAC env-steps/sec (acting + replay + learner)', 'value': m['sps'], 'env_steps': m['env_steps'],
                    'config': f'{env}, 128 envs, batch 512, 32 grad updates per step', 'alpha': m['alpha'], 'progress': log[-8:]}))
