"""Where PPO's time goes (config 5 shape): kernel-level summary from torch.profiler over a few training steps.
  python tools/ppo_profile.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from brax_b200.training import ppo

steps = 4096 * 5 * 16 * 6      # six training steps
ppo.train('ant', num_envs=4096, episode_length=1000, num_timesteps=4096 * 5 * 16 * 2, unroll_length=5, batch_size=2048,
          num_minibatches=32, num_update_epochs=4, reward_scaling=10.0)          # warm-up (graphs, allocator)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
  ppo.train('ant', num_envs=4096, episode_length=1000, num_timesteps=steps, unroll_length=5, batch_size=2048,
            num_minibatches=32, num_update_epochs=4, reward_scaling=10.0)
  torch.cuda.synchronize()
ev = prof.key_averages()
tot = sum(e.device_time_total for e in ev)
print(f'total device time {tot / 1e3:.1f} ms for {steps} env-steps -> device-bound rate {steps / (tot * 1e-6):.0f} env-steps/s')
for e in sorted(ev, key=lambda e: -e.device_time_total)[:14]:
  print(f'{100 * e.device_time_total / tot:5.1f}%  n={e.count:6d}  {e.key[:90]}')
