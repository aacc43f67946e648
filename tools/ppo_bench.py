"""BASELINE config 5: end-to-end PPO on Ant with the fused B200 env step.
Hyper-parameters of the reference colab (notebooks/training.ipynb:250): 4096 envs, unroll 5,
32 minibatches x batch 2048, 4 update epochs, reward_scaling 10, lr 3e-4, obs normalisation.
  python tools/ppo_bench.py [env_steps]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brax_b200.training import ppo  # noqa: E402

if __name__ == '__main__':
  steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3_000_000
  log = []
  import torch
  import torch.distributed as dist
  world = int(os.environ.get('WORLD_SIZE', '1'))
  if world > 1:   # torchrun: one process per GPU, env batch sharded, NCCL gradient all-reduce
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ['LOCAL_RANK'])))
  agent, m = ppo.train('ant', num_envs=4096, episode_length=1000, num_timesteps=steps, unroll_length=5,
                       batch_size=2048, num_minibatches=32, num_update_epochs=4, reward_scaling=10.0,
                       learning_rate=3e-4, entropy_cost=1e-2, discounting=0.97,
                       progress_fn=lambda t, mm: log.append((t, round(mm['episode_reward'], 2), round(mm['sps']))))
  if world == 1 or dist.get_rank() == 0:
    print(json.dumps({'metric': 'PPO env-steps/sec (rollout + policy + learner)', 'value': m['sps'], 'env_steps': m['env_steps'],
                      'n_gpus': world, 'config': 'ant, 4096 envs/GPU, unroll 5, 32x2048 minibatches, 4 epochs',
                      'progress': log[-8:]}))
  if world > 1:
    dist.destroy_process_group()
