"""SAC on the fused B200 env step (the second trainer `north_star` names as a consumer of the backend).

The physics + env arithmetic is the hand-written kernel (one launch per env step for the whole
batch); networks, replay and Adam are ordinary PyTorch.  Algorithm, losses and hyper-parameter
names follow the reference:
  brax/training/agents/sac/train.py:124-156     train() arguments (discounting 0.9, tau 0.005, ...)
  brax/training/agents/sac/losses.py:30-131     alpha / critic / actor losses, target_entropy = -0.5 * act,
                                                truncation-masked q error
  brax/training/agents/sac/networks.py:60-99    policy [obs, 256, 256, 2 * act] (tanh-normal), two
                                                q networks [obs + act, 256, 256, 1], relu
  brax/training/acting.py:33-62                 transition: discount = 1 - done, extras: truncation
  brax/training/replay_buffers.py:258-311       uniform sampling from a ring buffer

Multi-GPU: one process per GPU, env batch and replay sharded by rank, gradients averaged with an
NCCL all-reduce per update (`lax.pmean` in training/gradients.py:32).
"""
from __future__ import annotations

import math
import time
from typing import Callable, Dict, Optional

import torch
import torch.distributed as dist
from torch import nn

from brax_b200 import envs


def _mlp(sizes, act=nn.ReLU):
  layers = []
  for i in range(len(sizes) - 1):
    layers += [nn.Linear(sizes[i], sizes[i + 1]), act()]
  return nn.Sequential(*layers[:-1])


class SACNetworks(nn.Module):
  """reference sac/networks.py:60-99 (+ running observation statistics, acme/running_statistics.py)."""

  def __init__(self, obs_size: int, act_size: int, hidden=(256, 256), n_critics: int = 2, normalize_observations: bool = False):
    super().__init__()
    self.policy = _mlp([obs_size, *hidden, 2 * act_size])
    self.q = nn.ModuleList([_mlp([obs_size + act_size, *hidden, 1]) for _ in range(n_critics)])
    self.target_q = nn.ModuleList([_mlp([obs_size + act_size, *hidden, 1]) for _ in range(n_critics)])
    self.target_q.load_state_dict(self.q.state_dict())
    for p in self.target_q.parameters():
      p.requires_grad_(False)
    self.log_alpha = nn.Parameter(torch.zeros(()))
    self.normalize_observations = normalize_observations
    self.register_buffer('num_steps', torch.zeros(()))
    self.register_buffer('running_mean', torch.zeros(obs_size))
    self.register_buffer('running_var', torch.zeros(obs_size))

  @torch.no_grad()
  def update_normalization(self, obs):
    if not self.normalize_observations:
      return
    # running_statistics.update with pmap_axis_name (acme/running_statistics.py:174,262-271): the increment, the mean
    # update and the variance update are summed over ranks, so every rank holds the statistics of the global batch
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    inc = torch.tensor(float(obs.shape[0]), device=obs.device)
    if multi:
      dist.all_reduce(inc)
    total = self.num_steps + inc
    delta = obs - self.running_mean
    mean_update = delta.sum(0) / total
    if multi:
      dist.all_reduce(mean_update)
    self.running_mean += mean_update
    var_update = (delta * (obs - self.running_mean)).sum(0)
    if multi:
      dist.all_reduce(var_update)
    self.running_var += var_update
    self.num_steps.copy_(total)

  def normalize(self, obs):
    if not self.normalize_observations:
      return obs
    var = self.running_var / (self.num_steps + 1.0)
    return torch.clip((obs - self.running_mean) / (var.sqrt() + 1e-6), -5, 5)

  # NormalTanhDistribution (training/distribution.py:120-161): scale = softplus(.) + 0.001
  def dist_params(self, obs):
    loc, scale = torch.chunk(self.policy(self.normalize(obs)), 2, dim=-1)
    return loc, torch.nn.functional.softplus(scale) + 0.001

  @staticmethod
  def sample_pre_tanh(loc, scale):
    return loc + scale * torch.randn_like(loc)

  @staticmethod
  def log_prob(loc, scale, pre):
    lp = -0.5 * ((pre - loc) / scale) ** 2 - 0.5 * math.log(2 * math.pi) - torch.log(scale)
    lp = lp - 2 * (math.log(2) - pre - torch.nn.functional.softplus(-2 * pre))   # tanh bijector log-det
    return lp.sum(-1)

  def q_values(self, nets, obs, action):
    x = torch.cat([self.normalize(obs), action], -1)
    return torch.cat([q(x) for q in nets], -1)        # [B, n_critics]

  @torch.no_grad()
  def act(self, obs, deterministic: bool = False):
    loc, scale = self.dist_params(obs)
    return torch.tanh(loc if deterministic else self.sample_pre_tanh(loc, scale))


class ReplayBuffer:
  """Uniform sampling from a ring of transitions kept in HBM (reference replay_buffers.py:258-311)."""

  def __init__(self, capacity: int, obs_size: int, act_size: int, device):
    self.capacity, self.size, self.pos = int(capacity), 0, 0
    f = dict(dtype=torch.float32, device=device)
    self.obs = torch.empty((capacity, obs_size), **f); self.next_obs = torch.empty((capacity, obs_size), **f)
    self.action = torch.empty((capacity, act_size), **f)
    self.reward = torch.empty(capacity, **f); self.discount = torch.empty(capacity, **f); self.truncation = torch.empty(capacity, **f)
    self.size_t = torch.zeros((), **f)     # `size` on the device: the sampler inside a CUDA graph reads it

  def insert(self, obs, action, reward, discount, next_obs, truncation):
    n = obs.shape[0]
    idx = (torch.arange(n, device=obs.device) + self.pos) % self.capacity
    self.obs[idx], self.action[idx], self.reward[idx] = obs, action, reward
    self.discount[idx], self.next_obs[idx], self.truncation[idx] = discount, next_obs, truncation
    self.pos = (self.pos + n) % self.capacity
    self.size = min(self.size + n, self.capacity)
    self.size_t.fill_(float(self.size))

  def sample(self, batch_size: int):
    # floor(U[0,1) * size): graph-capturable (the bound is a device scalar, not a Python int)
    idx = (torch.rand(batch_size, device=self.obs.device) * self.size_t).long().clamp_(max=self.capacity - 1)
    return {'obs': self.obs[idx], 'action': self.action[idx], 'reward': self.reward[idx], 'discount': self.discount[idx],
            'next_obs': self.next_obs[idx], 'truncation': self.truncation[idx]}


def losses(net: SACNetworks, tr: Dict[str, torch.Tensor], reward_scaling: float, discounting: float, act_size: int):
  """(alpha_loss, critic_loss, actor_loss) as reference sac/losses.py:43-129."""
  target_entropy = -0.5 * act_size
  alpha = net.log_alpha.exp().detach()
  # alpha loss (:43-60)
  loc, scale = net.dist_params(tr['obs'])
  pre = net.sample_pre_tanh(loc, scale)
  log_prob = net.log_prob(loc, scale, pre)
  alpha_loss = (net.log_alpha.exp() * (-log_prob - target_entropy).detach()).mean()
  # critic loss (:62-104)
  q_old = net.q_values(net.q, tr['obs'], tr['action'])
  with torch.no_grad():
    nloc, nscale = net.dist_params(tr['next_obs'])
    npre = net.sample_pre_tanh(nloc, nscale)
    next_log_prob = net.log_prob(nloc, nscale, npre)
    next_q = net.q_values(net.target_q, tr['next_obs'], torch.tanh(npre))
    next_v = next_q.min(-1).values - alpha * next_log_prob
    target_q = tr['reward'] * reward_scaling + tr['discount'] * discounting * next_v
  q_error = (q_old - target_q[:, None]) * (1 - tr['truncation'])[:, None]
  critic_loss = 0.5 * (q_error ** 2).mean()
  # actor loss (:106-129); the q networks are held fixed for this term
  q_action = net.q_values(net.q, tr['obs'], torch.tanh(pre))
  actor_loss = (alpha * log_prob - q_action.min(-1).values).mean()
  return alpha_loss, critic_loss, actor_loss


def train(env_name: str = 'ant', num_timesteps: int = 1_000_000, episode_length: int = 1000, num_envs: int = 128,
          learning_rate: float = 1e-4, discounting: float = 0.9, seed: int = 0, batch_size: int = 256,
          normalize_observations: bool = False, reward_scaling: float = 1.0, tau: float = 0.005, min_replay_size: int = 0,
          max_replay_size: Optional[int] = None, grad_updates_per_step: int = 1, device=None, use_cuda_graph: bool = True,
          progress_fn: Optional[Callable[[int, Dict[str, float]], None]] = None, progress_every: int = 100):
  """Returns (networks, metrics).  metrics['sps'] = env-steps/sec including acting and learning."""
  world = dist.get_world_size() if dist.is_initialized() else 1
  rank = dist.get_rank() if dist.is_initialized() else 0
  device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
  if min_replay_size >= num_timesteps:
    raise ValueError('No training will happen because min_replay_size >= num_timesteps')
  env = envs.create(env_name, episode_length=episode_length, auto_reset=True, batch_size=num_envs, device=device,
                    env_id_offset=rank * num_envs)
  torch.manual_seed(seed + rank)
  net = SACNetworks(env.observation_size, env.action_size, normalize_observations=normalize_observations).to(device)
  if world > 1:
    for p in net.parameters():
      dist.broadcast(p.data, 0)
    net.target_q.load_state_dict(net.q.state_dict())
  graphed = use_cuda_graph and world == 1       # (several ranks: eager updates around the NCCL all-reduce)
  # learner GEMMs in TF32 on the tensor cores (XLA's default precision for the reference's float32 dots on GPU);
  # the physics step is plain FP32 and unaffected
  torch.backends.cuda.matmul.allow_tf32 = True
  policy_opt = torch.optim.Adam(net.policy.parameters(), lr=learning_rate, capturable=graphed, fused=True)
  q_opt = torch.optim.Adam(net.q.parameters(), lr=learning_rate, capturable=graphed, fused=True)
  alpha_opt = torch.optim.Adam([net.log_alpha], lr=3e-4, capturable=graphed, fused=True)        # reference train.py:233
  buf = ReplayBuffer((max_replay_size or num_timesteps) // world, env.observation_size, env.action_size, device)
  state = env.reset(seed)
  ep_reward = torch.zeros(num_envs, device=device)
  finished_sum = torch.zeros((), device=device); finished_n = torch.zeros((), device=device)

  def actor_step(st, random_action: bool):
    nonlocal ep_reward, finished_sum, finished_n
    action = torch.rand((num_envs, env.action_size), device=device) * 2 - 1 if random_action else net.act(st.obs)
    nst = env.step(st, action)
    net.update_normalization(st.obs)
    buf.insert(st.obs, action, nst.reward, 1 - nst.done, nst.obs, nst.info['truncation'])
    ep_reward = ep_reward + nst.reward
    finished_sum = finished_sum + (ep_reward * nst.done).sum(); finished_n = finished_n + nst.done.sum()
    ep_reward = ep_reward * (1 - nst.done)
    return nst

  def all_reduce_grads(params):
    if world > 1:
      flat = torch.cat([p.grad.reshape(-1) for p in params])
      dist.all_reduce(flat); flat /= world
      off = 0
      for p in params:
        p.grad.copy_(flat[off:off + p.numel()].view_as(p)); off += p.numel()

  total = 0
  while total < min_replay_size * world:      # prefill with the initial policy (reference train.py:430-452)
    state = actor_step(state, False)
    total += num_envs * world
  torch.cuda.synchronize()
  t0, it, metrics = time.perf_counter(), 0, {}
  q_params, pi_params = list(net.q.parameters()), list(net.policy.parameters())
  static_losses = torch.zeros(3, device=device)

  def update():
    """One gradient update (reference train.py:262-327): sample, three losses, three Adam steps, polyak."""
    tr = buf.sample(batch_size)
    a_loss, c_loss, p_loss = losses(net, tr, reward_scaling, discounting, env.action_size)
    alpha_opt.zero_grad(set_to_none=True); q_opt.zero_grad(set_to_none=True); policy_opt.zero_grad(set_to_none=True)
    a_loss.backward(inputs=[net.log_alpha])
    c_loss.backward(inputs=q_params)
    p_loss.backward(inputs=pi_params)
    all_reduce_grads([net.log_alpha]); all_reduce_grads(q_params); all_reduce_grads(pi_params)
    alpha_opt.step(); q_opt.step(); policy_opt.step()
    with torch.no_grad():      # polyak: target = target * (1 - tau) + q * tau (train.py:306-309)
      for tp, p in zip(net.target_q.parameters(), net.q.parameters()):
        tp.mul_(1 - tau).add_(p, alpha=tau)
      static_losses.copy_(torch.stack([a_loss.detach(), c_loss.detach(), p_loss.detach()]))

  update_graph = None
  if graphed:
    # the learner is launch-bound (three small MLPs, ~200 kernels per update): one CUDA graph per update
    if buf.size == 0:
      state = actor_step(state, False); total += num_envs
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
      for _ in range(3):      # warm-up outside capture (allocations, cuBLAS handles, Adam state)
        update()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    update_graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(update_graph):
      update()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
  while total < num_timesteps:
    state = actor_step(state, False)
    total += num_envs * world
    for _ in range(grad_updates_per_step):
      if update_graph is not None:
        update_graph.replay()
      else:
        update()
    it += 1
    if it % progress_every == 0 or total >= num_timesteps:
      torch.cuda.synchronize()
      a_l, c_l, p_l = static_losses.tolist()
      metrics = {'sps': (total - min_replay_size * world) / (time.perf_counter() - t0), 'alpha_loss': a_l, 'critic_loss': c_l, 'actor_loss': p_l,
                 'alpha': float(net.log_alpha.detach().exp()), 'episode_reward': float(finished_sum) / max(float(finished_n), 1.0),
                 'env_steps': total, 'iterations': it}
      if progress_fn and rank == 0:
        progress_fn(total, metrics)
      finished_sum.zero_(); finished_n.zero_()
  return net, metrics
