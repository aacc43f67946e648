"""PPO on the fused B200 env step (BASELINE config 5).

The physics + env arithmetic is the hand-written kernel (one launch per env
step for the whole batch); policy / value networks, GAE and Adam are ordinary
PyTorch, as in the PyTorch agent the reference ships
(`notebooks/training_torch.ipynb:94-480`): tanh-squashed diagonal Gaussian policy
`[obs, 64, 64, 2*act]`, value `[obs, 64, 64, 1]`, running observation
normalisation, clipped surrogate, entropy bonus.  Hyper-parameter names follow
`brax/training/agents/ppo/train.py:176-230`.

Multi-GPU: one process per GPU, env batch sharded by global env id (no
collective in the rollout), gradients averaged with an NCCL all-reduce per
minibatch -- the `lax.pmean` of `training/gradients.py:32`.
"""
from __future__ import annotations

import math
import time
from typing import Callable, Dict, Optional

import numpy as np
import torch
import torch.distributed as dist
from torch import nn

from brax_b200 import envs
from brax_b200.training import fused


class Agent(nn.Module):
  def __init__(self, obs_size: int, act_size: int, hidden=(64, 64), entropy_cost=1e-2, discounting=0.97,
               reward_scaling=10.0, lambda_=0.95, epsilon=0.3, normalize_advantage=True, clip_obs=None):
    super().__init__()
    def mlp(sizes):
      layers = []
      for i in range(len(sizes) - 1):
        layers += [nn.Linear(sizes[i], sizes[i + 1]), nn.SiLU()]
      return nn.Sequential(*layers[:-1])
    self.policy = mlp([obs_size, *hidden, 2 * act_size])
    self.value = mlp([obs_size, *hidden, 1])
    # running_statistics.RunningStatisticsState (acme/running_statistics.py:62-68): count, mean, summed_variance, std
    self.register_buffer('num_steps', torch.zeros(()))
    self.register_buffer('running_mean', torch.zeros(obs_size))
    self.register_buffer('running_var', torch.zeros(obs_size))     # summed_variance
    self.register_buffer('running_std', torch.ones(obs_size))
    self.entropy_cost, self.discounting, self.reward_scaling = entropy_cost, discounting, reward_scaling
    self.lambda_, self.epsilon = lambda_, epsilon
    self.normalize_advantage = normalize_advantage
    self.clip_obs = clip_obs
    self.fused_head = True      # CUDA float32: the loss head runs as bxg_ppo_head (the framework-op statement below otherwise)

  @torch.no_grad()
  def update_normalization(self, obs, group=None):
    """running_statistics.update (acme/running_statistics.py:120-300, Welford branch) with
    `pmap_axis_name` = the process group: the step increment, the mean update and the variance
    update are summed over ranks (`jax.lax.psum`, :174, :262-264, :270-271), so every rank ends
    with identical statistics.  `group=None` uses the default group when torch.distributed is
    initialised with more than one rank."""
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    flat = obs.reshape(-1, obs.shape[-1])
    inc = torch.tensor(float(flat.shape[0]), device=flat.device)
    if multi:
      dist.all_reduce(inc, group=group)
    count = self.num_steps + inc
    diff_to_old_mean = flat - self.running_mean
    mean_update = diff_to_old_mean.sum(0) / count
    if multi:
      dist.all_reduce(mean_update, group=group)
    mean = self.running_mean + mean_update
    variance_update = (diff_to_old_mean * (flat - mean)).sum(0)
    if multi:
      dist.all_reduce(variance_update, group=group)
    self.running_mean.copy_(mean)
    self.running_var.add_(variance_update)
    self.num_steps.copy_(count)
    self.running_std.copy_(torch.clip(torch.sqrt(torch.clamp(self.running_var, min=0) / count), 1e-6, 1e6))

  def normalize(self, obs):
    """running_statistics.normalize (acme/running_statistics.py:303-328)."""
    out = (obs - self.running_mean) / self.running_std
    return out if self.clip_obs is None else torch.clip(out, -self.clip_obs, self.clip_obs)

  @staticmethod
  def dist_create(logits):
    loc, scale = torch.chunk(logits, 2, dim=-1)
    return loc, torch.nn.functional.softplus(scale) + 0.001

  @staticmethod
  def dist_log_prob(loc, scale, pre_tanh):
    lp = -0.5 * ((pre_tanh - loc) / scale) ** 2 - 0.5 * math.log(2 * math.pi) - torch.log(scale)
    lp = lp - 2 * (math.log(2) - pre_tanh - torch.nn.functional.softplus(-2 * pre_tanh))
    return lp.sum(-1)

  @staticmethod
  def dist_entropy(loc, scale, noise=None):
    """ParametricDistribution.entropy (distribution.py:84-91): Gaussian entropy plus the tanh
    log-det-jacobian at one sample; `noise` stands for the sample's standard normal draw."""
    ent = 0.5 + 0.5 * math.log(2 * math.pi) + torch.log(scale)
    sample = loc + scale * (torch.randn_like(loc) if noise is None else noise)
    ent = ent + 2 * (math.log(2) - sample - torch.nn.functional.softplus(-2 * sample))
    return ent.sum(-1)

  @torch.no_grad()
  def act(self, obs):
    if obs.is_cuda and obs.dtype == torch.float32 and fused.supports(self.policy):
      # one hand-written launch: normalise, MLP, sample, tanh (brax_b200/csrc/bxg_train.cu)
      noise = torch.randn((obs.shape[0], self.policy[-1].out_features // 2), device=obs.device)
      return fused.policy_act(self.policy, self.running_mean, self.running_std, obs, noise, clip=self.clip_obs)
    logits = self.policy(self.normalize(obs))
    loc, scale = self.dist_create(logits)
    pre = loc + scale * torch.randn_like(loc)
    return torch.tanh(pre), logits, pre

  @torch.no_grad()
  def gae(self, truncation, termination, reward, values, bootstrap):
    if reward.is_cuda and reward.dtype == torch.float32:
      return fused.gae(truncation, termination, reward, values, bootstrap, self.lambda_, self.discounting)   # one launch
    return self.gae_reference(truncation, termination, reward, values, bootstrap)

  @torch.no_grad()
  def gae_reference(self, truncation, termination, reward, values, bootstrap):
    """compute_gae (agents/ppo/losses.py:38-101) in framework ops: the statement the kernel is tested against."""
    mask = 1.0 - truncation
    values_t1 = torch.cat([values[1:], bootstrap[None]], 0)
    deltas = (reward + self.discounting * (1 - termination) * values_t1 - values) * mask
    acc = torch.zeros_like(bootstrap)
    out = []
    for t in range(deltas.shape[0] - 1, -1, -1):
      acc = deltas[t] + self.discounting * (1 - termination[t]) * mask[t] * self.lambda_ * acc
      out.append(acc)
    vs = torch.stack(out[::-1]) + values
    vs_t1 = torch.cat([vs[1:], bootstrap[None]], 0)
    adv = (reward + self.discounting * (1 - termination) * vs_t1 - values) * mask
    return vs, adv

  def loss(self, td: Dict[str, torch.Tensor], entropy_noise=None, parts: bool = False):
    """compute_ppo_loss (agents/ppo/losses.py:143-303) on time-major data."""
    obs = self.normalize(td['obs'])                       # [T+1, B, obs]
    values = self.value(obs).squeeze(-1)
    logits = self.policy(obs[:-1])
    if self.fused_head and logits.is_cuda and logits.dtype == torch.float32:
      # everything from the network outputs on, forward and backward, in two hand-written launches (bxg_ppo_head)
      noise = torch.randn_like(td['pre']) if entropy_noise is None else entropy_noise
      l4 = fused.ppo_head(logits, values, td['logits'], td['pre'], td['reward'], td['done'], td['truncation'], noise,
                          reward_scaling=self.reward_scaling, lambda_=self.lambda_, discounting=self.discounting, epsilon=self.epsilon,
                          entropy_cost=self.entropy_cost, normalize_advantage=self.normalize_advantage)
      return (l4[0], l4[1], l4[2], l4[3]) if parts else l4[0]
    loc, scale = self.dist_create(logits)
    beh_loc, beh_scale = self.dist_create(td['logits'])
    lp = self.dist_log_prob(loc, scale, td['pre'])
    blp = self.dist_log_prob(beh_loc, beh_scale, td['pre'])
    reward = td['reward'] * self.reward_scaling
    termination = td['done'] * (1 - td['truncation'])
    vs, adv = self.gae(td['truncation'], termination, reward, values[:-1].detach(), values[-1].detach())
    if self.normalize_advantage:                           # losses.py:236-237 (jnp.std: population standard deviation)
      adv = (adv - adv.mean()) / (adv.std(unbiased=False) + 1e-8)
    rho = torch.exp(lp - blp)
    policy_loss = -torch.minimum(rho * adv, rho.clip(1 - self.epsilon, 1 + self.epsilon) * adv).mean()
    v_loss = 0.25 * ((vs - values[:-1]) ** 2).mean()       # mean(v_error^2) * 0.5 * vf_coefficient(0.5), losses.py:258-270
    ent_loss = -self.entropy_cost * self.dist_entropy(loc, scale, entropy_noise).mean()
    total = policy_loss + v_loss + ent_loss
    return (total, policy_loss, v_loss, ent_loss) if parts else total


def train(env_name: str = 'ant', num_envs: int = 2048, episode_length: int = 1000, num_timesteps: int = 1_000_000,
          unroll_length: int = 5, batch_size: int = 1024, num_minibatches: int = 32, num_update_epochs: int = 4,
          reward_scaling: float = 10.0, entropy_cost: float = 1e-2, discounting: float = 0.97, learning_rate: float = 3e-4,
          normalize_advantage: bool = True, randomization_fn: Optional[Callable] = None,
          seed: int = 0, device=None, use_cuda_graph: bool = True, progress_fn: Optional[Callable[[int, Dict[str, float]], None]] = None):
  """Returns (agent, metrics).  metrics['sps'] = env-steps/sec including policy
  inference and learning (the figure BASELINE config 5 asks for).

  randomization_fn: `fn(sys, rng) -> (sys_v, in_axes)` as in the reference trainer (agents/ppo/train.py:88-90, 268-276):
  `rng` holds one integer seed per env of this rank (the reference passes one PRNG key per env), `sys_v` the System
  with a leading env axis on the randomised leaves, `in_axes` 0 at those leaves and None elsewhere."""
  world = dist.get_world_size() if dist.is_initialized() else 1
  rank = dist.get_rank() if dist.is_initialized() else 0
  device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
  # lean pipeline state: the rollout never reads the derived State leaves, the step recomputes them on chip
  env = envs.create(env_name, episode_length=episode_length, auto_reset=True, batch_size=num_envs, device=device,
                    env_id_offset=rank * num_envs, lean=True)
  if randomization_fn is not None:
    import functools
    from brax_b200.envs.wrappers import training as wrappers
    env_seeds = np.random.SeedSequence([seed, rank]).generate_state(num_envs)
    env = wrappers.DomainRandomizationVmapWrapper(env, functools.partial(randomization_fn, rng=env_seeds))
    assert env.batch_size == num_envs, 'randomization_fn must return one System per env'
  torch.manual_seed(seed + rank)
  agent = Agent(env.observation_size, env.action_size, entropy_cost=entropy_cost, discounting=discounting,
                reward_scaling=reward_scaling, normalize_advantage=normalize_advantage).to(device)
  if world > 1:
    for p in agent.parameters():
      dist.broadcast(p.data, 0)
  # the learner's GEMMs run on the tensor cores in TF32, as XLA's default precision does for the reference's
  # float32 dots on GPU; the physics step is untouched by this switch (plain FP32 FMA)
  torch.backends.cuda.matmul.allow_tf32 = True
  torch.backends.cudnn.allow_tf32 = True
  opt = torch.optim.Adam(agent.parameters(), lr=learning_rate, capturable=use_cuda_graph, fused=True)
  sgd_graph = step_graph = static_mb = static_loss = static_flat = None
  state = env.reset(seed)
  # one training step consumes batch_size * num_minibatches trajectories of unroll_length steps
  traj_per_step = batch_size * num_minibatches
  rollouts_per_step = max(1, traj_per_step // num_envs)
  steps_per_train_step = rollouts_per_step * num_envs * unroll_length * world
  total, it = 0, 0
  ep_reward = torch.zeros(num_envs, device=device)
  finished_sum = torch.zeros((), device=device); finished_n = torch.zeros((), device=device)   # no host sync in the rollout
  metrics: Dict[str, float] = {}
  # ---- rollout of `unroll_length` env steps; optionally one CUDA graph ----------------------------
  from brax_b200.base import tree_map

  def unroll(st, ep_rew, fsum, fnum):
    o, lg, pr, rw, dn, tr = [st.obs], [], [], [], [], []
    for _ in range(unroll_length):
      action, logits, pre = agent.act(st.obs)
      st = env.step(st, action)
      o.append(st.obs); lg.append(logits); pr.append(pre)
      rw.append(st.reward); dn.append(st.done); tr.append(st.info['truncation'])
      ep_rew = ep_rew + st.reward
      fsum = fsum + (ep_rew * st.done).sum(); fnum = fnum + st.done.sum()
      ep_rew = ep_rew * (1 - st.done)
    return st, ep_rew, fsum, fnum, (torch.stack(o), torch.stack(lg), torch.stack(pr), torch.stack(rw), torch.stack(dn), torch.stack(tr))

  rollout_graph = None
  if use_cuda_graph:   # (no collective inside the rollout: graphs work the same on every rank)
    # static state buffers: the graph reads them, steps, and writes the final state back in place
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
      for _ in range(2):   # warm-up (model upload, allocator, cuBLAS)
        unroll(state, ep_reward, finished_sum, finished_n)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    static_state = state
    rollout_graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(rollout_graph):
      st2, er2, fs2, fn2, static_traj = unroll(static_state, ep_reward, finished_sum, finished_n)
      # write the carried values back into the static inputs
      def write_back(dst, src):
        if isinstance(dst, torch.Tensor) and dst.data_ptr() != src.data_ptr():
          dst.copy_(src)
        return dst
      tree_map(write_back, static_state.pipeline_state, st2.pipeline_state)
      static_state.obs.copy_(st2.obs); static_state.reward.copy_(st2.reward); static_state.done.copy_(st2.done)
      for k in ('steps', 'truncation'):
        static_state.info[k].copy_(st2.info[k])
      ep_reward.copy_(er2); finished_sum.copy_(fs2); finished_n.copy_(fn2)

  torch.cuda.synchronize()
  t0 = time.perf_counter()
  t_steady, total_steady = None, 0     # from the end of the first iteration on (it captures the learner's CUDA graphs)
  while total < num_timesteps:
    obs_l, logit_l, pre_l, rew_l, done_l, trunc_l = [], [], [], [], [], []
    for _ in range(rollouts_per_step):
      if rollout_graph is not None:
        rollout_graph.replay()
        traj = [t.clone() for t in static_traj]
      else:
        state, ep_reward, finished_sum, finished_n, traj = unroll(state, ep_reward, finished_sum, finished_n)
      obs_l.append(traj[0]); logit_l.append(traj[1]); pre_l.append(traj[2])
      rew_l.append(traj[3]); done_l.append(traj[4]); trunc_l.append(traj[5])
    td = {'obs': torch.cat(obs_l, 1), 'logits': torch.cat(logit_l, 1), 'pre': torch.cat(pre_l, 1),
          'reward': torch.cat(rew_l, 1), 'done': torch.cat(done_l, 1), 'truncation': torch.cat(trunc_l, 1)}
    agent.update_normalization(td['obs'][:-1])
    n_traj = td['reward'].shape[1]
    mb_size = n_traj // num_minibatches
    if use_cuda_graph and sgd_graph is None:
      # capture one minibatch update (loss, backward, Adam) into a CUDA graph: the
      # learner is launch-bound (small MLPs), replay removes the per-op overhead.  With several
      # ranks the update is two graphs around one NCCL all-reduce of the flattened gradient
      # (lax.pmean(grad) of the reference): [loss, backward, flatten] -> all_reduce -> [unflatten, Adam].
      static_mb = {k: torch.empty_like(v[:, :mb_size]) for k, v in td.items()}
      static_loss = torch.zeros((), device=device)
      params = list(agent.parameters())
      static_flat = torch.zeros(sum(p.numel() for p in params), device=device)

      def scatter_grads():
        off = 0
        for p in params:
          p.grad.copy_(static_flat[off:off + p.numel()].view_as(p)); off += p.numel()

      side = torch.cuda.Stream()
      side.wait_stream(torch.cuda.current_stream())
      with torch.cuda.stream(side):
        for _ in range(3):   # warm-up outside capture (allocations, cuBLAS handles)
          for k, v in td.items():
            static_mb[k].copy_(v[:, :mb_size])
          opt.zero_grad(set_to_none=False)
          agent.loss(static_mb).backward()
          if world > 1:
            torch.cat([p.grad.reshape(-1) for p in params], out=static_flat)
            dist.all_reduce(static_flat); static_flat /= world
            scatter_grads()
          opt.step()
      torch.cuda.current_stream().wait_stream(side)
      torch.cuda.synchronize()
      sgd_graph = torch.cuda.CUDAGraph()
      opt.zero_grad(set_to_none=world == 1)     # one rank: gradients are (re)created inside the graph
      with torch.cuda.graph(sgd_graph):
        if world > 1:                            # several ranks: persistent gradients, zeroed in the graph
          for p in params:
            p.grad.zero_()
        l_ = agent.loss(static_mb)
        l_.backward()
        static_loss.copy_(l_.detach())
        if world > 1:
          torch.cat([p.grad.reshape(-1) for p in params], out=static_flat)
        else:
          opt.step()
      step_graph = None
      if world > 1:
        step_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(step_graph):
          static_flat.div_(world)
          scatter_grads()
          opt.step()
    for _ in range(num_update_epochs):
      perm = torch.randperm(n_traj, device=device)
      if sgd_graph is not None:   # one gather per leaf and epoch; the minibatches are then plain slices
        shuffled = {k: torch.index_select(v, 1, perm[:mb_size * num_minibatches]) for k, v in td.items()}
      for i, mb in enumerate(perm[:mb_size * num_minibatches].view(num_minibatches, mb_size)):
        if sgd_graph is not None:
          for k, v in shuffled.items():
            static_mb[k].copy_(v[:, i * mb_size:(i + 1) * mb_size])
          sgd_graph.replay()
          if world > 1:
            dist.all_reduce(static_flat)
            step_graph.replay()
          loss = static_loss
          continue
        loss = agent.loss({k: v[:, mb] for k, v in td.items()})
        opt.zero_grad(set_to_none=True)
        loss.backward()
        if world > 1:   # lax.pmean(grad) of the reference, over NCCL
          flat = torch.cat([p.grad.reshape(-1) for p in agent.parameters()])
          dist.all_reduce(flat); flat /= world
          off = 0
          for p in agent.parameters():
            p.grad.copy_(flat[off:off + p.numel()].view_as(p)); off += p.numel()
        opt.step()
    total += steps_per_train_step
    it += 1
    torch.cuda.synchronize()
    now = time.perf_counter()
    if t_steady is None:
      t_steady, total_steady = now, total
    steady = (total - total_steady) / (now - t_steady) if total > total_steady else total / (now - t0)
    metrics = {'sps': total / (now - t0), 'sps_steady': steady, 'loss': float(loss.detach()),
               'episode_reward': float(finished_sum) / max(float(finished_n), 1.0), 'env_steps': total, 'iterations': it}
    if progress_fn and rank == 0:
      progress_fn(total, metrics)
    finished_sum.zero_(); finished_n.zero_()
  return agent, metrics
