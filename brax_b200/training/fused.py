"""The two hand-written kernels of the PPO loop (brax_b200/csrc/bxg_train.cu, C ABI in include/bxg.h).

`gae`          compute_gae (agents/ppo/losses.py:38-101): one launch instead of ~40 small ones per minibatch.
`policy_act`   normalise -> policy MLP [obs, 64, 64, 2 act] -> tanh-normal sample: one launch per env step of the
               rollout (training/acting.py:33-53) instead of ~15.
Both take and return torch CUDA tensors and launch on the current stream (CUDA-graph capturable)."""
from __future__ import annotations

import torch

from brax_b200 import native


def _ptr(t):
  assert t.is_cuda and t.is_contiguous() and t.dtype == torch.float32, (t.device, t.dtype, t.is_contiguous())
  return t.data_ptr()


def gae(truncation, termination, reward, values, bootstrap, lambda_: float, discount: float):
  """Time-major [T, B] float32 CUDA tensors, bootstrap [B] -> (vs, advantages), both [T, B]."""
  T, B = reward.shape
  vs, adv = torch.empty_like(reward), torch.empty_like(reward)
  args = [t.contiguous() for t in (truncation, termination, reward, values, bootstrap)]
  stream = torch.cuda.current_stream(reward.device).cuda_stream
  with torch.cuda.device(reward.device):
    native._check(native.lib().bxg_gae(*[_ptr(t) for t in args], T, B, float(lambda_), float(discount), _ptr(vs), _ptr(adv), stream), 'bxg_gae')
  return vs, adv


def supports(policy: torch.nn.Sequential) -> bool:
  """The kernel is compiled for the [obs, 64, 64, 2 act] swish policy (notebooks/training_torch.ipynb)."""
  lin = [m for m in policy if isinstance(m, torch.nn.Linear)]
  act = [m for m in policy if isinstance(m, torch.nn.SiLU)]
  return (len(lin) == 3 and len(act) == 2 and lin[0].out_features == 64 and lin[1].out_features == 64
          and lin[0].weight.is_cuda and lin[0].weight.dtype == torch.float32)


def policy_act(policy: torch.nn.Sequential, mean, std, obs, noise, clip=None, min_std: float = 0.001):
  """obs [n, obs], noise [n, act] standard normal -> (action = tanh(raw), logits [n, 2 act], raw action)."""
  l1, l2, l3 = [m for m in policy if isinstance(m, torch.nn.Linear)]
  n, no = obs.shape
  na = l3.out_features // 2
  logits = torch.empty((n, 2 * na), dtype=torch.float32, device=obs.device)
  pre = torch.empty((n, na), dtype=torch.float32, device=obs.device)
  action = torch.empty_like(pre)
  stream = torch.cuda.current_stream(obs.device).cuda_stream
  with torch.cuda.device(obs.device):
    native._check(native.lib().bxg_policy_act(
        _ptr(obs.contiguous()), _ptr(mean), _ptr(std), float(clip) if clip else 0.0, _ptr(l1.weight), _ptr(l1.bias), _ptr(l2.weight),
        _ptr(l2.bias), _ptr(l3.weight), _ptr(l3.bias), _ptr(noise.contiguous()), n, no, l1.out_features, l2.out_features, na,
        float(min_std), _ptr(logits), _ptr(pre), _ptr(action), stream), 'bxg_policy_act')
  return action, logits, pre


class _PPOHead(torch.autograd.Function):
  """compute_ppo_loss from the network outputs on, forward and backward in two hand-written launches
  (bxg_ppo_head).  The gradient wrt the logits and the values is produced in the forward pass."""

  @staticmethod
  def forward(ctx, logits, values, beh_logits, pre, reward, done, trunc, noise, cfg):
    T, B, A2 = logits.shape
    dev = logits.device
    args = [t.contiguous() for t in (logits, values, beh_logits, pre, reward, done, trunc, noise)]
    vs, adv = torch.empty((T, B), device=dev), torch.empty((T, B), device=dev)
    stats = torch.empty(2, dtype=torch.float64, device=dev)
    loss4 = torch.empty(4, device=dev)
    dlogits, dvalues = torch.empty_like(args[0]), torch.empty_like(args[1])
    stream = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
      native._check(native.lib().bxg_ppo_head(
          *[_ptr(t) for t in args], T, B, A2 // 2, float(cfg['reward_scaling']), float(cfg['lambda_']), float(cfg['discounting']),
          float(cfg['epsilon']), float(cfg['entropy_cost']), int(bool(cfg['normalize_advantage'])), float(cfg.get('min_std', 0.001)),
          _ptr(vs), _ptr(adv), stats.data_ptr(), _ptr(loss4), _ptr(dlogits), _ptr(dvalues), stream), 'bxg_ppo_head')
    ctx.save_for_backward(dlogits, dvalues)
    ctx.mark_non_differentiable()
    return loss4

  @staticmethod
  def backward(ctx, g):
    dlogits, dvalues = ctx.saved_tensors
    # loss4 = (total, policy, value, entropy): the gradients stored are those of the total
    return g[0] * dlogits, g[0] * dvalues, None, None, None, None, None, None, None


def ppo_head(logits, values, beh_logits, pre, reward, done, trunc, noise, **cfg):
  """-> tensor [total, policy_loss, v_loss, entropy_loss]; differentiate `[0]` (the total)."""
  return _PPOHead.apply(logits, values, beh_logits, pre, reward, done, trunc, noise, cfg)
