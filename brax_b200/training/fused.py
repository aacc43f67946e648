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
