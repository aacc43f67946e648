"""Golden vectors for the PPO arithmetic from the REFERENCE'S OWN SOURCE (build container only).

Imports, unmodified from /root/reference and run on NumPy float64 through tools/refshim/:
  brax.training.agents.ppo.losses   compute_gae, compute_ppo_loss        (losses.py:38-101,143-303)
  brax.training.distribution        NormalTanhDistribution (log_prob, entropy, create_dist)
  brax.training.acme.running_statistics  init_state, update, normalize    (running_statistics.py:58-328)
The policy / value networks are plain MLPs (swish) evaluated here in NumPy from weights stored in
the golden file: compute_ppo_loss only calls `.apply`.  jax.random.normal is replaced by draws that
are stored too, so the sampled entropy term can be reproduced.

  python tools/gen_ppo_golden.py     ->  tests/golden/ref_ppo.npz
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tools', 'refshim'))
sys.path.insert(0, '/root/reference')

import jax                      # noqa: E402  (the stand-in)
import jax.numpy as jnp         # noqa: E402

# ppo/networks.py builds flax.linen modules: not on the arithmetic path (the loss only calls .apply)
_nets = types.ModuleType('brax.training.agents.ppo.networks')
_nets.PPONetworks = type('PPONetworks', (), {})
sys.modules['brax.training.agents.ppo.networks'] = _nets

from brax.training import distribution, types as btypes          # noqa: E402  (the reference)
from brax.training.acme import running_statistics, specs          # noqa: E402  (the reference)
from brax.training.agents.ppo import losses                       # noqa: E402  (the reference)

assert losses.__file__.startswith('/root/reference/'), losses.__file__

T, B, OBS, ACT, H = 5, 12, 27, 8, 32


def mlp(params, x):
  n = len(params) // 2
  for i in range(n):
    x = x @ params[2 * i] + params[2 * i + 1]
    if i < n - 1:
      x = x * (1.0 / (1.0 + np.exp(-x)))     # linen.swish
  return x


def main():
  rng = np.random.default_rng(5)
  out = {}
  # ---- running statistics: three updates, then normalize ------------------------------------
  rs = running_statistics.init_state(specs.Array((OBS,), jnp.dtype('float32')))
  batches = [rng.normal(1.0 + k, 2.0, (T, B, OBS)) for k in range(3)]
  for k, b in enumerate(batches):
    rs = running_statistics.update(rs, jnp.array(b))
    out[f'rs_batch{k}'] = b
    out[f'rs_mean{k}'] = np.asarray(rs.mean); out[f'rs_std{k}'] = np.asarray(rs.std)
    out[f'rs_summed_variance{k}'] = np.asarray(rs.summed_variance)
    cnt = rs.count
    out[f'rs_count{k}'] = np.asarray(float(cnt.hi) * 2.0 ** 32 + float(cnt.lo) if hasattr(cnt, 'hi') else float(cnt))
  probe = rng.normal(0.0, 5.0, (7, OBS))
  out['rs_probe'] = probe
  out['rs_probe_normalized'] = np.asarray(running_statistics.normalize(jnp.array(probe), rs))
  out['rs_probe_normalized_clip5'] = np.asarray(running_statistics.normalize(jnp.array(probe), rs, max_abs_value=5.0))

  # ---- compute_gae ----------------------------------------------------------------------------
  trunc = (rng.uniform(size=(T, B)) < 0.15).astype(np.float64)
  done = np.maximum(trunc, (rng.uniform(size=(T, B)) < 0.2).astype(np.float64))
  term = done * (1 - trunc)
  rew = rng.normal(size=(T, B)); val = rng.normal(size=(T, B)); boot = rng.normal(size=(B,))
  vs, adv = losses.compute_gae(truncation=jnp.array(trunc), termination=jnp.array(term), rewards=jnp.array(rew),
                               values=jnp.array(val), bootstrap_value=jnp.array(boot), lambda_=0.95, discount=0.97)
  out.update(gae_truncation=trunc, gae_termination=term, gae_rewards=rew, gae_values=val, gae_bootstrap=boot,
             gae_vs=np.asarray(vs), gae_advantages=np.asarray(adv))

  # ---- compute_ppo_loss -----------------------------------------------------------------------
  def init(sizes):
    p = []
    for i in range(len(sizes) - 1):
      p += [rng.normal(0, 1.0 / np.sqrt(sizes[i]), (sizes[i], sizes[i + 1])), rng.normal(0, 0.1, (sizes[i + 1],))]
    return p
  pol, vf = init([OBS, H, H, 2 * ACT]), init([OBS, H, H, 1])
  for i, w in enumerate(pol):
    out[f'policy_{i}'] = w
  for i, w in enumerate(vf):
    out[f'value_{i}'] = w
  dist = distribution.NormalTanhDistribution(event_size=ACT)
  net = _nets.PPONetworks()
  net.parametric_action_distribution = dist
  net.policy_network = types.SimpleNamespace(apply=lambda norm, p, obs: jnp.array(mlp(p, np.asarray(running_statistics.normalize(obs, norm)))))
  net.value_network = types.SimpleNamespace(apply=lambda norm, p, obs: jnp.array(mlp(p, np.asarray(running_statistics.normalize(obs, norm)))[..., 0]))
  obs = rng.normal(1.0, 2.0, (B, T + 1, OBS))           # [B, T(+1), ...]: the loss swaps to time-major itself
  # behaviour policy: slightly different weights
  pol_b = [w + 0.05 * rng.normal(size=w.shape) for w in pol]
  logits_b = mlp(pol_b, np.asarray(running_statistics.normalize(jnp.array(obs[:, :-1]), rs)))
  loc_b, scale_b = logits_b[..., :ACT], np.log1p(np.exp(logits_b[..., ACT:])) + 0.001
  raw = loc_b + scale_b * rng.normal(size=loc_b.shape)
  logp_b = np.asarray(dist.log_prob(jnp.array(logits_b), jnp.array(raw)))
  reward = rng.normal(size=(B, T))
  trunc2 = (rng.uniform(size=(B, T)) < 0.15).astype(np.float64)
  done2 = np.maximum(trunc2, (rng.uniform(size=(B, T)) < 0.2).astype(np.float64))
  data = btypes.Transition(
      observation=jnp.array(obs[:, :-1]), action=jnp.array(np.tanh(raw)), reward=jnp.array(reward), discount=jnp.array(1 - done2),
      next_observation=jnp.array(obs[:, 1:]),
      extras={'state_extras': {'truncation': jnp.array(trunc2)},
              'policy_extras': {'raw_action': jnp.array(raw), 'log_prob': jnp.array(logp_b), 'distribution_params': jnp.array(logits_b)}})
  noise = rng.normal(size=(T, B, ACT))
  jax.random.normal = lambda key, shape=(), dtype=None: jnp.array(noise.reshape(shape))   # entropy's sample (distribution.py)
  out.update(loss_obs=obs, loss_raw_action=raw, loss_behaviour_logits=logits_b, loss_behaviour_log_prob=logp_b, loss_reward=reward,
             loss_truncation=trunc2, loss_done=done2, loss_entropy_noise=noise)
  for tag, kw in (('adv_norm', dict(normalize_advantage=True)), ('no_adv_norm', dict(normalize_advantage=False))):
    total, m = losses.compute_ppo_loss(losses.PPONetworkParams(policy=pol, value=vf), rs, data, jnp.array([0, 1]), net,
                                       entropy_cost=1e-2, discounting=0.97, reward_scaling=10.0, gae_lambda=0.95,
                                       clipping_epsilon=0.3, **kw)
    out[f'loss_{tag}_total'] = np.asarray(total)
    for k in ('policy_loss', 'v_loss', 'entropy_loss'):
      out[f'loss_{tag}_{k}'] = np.asarray(m[k])
    print(tag, float(total), {k: float(m[k]) for k in ('policy_loss', 'v_loss', 'entropy_loss')})
  path = os.path.join(ROOT, 'tests', 'golden', 'ref_ppo.npz')
  np.savez_compressed(path, **{k: np.asarray(v, np.float64) for k, v in out.items()})
  print('wrote', path, os.path.getsize(path) // 1024, 'KB')


if __name__ == '__main__':
  main()
