"""Domain-randomisation golden from the REFERENCE'S OWN SOURCE (build container only).

Runs `brax.envs.wrappers.training.wrap(env, randomization_fn=...)` -- i.e. the reference's
DomainRandomizationVmapWrapper (wrappers/training.py:223-260) under EpisodeWrapper and AutoResetWrapper -- unmodified
from /root/reference on the NumPy stand-ins of tools/refshim/ (float64; `jax.vmap` with the System-shaped `in_axes`
is a Python loop over the envs).  The randomisation is the reference test's kind (`ppo/train_test.py:229-239`:
`sys.tree_replace` of batched leaves + an `in_axes` tree of None / 0): per env, the links' centre-of-mass offsets,
masses and inertias, the actuator gears, the joint damping and the armature.

Writes tests/golden/ref_dr_<env>.npz: the randomised leaves (`rand_<path>`, float32 values), q0, qd0, obs0, the
actions and, per env-step, obs / reward / done / steps / truncation / metrics and every pipeline-state leaf.
tests/test_domain_randomization.py and tests/test_gpu_domain_randomization.py replay it.

  python tools/gen_reference_golden_dr.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import gen_reference_golden as G   # noqa: E402  (sets up the stand-ins and imports the reference)

import jax                         # noqa: E402  (the stand-in)
import jax.numpy as jp             # noqa: E402

f32 = lambda a: np.asarray(a, np.float64).astype(np.float32).astype(np.float64)   # noqa: E731


def randomised_leaves(sys_, n, seed):
  """{path: [n, ...] array} -- float32-representable values, so the float32 System of the product holds the same numbers."""
  rng = np.random.default_rng(seed)
  L, nv, nu = sys_.link.inertia.mass.shape[0], sys_.dof.damping.shape[0], sys_.actuator.gear.shape[0]
  scale = rng.uniform(0.7, 1.4, (n, L))
  return {
      'link.inertia.transform.pos': f32(np.asarray(sys_.link.inertia.transform.pos)[None] + rng.uniform(-0.02, 0.02, (n, L, 3))),
      'link.inertia.mass': f32(np.asarray(sys_.link.inertia.mass)[None] * scale),
      'link.inertia.i': f32(np.asarray(sys_.link.inertia.i)[None] * scale[:, :, None, None]),
      'actuator.gear': f32(np.asarray(sys_.actuator.gear)[None] * rng.uniform(0.7, 1.3, (n, nu))),
      'dof.damping': f32(np.asarray(sys_.dof.damping)[None] * rng.uniform(0.5, 2.0, (n, nv))),
      'dof.armature': f32(np.asarray(sys_.dof.armature)[None] * rng.uniform(0.8, 1.25, (n, nv))),
  }


def main():
  G.mjcf_stub.load = G._mjcf_load
  from brax.envs import ant as ref_ant, humanoid as ref_hum     # the reference
  from brax.envs.wrappers import training as ref_wrap           # the reference
  assert ref_wrap.__file__.startswith('/root/reference/')
  for name, cls, n, steps, ep_len in (('ant', ref_ant.Ant, 6, 8, 5), ('humanoid', ref_hum.Humanoid, 4, 6, 4)):
    env0 = cls(backend='generalized')
    leaves = randomised_leaves(env0.sys, n, seed=11)

    def rand_fn(sys_):
      sys_v = sys_.tree_replace({k: jp.array(v) for k, v in leaves.items()})
      in_axes = jax.tree.map(lambda x: None, sys_)
      in_axes = in_axes.tree_replace({k: 0 for k in leaves})
      return sys_v, in_axes

    env = ref_wrap.wrap(env0, episode_length=ep_len, action_repeat=1, randomization_fn=rand_fn)
    assert type(env.env.env).__name__ == 'DomainRandomizationVmapWrapper', type(env.env.env)
    rng = np.random.default_rng(5)
    st = env.reset(jax.random.split(jax.random.PRNGKey(9), n))
    out = {f'rand_{k}': v for k, v in leaves.items()}
    out.update({'q0': np.asarray(st.pipeline_state.q), 'qd0': np.asarray(st.pipeline_state.qd), 'obs0': np.asarray(st.obs),
                'episode_length': np.array(ep_len)})
    for f, get in G.LEAVES.items():
      out[f'init_ps_{f}'] = np.asarray(get(st.pipeline_state), np.float64)
    acts = f32(rng.uniform(-1, 1, (steps, n, env.action_size)) * (0.4 if name == 'humanoid' else 1.0))
    out['act'] = acts
    for k in range(steps):
      st = env.step(st, jp.array(acts[k]))
      rec = {'obs': st.obs, 'reward': st.reward, 'done': st.done, 'steps': st.info['steps'], 'truncation': st.info['truncation'],
             'q': st.pipeline_state.q, 'qd': st.pipeline_state.qd}
      rec.update({f'metric_{m}': v for m, v in st.metrics.items()})
      for f, get in G.LEAVES.items():
        rec[f'ps_{f}'] = get(st.pipeline_state)
      for kk, v in rec.items():
        out[f'step{k}_{kk}'] = np.asarray(v, np.float64)
    path = os.path.join(ROOT, 'tests', 'golden', f'ref_dr_{name}.npz')
    np.savez_compressed(path, **out)
    spread = float(np.abs(out['step0_q'] - out['step0_q'][:1]).max())
    print(f'dr {name}: {n} envs x {steps} env-steps (episode_length {ep_len}), done per step '
          f'{[float(out[f"step{k}_done"].sum()) for k in range(steps)]}, wrote {path} ({os.path.getsize(path) // 1024} KB), q spread {spread:.3g}')


if __name__ == '__main__':
  main()
