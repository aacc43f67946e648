"""Small rollout for compute-sanitizer (memcheck / racecheck / initcheck):
  compute-sanitizer --tool racecheck python tools/sanitize.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from brax_b200 import envs, workloads
from brax_b200.generalized import pipeline

dev = torch.device('cuda', 0)
for model in ('ant', 'humanoid'):
  sys_, q, qd = workloads.reset(model, 0, 37, 0, dev)
  st = pipeline.init(sys_, q, qd)
  for k in range(2):
    st = pipeline.step(sys_, st, workloads.action(model, 0, 37, 0, k, dev), n_frames=5)
  env = envs.create(model, episode_length=2, auto_reset=True, batch_size=21)
  es = env.reset(0)
  for k in range(3):
    es = env.step(es, torch.zeros((21, env.action_size), device=dev))
  torch.cuda.synchronize()
  assert torch.isfinite(st.q).all() and torch.isfinite(es.obs).all()
from brax_b200 import envs_assets
import numpy as np
for model in ('hopper', 'halfcheetah'):   # plane-capsule contacts: half-warp variant and the generic kernel
  sys_ = envs_assets.load(model)
  q = torch.as_tensor(np.asarray(sys_.init_q, np.float32), device=dev)[None].repeat(29, 1).contiguous()
  q[:, 1] -= 0.3 if model == 'halfcheetah' else 0.04
  st = pipeline.init(sys_, q, torch.zeros((29, sys_.nv), device=dev))
  for k in range(2):
    st = pipeline.step(sys_, st, torch.zeros((29, sys_.nu), device=dev), n_frames=5)
  torch.cuda.synchronize()
  assert torch.isfinite(st.q).all()
# fluid forces (generic variant), capsule-capsule contacts (variant 5), the 80-row variant, the classic-control env kinds
for model in ('swimmer', 'pusher', 'humanoidstandup', 'reacher', 'inverted_pendulum', 'inverted_double_pendulum'):
  env = envs.create(model, episode_length=2, auto_reset=True, batch_size=23)
  es = env.reset(0)
  if model == 'pusher':   # arm lowered onto the table, the object under the wrist: both contact kinds active
    q = es.pipeline_state.q.clone(); q[:, :7] = 0.0; q[:, 1] = 0.42
    w = pipeline.init(env.sys, q, es.pipeline_state.qd).x.pos[:, 6]
    q[:, 7], q[:, 8] = w[:, 1] + 0.07, w[:, 0] - 0.40
    es = es.replace(pipeline_state=pipeline.init(env.sys, q.contiguous(), es.pipeline_state.qd))
  for k in range(3):
    es = env.step(es, 0.5 * torch.ones((23, env.action_size), device=dev))
  torch.cuda.synchronize()
  assert torch.isfinite(es.obs).all() and torch.isfinite(es.pipeline_state.q).all(), model
print('sanitize rollout ok')
# round 2: lean state I/O (separate kernels), the model-specialised kernels (ids 10 / 11: Ant and Humanoid above run on
# them; BXG_NO_SPECIALISE covers their generic twins) and the PPO kernels
from brax_b200 import native
from brax_b200.training import fused, ppo
for model in ('ant', 'humanoid'):
  env = envs.create(model, episode_length=2, auto_reset=True, batch_size=19, lean=True)
  es = env.reset(0)
  for k in range(3):
    es = env.step(es, torch.zeros((19, env.action_size), device=dev))
  torch.cuda.synchronize()
  assert torch.isfinite(es.obs).all()
os.environ['BXG_NO_SPECIALISE'] = '1'
for model in ('ant', 'humanoid'):
  sys_, q, qd = workloads.reset(model, 0, 37, 0, dev)
  nm = native.NativeModel(sys_, 0)
  assert nm.kernel_id == native.plan(sys_)['variant']
  st = nm.init(q, qd)
  st = nm.step(st, workloads.action(model, 0, 37, 0, 0, dev), 5)
  torch.cuda.synchronize()
  assert torch.isfinite(st['q']).all()
del os.environ['BXG_NO_SPECIALISE']
a = ppo.Agent(27, 8).to(dev)
act, logits, pre = a.act(torch.randn((77, 27), device=dev))
T, B = 5, 130
td = {'obs': torch.randn((T + 1, B, 27), device=dev), 'logits': torch.randn((T, B, 16), device=dev), 'pre': torch.randn((T, B, 8), device=dev),
      'reward': torch.randn((T, B), device=dev), 'done': torch.zeros((T, B), device=dev), 'truncation': torch.zeros((T, B), device=dev)}
a.loss(td).backward()
torch.cuda.synchronize()
print('sanitize: done')
# per-env models (domain randomisation): the kernels read each env's constants from global memory
from brax_b200 import base
for model in ('ant', 'humanoid', 'hopper'):   # variants 0, 1 and (hopper's own variant has no per-env build) the generic 3
  sys_ = envs_assets.load(model)
  n = 27
  scale = np.random.default_rng(0).uniform(0.8, 1.25, (n, sys_.num_links())).astype(np.float32)
  leaves = {'link.inertia.mass': np.asarray(sys_.link.inertia.mass, np.float32)[None] * scale,
            'link.inertia.i': np.asarray(sys_.link.inertia.i, np.float32)[None] * scale[:, :, None, None]}
  in_axes = base.tree_map(lambda x: None, sys_).tree_replace({k: 0 for k in leaves})
  nm = native.BatchedNativeModel(base.unbatch(sys_.tree_replace(leaves), in_axes), 0)
  assert nm.kernel_id == {'ant': 0, 'humanoid': 1, 'hopper': 3}[model], nm.kernel_id
  q = torch.as_tensor(np.asarray(sys_.init_q, np.float32), device=dev)[None].repeat(n, 1).contiguous()
  if model == 'hopper':
    q[:, 1] -= 0.04
  st = nm.init(q, torch.zeros((n, sys_.nv), device=dev))
  lean = {k: st[k].clone() for k in native.LEAN_FIELDS}
  for k in range(2):
    a_ = 0.3 * torch.ones((n, sys_.nu), device=dev)
    st = nm.step(st, a_, 5)
    lean = nm.step(lean, a_, 5, lean=True)
  torch.cuda.synchronize()
  assert torch.isfinite(st['q']).all() and torch.equal(st['q'], lean['q']), model
print('sanitize: per-env models done')
