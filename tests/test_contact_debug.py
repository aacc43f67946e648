"""`brax_b200.contact.get` (what `pipeline.init / step(debug=True)` attach to the state; reference brax/contact.py:28-67)
against the oracle's colliders, for every contact kind: plane-sphere (Ant, Humanoid), plane-capsule end points
(Hopper, HalfCheetah) and capsule-capsule between moving links (Pusher)."""
import numpy as np
import pytest
import torch

from brax_b200 import base, contact, envs_assets
from oracle import oracle as O


def _states(name, n, seed=0):
  s = envs_assets.load(name)
  rng = np.random.default_rng(seed)
  q = (np.asarray(s.init_q, np.float64)[None] + rng.uniform(-0.15, 0.15, (n, s.nq)))
  if name in ('hopper', 'halfcheetah'):
    q[:, 1] -= {'hopper': 0.05, 'halfcheetah': 0.45}[name]
  if name == 'pusher':
    q[:, :7] = rng.uniform(-0.3, 0.3, (n, 7)); q[:, 1] = 0.42
  return s, q


@pytest.mark.parametrize('name', ['ant', 'humanoid', 'hopper', 'halfcheetah', 'pusher'])
def test_contact_get_matches_the_oracle_colliders(name):
  n = 16
  s, q = _states(name, n)
  o = O.Oracle(s, np.float64)
  st = o.init(q, np.zeros((n, s.nv)))
  x = base.Transform(pos=torch.as_tensor(st['x_pos']), rot=torch.as_tensor(st['x_rot']))
  c = contact.get(s, x)
  ncon = len(s.contact_pairs().geom1)
  assert c.dist.shape == (n, ncon) and c.pos.shape == (n, ncon, 3) and c.frame.shape == (n, ncon, 3, 3)
  assert c.friction.shape == (ncon, 5) and c.solref.shape == (ncon, 2) and c.solimp.shape == (ncon, 5)
  assert len(c.link_idx) == 2 and c.link_idx[1].shape == (ncon,)
  for e in range(n):
    dist, pos = o.contact(q[e])
    np.testing.assert_allclose(c.dist[e].numpy(), dist, rtol=1e-9, atol=1e-11, err_msg=f'{name} env {e} dist')
    np.testing.assert_allclose(c.pos[e].numpy(), pos, rtol=1e-9, atol=1e-11, err_msg=f'{name} env {e} pos')
  f = c.frame.numpy()
  eye = np.einsum('ncij,nckj->ncik', f, f)
  np.testing.assert_allclose(eye, np.broadcast_to(np.eye(3), eye.shape), atol=1e-9)      # orthonormal rows: normal, tangent, bitangent
  cp = s.contact_pairs()
  plane = np.asarray(cp.kind) != 2
  np.testing.assert_allclose(f[:, plane, 0], np.broadcast_to(np.asarray(cp.plane_normal)[plane], f[:, plane, 0].shape), atol=1e-12)
  assert (c.dist.numpy() < 0).any() or name == 'pusher'
  np.testing.assert_array_equal(c.link_idx[1].numpy(), np.asarray(cp.link_b))


def test_no_contacts_gives_none_and_the_counters_still_travel():
  s = envs_assets.load('reacher')
  x = base.Transform(pos=torch.zeros((3, s.num_links(), 3)), rot=torch.zeros((3, s.num_links(), 4)))
  assert contact.get(s, x) is None
  c = contact.get(s, x, solver_stats=torch.ones((3, 4), dtype=torch.int32))
  assert c.dist.shape == (3, 0) and c['stats'].shape == (3, 4)
