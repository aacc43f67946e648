"""pipeline.init / step(debug=True) attach the state's contacts (reference generalized/pipeline.py:58-60,91-92)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name', ['ant', 'humanoid', 'hopper', 'pusher', 'reacher'])
def test_debug_contact_on_the_gpu(name):
  import torch
  from brax_b200 import contact, envs_assets
  from brax_b200.generalized import pipeline
  s = envs_assets.load(name)
  n = 33
  rng = np.random.default_rng(1)
  q = np.asarray(s.init_q, np.float32)[None] + rng.uniform(-0.1, 0.1, (n, s.nq)).astype(np.float32)
  if name == 'hopper':
    q[:, 1] -= 0.05
  dev = torch.device('cuda', 0)
  st = pipeline.init(s, torch.as_tensor(q, device=dev), torch.zeros((n, s.nv), device=dev), debug=True)
  ncon = len(s.contact_pairs().geom1)
  if ncon == 0:
    assert st.contact is None
  else:
    assert st.contact.dist.shape == (n, ncon) and st.contact.solver_stats is None
  assert pipeline.init(s, torch.as_tensor(q, device=dev), torch.zeros((n, s.nv), device=dev)).contact is None
  act = torch.as_tensor(rng.uniform(-1, 1, (n, s.nu)).astype(np.float32), device=dev)
  st2 = pipeline.step(s, st, act, debug=True, n_frames=3)
  c = st2.contact
  assert c.solver_stats.shape == (n, 4) and (c['stats'][:, 0] >= 3).all()      # at least one solver iteration per substep
  if ncon:
    # the kernel's distances (its own colliders, on chip) = the torch restatement on the state's link transforms
    ref = contact.get(s, st2.x)
    np.testing.assert_allclose(c.dist.cpu().numpy(), ref.dist.cpu().numpy(), rtol=1e-5, atol=2e-6)
    assert torch.equal(c.pos, ref.pos) and torch.equal(c.frame, ref.frame)
    assert c['con_dist'] is c.dist
  assert pipeline.step(s, st, act).contact is None
