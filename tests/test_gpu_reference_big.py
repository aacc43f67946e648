"""The CUDA path against the reference-sized reference-source rollouts (tests/golden/ref_big_*.npz):
one-env-step maps from the reference's state at the checkpoints, substep by substep, with the
reference run's own branch decisions as the yardstick (tests/test_reference_golden_big.py holds the
float64 / emulator side and explains the gates)."""
import numpy as np
import pytest

from tests.test_reference_golden_big import F32_GATES, MODELS, SIG, SimStepper, f32_one_env_step_maps, load

pytestmark = pytest.mark.gpu


class GpuStepper:
  def __init__(self, s):
    import torch
    from brax_b200 import native
    self.torch, self.native = torch, native
    self.dev = torch.device('cuda', 0)
    self.nm = native.NativeModel(s, 0)

  def init(self, q, qd):
    t = self.torch
    st = self.nm.init(t.as_tensor(q.astype(np.float32), device=self.dev), t.as_tensor(qd.astype(np.float32), device=self.dev))
    return _Host(self, st)

  def substep(self, st, act):
    t = self.torch
    diag = self.nm.alloc_diag(act.shape[0])
    out = self.nm.step(st.bufs, t.as_tensor(act, device=self.dev), 1, diag=diag)
    return _Host(self, out), diag['stats'].cpu().numpy(), diag['con_dist'].cpu().numpy()


class _Host(dict):
  """Device State whose leaves read / write like the numpy dicts of the CPU steppers."""

  def __init__(self, stepper, bufs):
    self.bufs, self.stepper = bufs, stepper

  def __getitem__(self, k):
    return _Leaf(self.bufs[k])


class _Leaf:
  def __init__(self, t):
    self.t = t
    self.dtype = np.float32

  def __setitem__(self, idx, v):
    import torch
    self.t.copy_(torch.as_tensor(np.asarray(v, np.float32), device=self.t.device))

  def __array__(self, dtype=None, copy=None):
    a = self.t.cpu().numpy()
    return a if dtype is None else a.astype(dtype)

  def __sub__(self, o):
    return np.asarray(self) - o


@pytest.mark.parametrize('name', MODELS)
def test_cuda_path_against_the_reference_run(name):
  s, g = load(name)
  ok, err = f32_one_env_step_maps(name, GpuStepper(s))
  frac_gate, inside_gate = F32_GATES[name]
  assert ok.mean() >= frac_gate, (name, ok.mean())
  assert (err[ok] <= 1.0).mean() >= inside_gate, (name, (err[ok] <= 1.0).mean())
  assert np.median(err[ok]) <= 0.15 and np.percentile(err[ok], 90) <= 0.5, (name, np.percentile(err[ok], [50, 90]))
  # and, being the same arithmetic, exactly what the host emulation of the kernel source gives
  ok_h, err_h = f32_one_env_step_maps(name, SimStepper(s))
  np.testing.assert_array_equal(ok, ok_h)
  np.testing.assert_array_equal(err, err_h)
