"""The kernel's own elementary functions (brax_b200/csrc/bxg_core.cuh).

`r_sincos` replaces libm's / CUDA's `sincosf` so that host emulator and device agree bit for bit;
it has to be as accurate as the functions it replaces (XLA's and libm's float32 sin / cos are good
to 1-2 ulp; so is this one over the angles the physics sees)."""
import ctypes

import numpy as np

from tests.simt import sim as S


def _sincos(x):
  S.build()
  lib = ctypes.CDLL(S._SO)
  x = np.ascontiguousarray(x, np.float32)
  s, c = np.empty_like(x), np.empty_like(x)
  lib.sim_sincos(x.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(x.size), s.ctypes.data_as(ctypes.c_void_p), c.ctypes.data_as(ctypes.c_void_p))
  return s, c


def _ulps(got, exact):
  ulp = np.spacing(np.abs(exact.astype(np.float32))).astype(np.float64)
  return np.abs(got.astype(np.float64) - exact) / ulp


def test_sincos_accuracy_over_the_physics_range():
  rng = np.random.default_rng(0)
  for lim, gate in ((0.8, 1.6), (3.2, 1.7), (100.0, 1.7), (1.0e4, 3.0)):
    x = rng.uniform(-lim, lim, 1_000_000).astype(np.float32)
    s, c = _sincos(x)
    es, ec = _ulps(s, np.sin(x.astype(np.float64))), _ulps(c, np.cos(x.astype(np.float64)))
    # near the zeros of sin / cos at large arguments the three-term reduction limits the RELATIVE error: there the
    # absolute error (<= 1.5e-7, about one ulp of 1) is what is bounded
    abs_s = np.abs(s.astype(np.float64) - np.sin(x.astype(np.float64))).max()
    abs_c = np.abs(c.astype(np.float64) - np.cos(x.astype(np.float64))).max()
    assert abs_s <= 1.5e-7 and abs_c <= 1.5e-7, (lim, abs_s, abs_c)
    if lim <= 100.0:
      assert es.max() <= gate and ec.max() <= gate, (lim, es.max(), ec.max())
    assert es.mean() < 0.4 and ec.mean() < 0.4


def test_sincos_special_values():
  x = np.array([0.0, -0.0, np.pi / 2, -np.pi / 2, np.pi, 1e-20, 3e9, -3e9, np.inf, -np.inf, np.nan], np.float32)
  s, c = _sincos(x)
  assert s[0] == 0 and c[0] == 1 and s[1] == 0 and c[1] == 1
  assert abs(s[2] - 1) < 1e-7 and abs(c[2]) < 1e-7 and abs(s[3] + 1) < 1e-7 and abs(s[4]) < 2e-7 and abs(c[4] + 1) < 1e-7
  assert s[5] == np.float32(1e-20) and c[5] == 1
  assert np.all(np.abs(s[6:8]) <= 1.0001) and np.all(np.abs(c[6:8]) <= 1.0001)      # far outside the range: only bounded
  assert np.isnan(s[8:]).all() and np.isnan(c[8:]).all()
  # sin^2 + cos^2 = 1 to rounding everywhere in range
  y = np.linspace(-50, 50, 100001).astype(np.float32)
  s, c = _sincos(y)
  assert np.abs(s.astype(np.float64) ** 2 + c.astype(np.float64) ** 2 - 1).max() < 3e-7
