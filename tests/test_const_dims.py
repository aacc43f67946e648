"""The generated constexpr layouts (brax_b200/csrc/gen/bxg_dims_*.h) that the model-specialised kernels are compiled
with must be exactly what `pack_model` produces for the shipped Ant / Humanoid assets today: a stale header would
silently send those models back to the generic kernels (correct, slower)."""
import os

from brax_b200 import envs_assets, native
from tests.conftest import ROOT
from tools import gen_const_dims as G


def test_generated_headers_are_current():
  for model in G.MODELS:
    path = os.path.join(ROOT, 'brax_b200', 'csrc', 'gen', f'bxg_dims_{model}.h')
    assert open(path).read() == G.header(model), f'{path} is stale: run python tools/gen_const_dims.py and rebuild'


def test_ant_and_humanoid_take_the_specialised_kernels_and_nothing_else_does():
  assert native.plan(envs_assets.load('ant'))['kernel_id'] == 10
  assert native.plan(envs_assets.load('humanoid'))['kernel_id'] == 11
  for other in ('hopper', 'halfcheetah', 'walker2d', 'humanoidstandup', 'reacher'):
    p = native.plan(envs_assets.load(other))
    assert p['kernel_id'] == p['variant']
  # a different solver setting on the same topology is a different layout contract: generic kernel
  ant = envs_assets.load('ant').replace(solver_iterations=6)
  assert native.plan(ant)['kernel_id'] == native.plan(ant)['variant'] == 0
