import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)


def pytest_configure(config):
  config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pytest_collection_modifyitems(config, items):
  """`gpu` tests need a CUDA device and the in-tree extension: skip them (visibly) elsewhere.
  On a GPU box a missing libbxg.so is NOT skipped: those tests must fail loudly there."""
  try:
    import torch
    has_cuda = torch.cuda.is_available()
  except Exception:
    has_cuda = False
  if has_cuda:
    return
  skip = pytest.mark.skip(reason='no CUDA device (the product has no CPU fallback)')
  for item in items:
    if 'gpu' in item.keywords:
      item.add_marker(skip)


def golden(name):
  from brax_b200.io import model_json
  return model_json.load(os.path.join(ROOT, 'tests', 'golden', f'{name}.json'))


@pytest.fixture(scope='session')
def ant():
  from brax_b200 import envs_assets
  return envs_assets.load('ant')


@pytest.fixture(scope='session')
def humanoid():
  from brax_b200 import envs_assets
  return envs_assets.load('humanoid')
