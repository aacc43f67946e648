import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)


def pytest_configure(config):
  config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


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
