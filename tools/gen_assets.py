"""Regenerates the compiled model constants shipped with the package.

Run in the build container (needs /root/reference for the MJCF sources):
    python tools/gen_assets.py
Writes brax_b200/assets/{ant,humanoid,halfcheetah,hopper,walker2d,humanoidstandup,pusher,inverted_pendulum,inverted_double_pendulum,
reacher,swimmer}.json and the small pendulum fixtures under tests/golden/ that the known-answer tests use.  The XML files themselves
are not copied into this repository; only the numbers the hot path consumes.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from brax_b200.io import mjcf, model_json  # noqa: E402

REF = '/root/reference/brax'
ASSETS = {
    'ant': f'{REF}/envs/assets/ant.xml',
    'humanoid': f'{REF}/envs/assets/humanoid.xml',
    # plane-capsule models
    'halfcheetah': f'{REF}/envs/assets/half_cheetah.xml',
    'hopper': f'{REF}/envs/assets/hopper.xml',
    'walker2d': f'{REF}/envs/assets/walker2d.xml',
    'pusher': f'{REF}/envs/assets/pusher.xml',   # capsule-capsule pairs between moving links
    'humanoidstandup': f'{REF}/envs/assets/humanoidstandup.xml',   # 15 contacts: 77 constraint rows (generic kernel variant)
}
FIXTURES = ['triple_pendulum', 'single_pendulum_motor', 'single_pendulum_position',
            'single_pendulum_velocity', 'single_pendulum_position_frclimit',
            'double_pendulum', 'single_pendulum', 'triple_pendulum_motor']

# the classic-control env models (slide joints, 2-dof links, no contacts; Swimmer: fluid forces): shipped as
# assets and kept as parity fixtures
ENV_FIXTURES = ['inverted_pendulum', 'inverted_double_pendulum', 'reacher', 'swimmer']

if __name__ == '__main__':
  for name, path in ASSETS.items():
    out = os.path.join(ROOT, 'brax_b200', 'assets', f'{name}.json')
    model_json.save(mjcf.load(path), out)
    print('wrote', out)
  for name in FIXTURES:
    out = os.path.join(ROOT, 'tests', 'golden', f'{name}.json')
    model_json.save(mjcf.load(f'{REF}/test_data/{name}.xml'), out)
    print('wrote', out)
  for name in ENV_FIXTURES:
    out = os.path.join(ROOT, 'tests', 'golden', f'{name}.json')
    model_json.save(mjcf.load(f'{REF}/envs/assets/{name}.xml'), out)
    print('wrote', out)
    out = os.path.join(ROOT, 'brax_b200', 'assets', f'{name}.json')
    model_json.save(mjcf.load(f'{REF}/envs/assets/{name}.xml'), out)
    print('wrote', out)
