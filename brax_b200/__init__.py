"""brax_b200 -- B200-native `generalized` physics step with the Brax API.

    from brax_b200 import envs_assets
    from brax_b200.generalized import pipeline
    sys = envs_assets.load('ant')
    state = pipeline.init(sys, q, qd)            # q [n, nq], qd [n, nv] on cuda
    state = pipeline.step(sys, state, act, n_frames=5)
"""
__version__ = '0.1.0'
