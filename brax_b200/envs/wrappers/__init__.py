"""Reference `brax/envs/wrappers/`: only the training wrappers exist here, and they are not
objects around the env but switches on it (their arithmetic is fused into the step kernel)."""
from brax_b200.envs.wrappers import training  # noqa: F401
