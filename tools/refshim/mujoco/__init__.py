"""Containers only: the MuJoCo C library and mujoco-mjx are not installable here."""
from mujoco import mjx  # noqa: F401


class MjModel:   # type annotation target in brax/base.py
  pass
