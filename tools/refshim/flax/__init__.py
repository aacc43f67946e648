from flax import struct  # noqa: F401
