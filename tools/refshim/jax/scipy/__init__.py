from jax.scipy import linalg  # noqa: F401
