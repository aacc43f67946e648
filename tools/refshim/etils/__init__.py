from etils import epath  # noqa: F401
