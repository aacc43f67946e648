"""Only what the reference envs use to name their asset: resource_path('brax') / 'envs/assets/x.xml'."""
import pathlib

Path = pathlib.PurePosixPath


def resource_path(package):
  return pathlib.PurePosixPath('/root/reference') / package
