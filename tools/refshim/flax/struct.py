"""flax.struct on plain frozen dataclasses registered as pytrees."""
import dataclasses

from jax.tree_util import register_dataclass  # noqa: F401


def field(pytree_node=True, **kw):
  md = dict(kw.pop('metadata', {}) or {})
  md['pytree_node'] = pytree_node
  return dataclasses.field(metadata=md, **kw)


def dataclass(cls=None, **kw):
  def wrap(c):
    c = dataclasses.dataclass(frozen=True)(c)
    if 'replace' not in c.__dict__:
      c.replace = lambda self, **upd: dataclasses.replace(self, **upd)
    return register_dataclass(c)
  return wrap(cls) if cls is not None else wrap


class PyTreeNode:
  def __init_subclass__(cls, **kw):
    super().__init_subclass__(**kw)
    dataclass(cls)
