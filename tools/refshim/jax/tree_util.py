"""Pytrees: None / tuple / list / dict nodes and dataclasses registered by flax.struct or mjx.PyTreeNode."""
import dataclasses

_registered = set()


def register_dataclass(cls):
  _registered.add(cls)
  return cls


def _node_fields(obj):
  return [f for f in dataclasses.fields(obj) if f.metadata.get('pytree_node', True)]


def _is_node(x):
  return x is None or isinstance(x, (tuple, list, dict)) or type(x) in _registered


def tree_map(f, tree, *rest, is_leaf=None):
  if is_leaf is not None and is_leaf(tree):
    return f(tree, *rest)
  if tree is None:
    return None
  if isinstance(tree, (tuple, list)):
    out = [tree_map(f, t, *[r[i] for r in rest], is_leaf=is_leaf) for i, t in enumerate(tree)]
    if hasattr(tree, '_fields'):
      return type(tree)(*out)
    return type(tree)(out)
  if isinstance(tree, dict):
    return {k: tree_map(f, v, *[r[k] for r in rest], is_leaf=is_leaf) for k, v in tree.items()}
  if type(tree) in _registered:
    kw = {fl.name: tree_map(f, getattr(tree, fl.name), *[getattr(r, fl.name) for r in rest], is_leaf=is_leaf)
          for fl in _node_fields(tree)}
    return dataclasses.replace(tree, **kw)
  return f(tree, *rest)


def tree_leaves(tree, is_leaf=None):
  out = []
  tree_map(lambda x: out.append(x), tree, is_leaf=is_leaf)
  return out


def tree_flatten(tree, is_leaf=None):
  return tree_leaves(tree, is_leaf), tree


def tree_unflatten(treedef, leaves):
  it = iter(leaves)
  return tree_map(lambda _: next(it), treedef)


def tree_structure(tree):
  """A comparable description of the container structure (leaves replaced by '*')."""
  if tree is None:
    return None
  if isinstance(tree, (tuple, list)):
    return (type(tree).__name__, tuple(tree_structure(t) for t in tree))
  if isinstance(tree, dict):
    return ('dict', tuple((k, tree_structure(v)) for k, v in sorted(tree.items())))
  if type(tree) in _registered:
    return (type(tree).__name__, tuple((fl.name, tree_structure(getattr(tree, fl.name))) for fl in _node_fields(tree)))
  return '*'
