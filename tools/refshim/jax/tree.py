from jax.tree_util import tree_map as map, tree_leaves as leaves, tree_flatten as flatten, tree_unflatten as unflatten  # noqa: F401,A001
