import numpy as _np
from jax import numpy as jp


def segment_sum(data, segment_ids, num_segments=None, **kw):
  data = _np.asarray(data)
  n = int(num_segments if num_segments is not None else _np.max(segment_ids) + 1)
  out = _np.zeros((n,) + data.shape[1:], data.dtype)
  _np.add.at(out, _np.asarray(segment_ids), data)
  return jp._wrap(out)
