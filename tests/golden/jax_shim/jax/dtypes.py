import numpy as _np


def result_type(*args):
    # jax with x64 enabled: Python floats are weak-typed float64
    return _np.result_type(*[_np.asarray(a).dtype if not isinstance(a, _np.dtype) else a for a in args])
