"""NumPy-backed stand-in for the subset of the `jax` API that the reference's hot path touches.

TEST INFRASTRUCTURE ONLY.  jax/jaxlib/equinox are not installable in this image (no network, no
wheel), so the reference cannot be imported as is.  With this directory and the reference's
`src/` on `sys.path`, the reference's OWN, UNMODIFIED source files (`jaxoplanet/core/kepler.py`,
`core/limb_dark.py`, `orbits/keplerian.py`, `orbits/transit.py`, `light_curves/*.py`,
`object_stack.py`, `utils.py`) execute eagerly on NumPy float64: every `jnp.*` call lands on the
NumPy function of the same name, `jit` is the identity, `vmap` is a Python loop over the mapped
axis, `custom_jvp` keeps the primal function and records the rule.  What that run pins is the
reference's formulas, operation order, branch conditions and API plumbing.  What it cannot pin is
XLA's own libm (sin/cos/atan2/cbrt/tan differ from NumPy's by <= 1 ulp) and XLA's fusion, and
derivatives (no tracing here): those stay with the oracle's torch-autograd restatement.

Nothing in the product imports this; only tests/golden/make_reference_vectors.py and
tests/test_reference_shim.py do.
"""

from __future__ import annotations

import functools

import numpy as _np

from . import numpy  # noqa: F401  (jax.numpy)
from . import api_util, dtypes, extend, interpreters, tree_util  # noqa: F401
from .tree_util import tree_flatten as _flatten


def jit(fun=None, **_kwargs):
    if fun is None:
        return lambda f: f
    return fun


class custom_jvp:  # noqa: N801  (jax's own name)
    """Keeps the primal function; `defjvp` records the rule so that tests can call it."""

    def __init__(self, fun, nondiff_argnums=()):
        self.fun = fun
        self.jvp_rule = None
        functools.update_wrapper(self, fun)

    def defjvp(self, rule, symbolic_zeros=False):
        self.jvp_rule = rule
        return rule

    def __call__(self, *args, **kwargs):
        return self.fun(*args, **kwargs)


def _axis_size(leaves_axes):
    sizes = {(_np.shape(leaf)[ax]) for leaf, ax in leaves_axes if ax is not None}
    if len(sizes) != 1:
        raise ValueError(f"vmap: inconsistent or missing mapped axis sizes {sizes}")
    return sizes.pop()


def vmap(fun, in_axes=0, out_axes=0):
    """`jax.vmap` as a Python loop: call `fun` once per index of the mapped axis and stack."""

    @functools.wraps(fun)
    def mapped(*args):
        axes = tuple(in_axes) if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        if len(axes) != len(args):
            raise ValueError("vmap: in_axes must match the positional arguments")
        flat = []  # per argument: (leaves, treedef, per-leaf axes)
        for arg, ax in zip(args, axes):
            leaves, treedef = _flatten(arg)
            leaf_axes = api_util.flatten_axes("vmap in_axes", treedef, ax)
            flat.append((leaves, treedef, leaf_axes))
        n = _axis_size([(leaf, a) for leaves, _, la in flat for leaf, a in zip(leaves, la)])
        results = []
        out_tree = None
        for i in range(n):
            call_args = []
            for leaves, treedef, leaf_axes in flat:
                picked = [
                    leaf if a is None else _np.take(_np.asarray(leaf), i, axis=a)
                    for leaf, a in zip(leaves, leaf_axes)
                ]
                call_args.append(treedef.unflatten(picked))
            out_leaves, out_tree = _flatten(fun(*call_args))
            results.append(out_leaves)
        out_axes_flat = api_util.flatten_axes("vmap out_axes", out_tree, out_axes)
        stacked = [
            parts[0] if a is None else numpy.stack([_np.asarray(p) for p in parts], axis=a)
            for a, *parts in zip(out_axes_flat, *results)
        ]
        return out_tree.unflatten(stacked)

    return mapped
