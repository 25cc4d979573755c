from .tree_util import TreeDef, tree_flatten


def flatten_axes(name, treedef, axis_tree):
    """Broadcast a prefix tree of axis specs (int / None / containers of them) to one spec per leaf."""

    def rec(td, spec):
        if spec is None or isinstance(spec, int):
            return [spec] * td.num_leaves
        if td.kind == "leaf":
            raise ValueError(f"{name}: axis spec {spec!r} is deeper than the value")
        kids = list(spec.values()) if isinstance(spec, dict) else list(spec)
        if len(kids) != len(td.children):
            raise ValueError(f"{name}: axis spec {spec!r} does not match {td!r}")
        out = []
        for c, s in zip(td.children, kids):
            out += rec(c, s)
        return out

    assert isinstance(treedef, TreeDef)
    return rec(treedef, axis_tree)


def debug_info(*_args, **_kwargs):
    return None


__all__ = ["flatten_axes", "debug_info", "tree_flatten"]
