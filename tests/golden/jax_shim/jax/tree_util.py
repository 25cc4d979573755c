"""Minimal pytrees: None, tuple, list, dict and equinox.Module nodes; everything else is a leaf."""


class TreeDef:
    def __init__(self, kind, meta, children):
        self.kind, self.meta, self.children = kind, meta, tuple(children)

    @property
    def num_leaves(self):
        return 1 if self.kind == "leaf" else sum(c.num_leaves for c in self.children)

    def _key(self):
        return (self.kind, self.meta, tuple(c._key() for c in self.children))

    def __eq__(self, other):
        return isinstance(other, TreeDef) and self._key() == other._key()

    def __ne__(self, other):
        return not self == other

    def __hash__(self):
        return hash(self._key())

    def __repr__(self):
        if self.kind == "leaf":
            return "*"
        return f"{self.kind}{self.meta if self.meta is not None else ''}({', '.join(map(repr, self.children))})"

    def unflatten(self, leaves):
        it = iter(leaves)
        out = self._build(it)
        return out

    def _build(self, it):
        if self.kind == "leaf":
            return next(it)
        if self.kind == "none":
            return None
        kids = [c._build(it) for c in self.children]
        if self.kind == "tuple":
            return tuple(kids)
        if self.kind == "list":
            return list(kids)
        if self.kind == "dict":
            return dict(zip(self.meta, kids))
        cls, names = self.meta
        obj = object.__new__(cls)
        for name, kid in zip(names, kids):
            object.__setattr__(obj, name, kid)
        return obj


def _children(tree):
    from equinox import Module, module_fields

    if tree is None:
        return "none", None, ()
    if isinstance(tree, tuple):
        return "tuple", None, tree
    if isinstance(tree, list):
        return "list", None, tree
    if isinstance(tree, dict):
        keys = tuple(sorted(tree))
        return "dict", keys, [tree[k] for k in keys]
    if isinstance(tree, Module):
        names = module_fields(type(tree))
        return "module", (type(tree), names), [getattr(tree, n, None) for n in names]
    return None


def tree_flatten(tree, is_leaf=None):
    leaves = []

    def rec(node):
        if is_leaf is not None and node is not None and is_leaf(node):
            leaves.append(node)
            return TreeDef("leaf", None, ())
        info = _children(node)
        if info is None:
            leaves.append(node)
            return TreeDef("leaf", None, ())
        kind, meta, kids = info
        return TreeDef(kind, meta, [rec(k) for k in kids])

    treedef = rec(tree)
    return leaves, treedef


def tree_structure(tree, is_leaf=None):
    return tree_flatten(tree, is_leaf)[1]


def tree_leaves(tree, is_leaf=None):
    return tree_flatten(tree, is_leaf)[0]


def tree_unflatten(treedef, leaves):
    return treedef.unflatten(leaves)


def tree_map(f, tree, *rest, is_leaf=None):
    leaves, treedef = tree_flatten(tree, is_leaf)
    others = [tree_flatten(r, is_leaf)[0] for r in rest]
    return treedef.unflatten([f(*xs) for xs in zip(leaves, *others)])
