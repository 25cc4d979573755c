"""`jax.extend.linear_util`: generator-based function transformations (wrap_init / transformation)."""


class _Store:
    def __init__(self):
        self.val = None

    def __call__(self):
        return self.val


class WrappedFun:
    def __init__(self, f, transforms=()):
        self.f, self.transforms = f, tuple(transforms)

    def wrap(self, gen, static_args, store):
        return WrappedFun(self.f, ((gen, static_args, store),) + self.transforms)

    def call_wrapped(self, *args, **kwargs):
        stack = []
        for gen, static_args, store in self.transforms:
            g = gen(*(tuple(static_args) + tuple(args)), **kwargs)
            args, kwargs = next(g)
            stack.append((g, store))
        ans = self.f(*args, **kwargs)
        while stack:
            g, store = stack.pop()
            ans = g.send(ans)
            if store is not None:
                ans, store.val = ans
        return ans


def wrap_init(f, params=None, *, debug_info=None):
    return WrappedFun(f)


def transformation(gen):
    return lambda fun, *static_args: fun.wrap(gen, static_args, None)


def transformation_with_aux(gen):
    def apply(fun, *static_args):
        store = _Store()
        return fun.wrap(gen, static_args, store), store

    return apply
