"""`equinox` stand-in: Module = a plain attribute container whose annotated fields are pytree children
(see jax/__init__.py in this directory for why this exists)."""

import pprint


def module_fields(cls):
    names = []
    for klass in reversed(cls.__mro__):
        for name in getattr(klass, "__annotations__", {}):
            if name not in names and not name.startswith("__"):
                names.append(name)
    return tuple(names)


def field(*, default=None, static=False, converter=None, **_kw):
    return default


class Module:
    def __init_subclass__(cls, **kwargs):
        super().__init_subclass__(**kwargs)
        if "__init__" not in cls.__dict__ and not any(
            "__init__" in k.__dict__ for k in cls.__mro__[1:] if k not in (Module, object) and issubclass(k, Module)
        ):
            names = module_fields(cls)

            def __init__(self, *args, **kw):
                if len(args) > len(names):
                    raise TypeError("too many positional arguments")
                vals = dict(zip(names, args))
                for k, v in kw.items():
                    if k not in names or k in vals:
                        raise TypeError(f"unexpected or duplicate argument {k!r}")
                    vals[k] = v
                for n in names:
                    if n in vals:
                        object.__setattr__(self, n, vals[n])
                    elif hasattr(type(self), n):
                        object.__setattr__(self, n, getattr(type(self), n))
                    else:
                        raise TypeError(f"missing argument {n!r}")
                _run_checks(self)

            cls.__init__ = __init__
        else:
            own = cls.__dict__.get("__init__")
            if own is not None:

                def __init__(self, *args, _own=own, **kw):
                    _own(self, *args, **kw)
                    if type(self).__init__ is __init__:
                        _run_checks(self)

                __init__.__wrapped__ = own
                cls.__init__ = __init__


def _run_checks(obj):
    for klass in reversed(type(obj).__mro__):
        chk = klass.__dict__.get("__check_init__")
        if chk is not None:
            chk(obj)


def tree_pformat(obj, **_kw):
    return pprint.pformat({n: getattr(obj, n, None) for n in module_fields(type(obj))})
