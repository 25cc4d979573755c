from . import ad  # noqa: F401
