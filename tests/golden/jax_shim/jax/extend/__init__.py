from . import linear_util  # noqa: F401
