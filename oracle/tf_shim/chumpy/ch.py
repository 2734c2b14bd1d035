from . import Ch  # noqa: F401
