from .helpers import to_2tuple  # noqa: F401
