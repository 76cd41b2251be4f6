from . import timeline  # noqa: F401
