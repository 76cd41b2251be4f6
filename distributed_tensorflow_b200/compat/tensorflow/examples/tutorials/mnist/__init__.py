from . import input_data  # noqa: F401
