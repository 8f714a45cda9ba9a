from . import time_limit  # noqa: F401
