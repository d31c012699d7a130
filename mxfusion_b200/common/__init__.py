from . import config, exceptions, constants  # noqa: F401
