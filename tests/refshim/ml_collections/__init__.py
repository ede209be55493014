from . import config_dict  # noqa: F401
