"""Stand-in for the slice of `brax` 0.12.1 the reference's env path uses (see ../README.md)."""
from . import base, math  # noqa: F401
