"""See README.md.  install() puts the stand-ins and the reference on sys.path (generator script only)."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def install(reference_root: str = "/root/reference") -> None:
    for p in (reference_root, os.path.join(HERE, "stubs"), HERE):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
