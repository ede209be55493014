"""Stand-in for the slice of `jax` the reference's env path uses (see ../README.md)."""
import numpy as _np

from . import nn, numpy, random  # noqa: F401

Array = _np.ndarray


def jit(f, *a, **k):
    return f


def vmap(f, in_axes=0, out_axes=0):
    """Loop over the leading axis of every argument, stack the (tuple of) results."""
    def g(*args):
        n = len(args[0])
        outs = [f(*[numpy.asarray(x[i]) for x in args]) for i in range(n)]
        if isinstance(outs[0], tuple):
            return tuple(numpy.stack([o[j] for o in outs]) for j in range(len(outs[0])))
        return numpy.stack(outs)
    return g


def tree_map(f, tree):
    return tree.tree_map(f)
