"""Stub of matplotlib.pyplot for importing the reference (see ../README.md)."""


def subplots(*a, **kw):
    raise NotImplementedError("matplotlib is stubbed")


def tight_layout(*a, **kw):
    raise NotImplementedError("matplotlib is stubbed")
