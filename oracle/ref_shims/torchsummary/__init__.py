"""Stub of torchsummary for importing the reference (see ../README.md)."""


def summary(*a, **kw):
    return None
