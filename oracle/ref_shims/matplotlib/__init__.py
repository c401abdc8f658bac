"""Stub of matplotlib for importing the reference (see ../README.md)."""
