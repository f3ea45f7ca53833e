"""Stub of the in-house `magic` module (feast.py:13): an IPython-shell debugging hook."""


def ipsh(*args, **kwargs):
    return None
