"""Stub of the unreleased `ttns2` tensor-network package; only the names ttnsVector.py:10-15
imports exist, none is ever called on the NumpyVector path."""
